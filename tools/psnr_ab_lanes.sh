#!/usr/bin/env bash
# PSNR@1000 on nerf/fox against the generator's lanes per ray (per-warp slot reservation in the generator, per-CTA in the loss kernel)
N=${1:-2}
for G in 2 8 32; do tools/psnr_ab.sh "nerf.training.slot_reservation=3 nerf.training.gen_lanes_per_ray=$G" $N; done
tools/psnr_ab.sh "nerf.training.slot_reservation=0 nerf.training.gen_lanes_per_ray=1" $N
tools/psnr_ab.sh "nerf.training.slot_reservation=3" 3
