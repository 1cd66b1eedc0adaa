#!/usr/bin/env bash
# PSNR@1000 on nerf/fox against the order in which rays take their slots in the compacted batch (profiles/r2/psnr_ab.md, "Round 2, second half"):
# the default (groups of 32 consecutive rays, shuffled), one atomic per ray, ascending ray id, the reference's generator capacity, one thread per ray.
N=${1:-4}
tools/psnr_ab.sh "" $N
tools/psnr_ab.sh "nerf.training.compaction_order=1" $N
tools/psnr_ab.sh "nerf.training.compaction_order=2" $N
tools/psnr_ab.sh "nerf.training.drop_overflowing_rays=1" $N
tools/psnr_ab.sh "nerf.training.compaction_order=1 nerf.training.drop_overflowing_rays=1 nerf.training.gen_lanes_per_ray=1" $N
