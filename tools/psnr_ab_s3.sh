#!/usr/bin/env bash
# third A/B of the PSNR investigation: does waiting for the device after every step change anything?  Held-out views kept as small images.
export KEEP_NPZ=1
EXTRA="--save-views" tools/psnr_ab.sh "nerf.training.slot_reservation=1" 2
EXTRA="--save-views --sync-every-step" tools/psnr_ab.sh "nerf.training.slot_reservation=1" 3
EXTRA="--save-views" tools/psnr_ab.sh "nerf.training.slot_reservation=1 nerf.training.gen_lanes_per_ray=1" 1
