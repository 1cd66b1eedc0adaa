#!/usr/bin/env bash
# fourth A/B of the PSNR investigation: shuffled compaction order / no generator drops (the new defaults) against each knob's old setting
tools/psnr_ab.sh "" 4
tools/psnr_ab.sh "nerf.training.drop_overflowing_rays=1" 2
tools/psnr_ab.sh "nerf.training.compaction_order=1" 2
