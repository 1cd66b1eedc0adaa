#!/usr/bin/env bash
# fifth A/B of the PSNR investigation: compaction order by groups of 32 rays (shuffled) and by ray id
tools/psnr_ab.sh "" 4
tools/psnr_ab.sh "nerf.training.compaction_order=2" 3
