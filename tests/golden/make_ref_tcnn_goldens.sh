#!/bin/bash
# Regenerates tests/golden/ref_tcnn_*.bin.gz and ref_field_*.bin.gz: outputs of the REFERENCE network code (ngp::NerfNetwork<__half> on the
# reference's tiny-cuda-nn, compiled from /root/reference by oracle/ref/Makefile) run on a B200.
#   here      : make -C oracle/ref tcnn harness            (needs /root/reference)
#   GPU box   : gpurun -- 'bash tests/golden/make_ref_tcnn_goldens.sh gpurun_out'
#   here again: gzip -9n gpurun_out/ref_{tcnn,field}_*.bin && mv gpurun_out/ref_{tcnn,field}_*.bin.gz tests/golden/
set -e
out=${1:-gpurun_out}
mkdir -p "$out"
./oracle/_ref/ref_tcnn "$out"
ls -la "$out"/ref_tcnn_*.bin "$out"/ref_field_*.bin
