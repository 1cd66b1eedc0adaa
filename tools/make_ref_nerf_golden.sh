#!/bin/bash
# Runs the REFERENCE's own NeRF kernels (oracle/_ref/ref_nerf, built by `make -C oracle/ref nerf` in the build container from
# /root/reference) over the seeded cases of tools/ref_nerf_cases.py on a GPU box and packs what they produced:
#   gpurun -- 'bash tools/make_ref_nerf_golden.sh [case ...]'   ->   gpurun_out/ref_nerf_<case>.npz
# tests/golden/make_ref_nerf_goldens.py then reduces those to the committed tests/golden/ref_nerf_<case>.npz.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=/tmp/ref_nerf_cases
rm -rf $R
python tools/ref_nerf_cases.py write $R "$@" || exit 1
for c in $R/*; do
	echo "== $(basename $c)"
	timeout 60 oracle/_ref/ref_nerf $c 2>&1 | tail -8 || echo "FAILED $c"
done
python tools/ref_nerf_cases.py pack $R gpurun_out
