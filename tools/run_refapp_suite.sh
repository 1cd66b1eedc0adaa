#!/usr/bin/env bash
# GPU-box driver for tools/ref_app.py: the reference application and this repo on the same scenes, back to back.
# usage: tools/run_refapp_suite.sh [scene] [steps]   -> gpurun_out/refapp/*.json|npz|ingp
set -u
SCENE=${1:-fox}; STEPS=${2:-1000}
O=gpurun_out/refapp; mkdir -p $O
run() { echo "=== $*"; timeout 600 python tools/ref_app.py "$@" 2>&1 | tail -4; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
run --impl reference --scene $SCENE --enc L16F2 --jit 1 --train-mode Nerf     --steps $STEPS --out $O/ref_${SCENE}_L16F2_jit1_nerf.json
run --impl reference --scene $SCENE --enc L16F2 --jit 0 --train-mode Nerf     --steps $STEPS --out $O/ref_${SCENE}_L16F2_jit0_nerf.json
run --impl reference --scene $SCENE --enc L8F4  --jit 1 --train-mode Nerf     --steps $STEPS --out $O/ref_${SCENE}_L8F4_jit1_nerf.json
run --impl reference --scene $SCENE --enc L8F4  --jit 1 --train-mode RflRelax --steps $STEPS --out $O/ref_${SCENE}_L8F4_jit1_rflrelax.json
run --impl ngp_b200  --scene $SCENE --enc L16F2 --train-mode Nerf --steps $STEPS --out $O/b200_${SCENE}_L16F2_nerf.json
run --impl ngp_b200  --scene $SCENE --enc L8F4  --train-mode Nerf --steps $STEPS --out $O/b200_${SCENE}_L8F4_nerf.json
# pixel parity on identical weights: each side renders the other's snapshot
run --impl ngp_b200  --scene $SCENE --enc L16F2 --no-train --load-snapshot $O/ref_${SCENE}_L16F2_jit1_nerf.ingp --out $O/b200_on_refsnap_${SCENE}_L16F2.json
run --impl reference --scene $SCENE --enc L16F2 --no-train --load-snapshot $O/b200_${SCENE}_L16F2_nerf.ingp --out $O/ref_on_b200snap_${SCENE}_L16F2.json
run --impl reference --scene $SCENE --enc L16F2 --no-train --load-snapshot $O/ref_${SCENE}_L16F2_jit1_nerf.ingp --out $O/ref_on_refsnap_${SCENE}_L16F2.json
rm -f $O/*.ingp   # keep the merge-back under 64 MiB (33 MB per snapshot)
ls -la $O
