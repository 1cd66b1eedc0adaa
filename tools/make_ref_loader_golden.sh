#!/bin/bash
# Regenerates tests/golden/ref_loader.json on a GPU box: the REFERENCE's own ngp::load_nerf (oracle/_ref/ref_loader, built by
# `make -C oracle/ref loader` in the build container from /root/reference) over the scenes tests/loader_scenes.py writes.
#   gpurun -- 'bash tools/make_ref_loader_golden.sh'  &&  cp gpurun_out/ref_loader.json tests/golden/ref_loader.json
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=/tmp/loader_scenes
rm -rf $S
NAMES=$(python tests/loader_scenes.py $S 2>/dev/null)
ARGS=""
for n in $NAMES; do ARGS="$ARGS $S/$n"; done
oracle/_ref/ref_loader gpurun_out/ref_loader.json $ARGS 2> gpurun_out/ref_loader.stderr || { tail -c 2000 gpurun_out/ref_loader.stderr; exit 1; }
python - <<'P'
import json
d = json.load(open("gpurun_out/ref_loader.json"))
for k, v in d.items():
    print(k, "ERROR " + v["error"] if "error" in v else f"{v['n_images']} images, scale {v['scale']}, aabb_scale {v['aabb_scale']}")
P
