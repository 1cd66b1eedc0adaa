#!/usr/bin/env bash
# Builds the UNMODIFIED reference (NVlabs/instant-ngp, headless) for sm_100 from a scratch copy of /root/reference and
# installs the artefacts a GPU box needs into baseline/_ref/ (git-ignored, travels with gpurun):
#   baseline/_ref/pyngp*.so, baseline/_ref/configs/, baseline/_ref/data/nerf/fox, data/image
# Recipe = BASELINE.md section 3.1 / SURVEY.md section 8c.  Nothing from the reference enters the git history.
set -euo pipefail
REF=${REF:-/root/reference}
SCRATCH=${SCRATCH:-/tmp/refbuild}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
JOBS=${JOBS:-$(nproc)}
[ -d "$REF" ] || { echo "no $REF: using prebuilt $OUT" >&2; exit 0; }
mkdir -p "$SCRATCH" "$OUT"
# the scene and the configs first (cheap; bench.py needs data/nerf/fox even when the application itself cannot be built in time)
rm -rf "$OUT/configs" "$OUT/data"
cp -r "$REF/configs" "$OUT/configs"
mkdir -p "$OUT/data/nerf" "$OUT/data/image"
cp -r "$REF/data/nerf/fox" "$OUT/data/nerf/fox"
cp -r "$REF"/data/image/* "$OUT/data/image/" || true
chmod -R u+w "$OUT/configs" "$OUT/data"
[ "${DATA_ONLY:-0}" = "1" ] && { echo "data installed in $OUT"; exit 0; }
if [ ! -d "$SCRATCH/src" ]; then cp -r "$REF" "$SCRATCH/src"; chmod -R u+w "$SCRATCH/src"; fi
export TCNN_CUDA_ARCHITECTURES=100
cmake -S "$SCRATCH/src" -B "$SCRATCH/build" -G Ninja -DNGP_BUILD_WITH_GUI=OFF -DCMAKE_BUILD_TYPE=Release \
      -DCMAKE_CUDA_ARCHITECTURES=100 > "$SCRATCH/cmake.log" 2>&1
cmake --build "$SCRATCH/build" -j "$JOBS" > "$SCRATCH/build.log" 2>&1
cp "$SCRATCH"/build/pyngp*.so "$OUT/"
echo "reference installed in $OUT"; ls -la "$OUT"
