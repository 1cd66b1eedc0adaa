"""Reduces gpurun_out/ref_nerf_<case>.npz (what the REFERENCE's own NeRF kernels produced on a B200: tools/make_ref_nerf_golden.sh ->
oracle/_ref/ref_nerf, oracle/ref/ref_nerf_harness.cu) to the committed tests/golden/ref_nerf_<case>.npz:

  * train cases keep everything the comparison needs (ray list, step counts, coordinates, per-variant counts / losses / gradients);
    the compacted coordinates (plain copies of the inputs) only for the smallest case;
  * grid cases replace the sample positions and cell indices — which the oracle reproduces bit for bit — by their SHA-256 plus the
    first 4096 rows, and store a re-marked grid (mark_untrained_density_grid on a populated grid) as its difference from the grid
    before.

    bash tools/make_ref_nerf_golden.sh   (on a GPU box, through gpurun)   &&   python tests/golden/make_ref_nerf_goldens.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
src = ROOT / "gpurun_out"
dst = Path(__file__).resolve().parent
files = sorted(src.glob("ref_nerf_*.npz"))
if not files:
    sys.exit("gpurun_out/ref_nerf_*.npz missing: run tools/make_ref_nerf_golden.sh on a GPU box first")
for f in files:
    g = np.load(f)
    out = {}
    for k in g.files:
        a = g[k]
        if k == "loss0_coords" and f.stem != "ref_nerf_train_aabb4":
            continue
        if k.endswith("_positions") or k.endswith("_indices") and k.startswith("grid"):
            out[k + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
            out[k + "_head"] = (a.reshape(-1, 3) if k.endswith("_positions") else a)[:4096].copy()
            continue
        if k.endswith("_marked") and not k.startswith("grid0"):
            prev = g[f"grid{int(k[4]) - 1}_grid"]
            changed = np.flatnonzero(a != prev).astype(np.uint32)
            out[k + "_changed_idx"] = changed
            out[k + "_changed_val"] = a[changed]
            continue
        out[k] = a
    np.savez_compressed(dst / f.name, **out)
    print(f.name, (dst / f.name).stat().st_size, "bytes")
