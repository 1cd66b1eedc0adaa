#!/usr/bin/env python
"""step_consistency.py — does ONE training step (generator .. loss .. forward/backward) give the same gradients whatever the schedule?

nerf/fox is trained with one thread per ray to a list of checkpoints; at each, the exact state (.ngpb: fp32 weights, optimizer, occupancy
grid, RNG streams, controller) is written out.  Every variant then reloads the checkpoint, runs `train_compute_grads` once and its fp16
gradient buffer is compared with the one-thread-per-ray run's (twice, which gives the noise floor of the fp16 reductions' order).
Variants: lanes per ray of the generator, compaction order (0 groups of 32 rays shuffled, 1 one atomic per ray, 2 ray id), the reference's
generator capacity (drop), the inference schedule.  (profiles/r2/order/step_consistency.json was taken with the experiment's `slot`
knob in place of `order`: 1 = one atomic per warp / ray, 0 = one per CTA of 32 / 16 rays.)

    python tools/step_consistency.py --checkpoints 0,8,32,100,300,700 > gpurun_out/step_consistency.json
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

BATCH = 1 << 18


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoints", default="0,8,32,100,300,700")
    ap.add_argument("--tmp", default="/tmp/step_consistency")
    args = ap.parse_args()
    import ref_app as R

    cks = sorted(int(x) for x in args.checkpoints.split(","))
    tmp = Path(args.tmp)
    tmp.mkdir(parents=True, exist_ok=True)
    cudart = C.CDLL("libcudart.so.12") if not Path("/usr/local/cuda/lib64/libcudart.so").exists() else C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]

    impl = R.B200(False, "Nerf")
    split, _ = R.fox_split()
    impl.load_transforms(split["train"])
    impl.set_network(R.network_config("L16F2"))
    tb = impl.tb
    B = importlib.import_module("instant-ngp_b200.binding")
    n_params = int(B.lib().ngp_testbed_n_params(tb._h))
    desc = tb.desc()
    n_mlp = n_params - int(desc.grid.n_params)

    def set_variant(v):
        tb._set("nerf.training.gen_lanes_per_ray", v["lanes"])
        tb._set("nerf.training.compaction_order", v["order"])
        tb._set("nerf.training.drop_overflowing_rays", v.get("drop", 0))
        tb._set("nerf.training.full_inference", v.get("full_inference", 2))

    tb._set("nerf.training.gen_lanes_per_ray", 1)
    tb._set("nerf.training.compaction_order", 1)
    step = 0
    for ck in cks:
        while step < ck:
            impl.train()
            step += 1
        impl.sync()
        tb.save_snapshot(str(tmp / f"ck_{ck}.ngpb"))

    def one_step(ck, v):
        tb.load_snapshot(str(tmp / f"ck_{ck}.ngpb"))
        impl._apply_modes()
        set_variant(v)
        before = tb.counters()
        tb.train_compute_grads(BATCH)
        tb.sync()
        g = np.empty(n_params, dtype=np.float16)
        rc = cudart.cudaMemcpy(g.ctypes.data, C.c_void_p(tb.grads_ptr()), n_params * 2, 2)
        assert rc == 0, rc
        after = tb.counters()
        tb.train_apply_grads()
        tb.sync()
        return g.astype(np.float32), before, after

    variants = [dict(lanes=1, order=1), dict(lanes=1, order=0), dict(lanes=2, order=1), dict(lanes=16, order=1), dict(lanes=16, order=0), dict(lanes=16, order=2),
                dict(lanes=0, order=0), dict(lanes=0, order=0, drop=1), dict(lanes=1, order=1, full_inference=1), dict(lanes=16, order=1, full_inference=1),
                dict(lanes=16, order=1, full_inference=0)]
    out = {"n_params": n_params, "n_mlp": n_mlp, "checkpoints": []}
    for ck in cks:
        base, b0, a0 = one_step(ck, dict(lanes=1, order=1))
        nb = float(np.linalg.norm(base))
        rec = {"step": ck, "controller_before": b0, "controller_after": a0, "grad_norm": nb, "grad_norm_mlp": float(np.linalg.norm(base[:n_mlp])),
               "nonfinite": int((~np.isfinite(base)).sum()), "variants": []}
        for v in variants:
            g, b, a = one_step(ck, v)
            d = g - base
            rec["variants"].append(dict(v, rel_diff=float(np.linalg.norm(d) / max(nb, 1e-30)), rel_diff_mlp=float(np.linalg.norm(d[:n_mlp]) / max(np.linalg.norm(base[:n_mlp]), 1e-30)),
                                        max_abs_diff=float(np.abs(d).max()), grad_norm=float(np.linalg.norm(g)), nonfinite=int((~np.isfinite(g)).sum()),
                                        n_entries_differ=int((d != 0).sum()), measured=a["measured_batch_size"], measured_before=a["measured_batch_size_before_compaction"],
                                        rays_per_batch_next=a["rays_per_batch"]))
        out["checkpoints"].append(rec)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
