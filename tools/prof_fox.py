#!/usr/bin/env python
"""prof_fox.py — per-phase device times of this repo's Testbed on nerf/fox (train split), plus the 1920x1080 render, for profiling runs:

    python tools/prof_fox.py [--enc L16F2|L8F4] [--steps 1000] [--profile-steps 32] [--render 5] [--option name=value ...]

Prints one JSON line.  Under ncu, `--nvtx-range`-free: use `--steps`/`--profile-steps` to place the launches of interest."""
from __future__ import annotations

import argparse
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--enc", default="L16F2")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--profile-steps", type=int, default=32)
    ap.add_argument("--render", type=int, default=5)
    ap.add_argument("--scene", default="fox", choices=["fox", "ball"])
    ap.add_argument("--option", action="append", default=[])
    ap.add_argument("--scatter-aggregation", type=int, default=1)
    args = ap.parse_args()
    import ref_app as R

    impl = R.B200(False, "Nerf")
    tb = impl.tb
    if args.scatter_aggregation != 1:
        from importlib import import_module
        import_module("instant-ngp_b200.binding").lib().ngp_set_scatter_aggregation(args.scatter_aggregation)
    if args.scene == "fox":
        split, _ = R.fox_split()
        impl.load_transforms(split["train"])
    else:
        ball = R.ball_scene()
        impl.load_arrays(ball["imgs"], ball["cams"], ball["focal"])
    impl.set_network(R.network_config(args.enc))
    for kv in args.option:
        k, v = kv.split("=")
        tb._set(k, float(v))
    tb.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tb.train(R.BATCH)
    tb.sync()
    t1 = time.perf_counter()
    tail = max(1, args.steps // 2)
    t2 = time.perf_counter()
    for _ in range(tail):
        tb.train(R.BATCH)
    tb.sync()
    t3 = time.perf_counter()
    rec = {"scene": args.scene, "enc": args.enc, "ms_per_step_first": (t1 - t0) / max(args.steps, 1) * 1e3, "ms_per_step_steady": (t3 - t2) / tail * 1e3,
           "counters": tb.counters(), "options": args.option,
           "scatter_aggregation": args.scatter_aggregation}
    tb.set_profiling(True)
    for _ in range(args.profile_steps):
        tb.train(R.BATCH)
    tb.sync()
    ph = tb.phase_ms()
    tb.set_profiling(False)
    n = max(ph.pop("steps"), 1)
    rec["phase_ms"] = {k: round(v / n, 4) for k, v in ph.items()}
    if args.render:
        impl.prepare_eval()
        tb.set_camera_to_training_view(0)
        ts = []
        for _ in range(args.render):
            a = time.perf_counter()
            img = tb.render(1920, 1080, 1, True)
            ts.append((time.perf_counter() - a) * 1e3)
        rec["render_ms"] = [round(t, 2) for t in ts]
        rec["render_coverage"] = float((img[..., 3] > 0.5).mean())
        rec["render_steps"] = int(getattr(tb, "last_render_steps", 0))
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
