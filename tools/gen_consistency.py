#!/usr/bin/env python
"""gen_consistency.py — do all schedules of the training sample generator produce the same rays on a TRAINED scene?

tests/test_gpu_march.py pins every schedule (lanes per ray, walk, speculation) against the CPU oracle on synthetic
scenes whose occupancy is a sphere.  This tool repeats the comparison where training runs: nerf/fox after `--steps` training steps
(fragmented three-cascade occupancy, OpenCV lens, cone stepping), both arithmetic flavours, every schedule against one thread per ray.

    python tools/gen_consistency.py [--steps 600] [--rays 3584] > gpurun_out/gen_consistency.json
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
sys.path.insert(0, str(ROOT / "tests"))


def pcg32_seed(initstate: int, initseq: int = 1):
    """pcg32::seed (pcg32.h): state, increment"""
    M, mul = (1 << 64) - 1, 6364136223846793005
    inc = ((initseq << 1) | 1) & M
    state = (0 * mul + inc) & M
    state = (state + initstate) & M
    state = (state * mul + inc) & M
    return state, inc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--rays", type=int, default=3584)
    ap.add_argument("--seeds", type=int, default=3)
    args = ap.parse_args()
    import torch

    import ref_app as R
    import util

    impl = R.B200(False, "Nerf")
    split, _ = R.fox_split()
    impl.load_transforms(split["train"])
    impl.set_network(R.network_config("L16F2"))
    impl.tb._set("nerf.training.gen_lanes_per_ray", 1)
    for _ in range(args.steps):
        impl.train()
    impl.sync()
    B = importlib.import_module("instant-ngp_b200.binding")
    lib, h = B.lib(), impl.tb._h
    n_views = impl.n_views()
    arr = (B.TrainView * n_views)()
    for i in range(n_views):
        B.check(lib.ngp_testbed_get_view(h, i, C.byref(arr[i])))
    t_views = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).cuda()
    n_bf = 128 ** 3 // 8 * 8
    bf = np.zeros(n_bf, dtype=np.uint8)
    B.check(lib.ngp_testbed_get_density_grid(h, None, 0, bf.ctypes.data_as(C.c_void_p), n_bf))
    t_bf = torch.from_numpy(bf).cuda()
    occupancy = [float(np.unpackbits(bf[c * 128 ** 3 // 8:(c + 1) * 128 ** 3 // 8]).mean()) for c in range(3)]

    n_rays, max_samples = args.rays, args.rays * 1024
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    t_cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    t_ri = torch.zeros(n_rays, dtype=torch.int32, device="cuda")
    t_rays = torch.zeros(n_rays, 6, dtype=torch.float32, device="cuda")
    t_ns = torch.zeros(n_rays, 2, dtype=torch.int32, device="cuda")
    t_co = torch.zeros(max_samples, 7, dtype=torch.float32, device="cuda")

    def run(seed, math_mode, lanes, walk=0, spec=0):
        cfg = util.make_train_cfg(aabb_scale=4)
        cfg.math_mode, cfg.gen_lanes_per_ray, cfg.gen_walk_empty, cfg.gen_speculation = math_mode, lanes, walk, spec
        s, inc = pcg32_seed(seed)
        t_cnt.zero_(); t_ns.zero_(); t_ri.zero_()
        rc = lib.ngp_nerf_generate_training_samples(stream, n_rays, 0, n_rays, s, inc, C.byref(cfg), t_views.data_ptr(), n_views, t_bf.data_ptr(), max_samples,
                                                    t_cnt.data_ptr(), t_ri.data_ptr(), t_rays.data_ptr(), t_ns.data_ptr(), t_co.data_ptr())
        assert rc == 0, lib.ngp_last_error()
        torch.cuda.synchronize()
        cnt = t_cnt.cpu().numpy().view(np.uint32)
        k = int(cnt[0])
        ri = t_ri.cpu().numpy().view(np.uint32)[:k]
        ns = t_ns.cpu().numpy().view(np.uint32)[:k]
        co = t_co[:min(int(cnt[1]), max_samples)].cpu().numpy()
        rays = t_rays.cpu().numpy()[:k]
        return {int(r): (int(ns[j, 0]), rays[j].tobytes(), co[ns[j, 1]:ns[j, 1] + ns[j, 0]].copy()) for j, r in enumerate(ri)}, int(cnt[1])

    out = {"steps": args.steps, "rays": n_rays, "occupancy_per_cascade": occupancy, "cases": []}
    variants = [(16, 0, 0), (32, 0, 0), (8, 0, 0), (4, 0, 0), (2, 0, 0), (16, 1, 1), (16, 64, 16), (1, 0, 0)]
    for math_mode in (1, 0):
        for seed in range(1, args.seeds + 1):
            base, n_base = run(seed, math_mode, 1)
            for lanes, walk, spec in variants:
                got, n_got = run(seed, math_mode, lanes, walk, spec)
                rec = {"math_mode": math_mode, "seed": seed, "lanes": lanes, "walk": walk, "speculation": spec,
                       "n_samples": n_got, "n_samples_base": n_base, "rays_base": len(base), "rays_got": len(got)}
                diff_set = sorted(set(base) ^ set(got))
                diff_count, diff_coord, worst = [], 0, 0.0
                for r in set(base) & set(got):
                    b, g = base[r], got[r]
                    if b[0] != g[0]:
                        diff_count.append((r, b[0], g[0]))
                    elif b[1] != g[1] or b[2].tobytes() != g[2].tobytes():
                        diff_coord += 1
                        worst = max(worst, float(np.abs(b[2] - g[2]).max()))
                rec.update(rays_only_in_one=len(diff_set), rays_count_differs=len(diff_count), rays_coords_differ=diff_coord, worst_coord_diff=worst,
                           examples=[list(map(int, x)) for x in diff_count[:8]])
                out["cases"].append(rec)
    out["all_identical"] = all(c["rays_only_in_one"] == 0 and c["rays_count_differs"] == 0 and c["rays_coords_differ"] == 0 for c in out["cases"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
