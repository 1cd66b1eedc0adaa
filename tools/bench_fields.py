#!/usr/bin/env python
"""Encoding + MLP micro-benchmark of the image / SDF primitives (BASELINE configs #1 and #4): forward, fused
forward+loss+backward, and forward+backward+optimizer of NetworkWithInputEncoding at N in {65 536, 262 144, 1 048 576} samples,
positions U[0,1)^D, "trained-like" parameters.  CUDA-event timing on the launching stream, 20 iterations after 5 warm-ups, inputs
rotated over 8 position sets so that no iteration re-reads the previous one's lines.  One JSON line per configuration.

    python tools/bench_fields.py            (needs a B200; not part of bench.py's contract)"""
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch

    import util

    P = importlib.import_module("instant-ngp_b200")
    lib = P.load_library()
    st = torch.cuda.current_stream().cuda_stream
    # configs/image/base.json (T 2^24 on a 2048-wide image) and configs/sdf/base.json
    cases = [("image", dict(n_pos_dims=2, log2_T=24, per_level_scale=1.3195079, n_hidden=2, n_out=3), 0),
             ("sdf", dict(n_pos_dims=3, log2_T=19, per_level_scale=1.3819129, n_hidden=2, n_out=1), 2)]
    for name, kw, loss in cases:
        d, L = util.make_field_desc(**kw)
        n_params = d.n_params
        rng = np.random.default_rng(0)
        p16 = torch.from_numpy(np.clip(rng.normal(0, 0.1, size=n_params), -1, 1).astype(np.float16)).cuda()
        p32 = p16.float()
        ema = p16.clone()
        grads = torch.zeros(n_params, dtype=torch.float16, device="cuda")
        m1, m2 = torch.zeros(n_params, device="cuda"), torch.zeros(n_params, device="cuda")
        steps = torch.zeros(n_params, dtype=torch.int32, device="cuda")
        tmp = torch.zeros(d.n_mlp_params, device="cuda")
        adam = P.AdamCfg(learning_rate=1e-2, beta1=0.9, beta2=0.99, epsilon=1e-15, l2_reg=1e-6, loss_scale=128.0, ema_decay=0.95, ema_step=1,
                         optimize_matrix_params=1, optimize_non_matrix_params=1)
        for n in (1 << 16, 1 << 18, 1 << 20):
            D, n_out = kw["n_pos_dims"], kw["n_out"]
            pos = [torch.rand(n, D, device="cuda") for _ in range(8)]
            tgt = torch.rand(n, n_out, device="cuda")
            out = torch.zeros(n, 16, dtype=torch.float16, device="cuda")

            def timed(fn, iters=20, warm=5):
                for i in range(warm):
                    fn(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(iters):
                    fn(i)
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / iters

            def fwd(i):
                assert lib.ngp_field_inference(C.byref(d), st, n, pos[i % 8].data_ptr(), p16.data_ptr(), out.data_ptr(), 16) == 0

            def fwd_bwd(i):
                assert lib.ngp_field_train_step(C.byref(d), st, n, pos[i % 8].data_ptr(), tgt.data_ptr(), loss, 128.0, None, p16.data_ptr(), grads.data_ptr(), None, None) == 0

            def step(i):
                fwd_bwd(i)
                adam.ema_step = i + 1
                assert lib.ngp_optimizer_step_flat(d.n_mlp_params, n_params, st, C.byref(adam), p32.data_ptr(), p16.data_ptr(), ema.data_ptr(), grads.data_ptr(),
                                                   m1.data_ptr(), m2.data_ptr(), steps.data_ptr()) == 0

            t_f, t_fb, t_s = timed(fwd), timed(fwd_bwd), timed(step)
            corners = 1 << D
            fwd_bytes = 16 * corners * 4 + 4 * D + 32           # table reads + position + padded fp16 output row
            fb_bytes = 2 * 16 * corners * 4 + 16 * corners * 4 + 4 * D + 4 * n_out   # reads + reduction RMW traffic + position + targets
            print(json.dumps({"case": name, "n": n, "n_params": int(n_params), "fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_bwd_opt_ms": t_s,
                              "fwd_msamples_per_s": n / t_f / 1e3, "fwd_bwd_msamples_per_s": n / t_fb / 1e3, "step_msamples_per_s": n / t_s / 1e3,
                              "fwd_algorithmic_gbs": n * fwd_bytes / t_f / 1e6, "fwd_bwd_algorithmic_gbs": n * fb_bytes / t_fb / 1e6}), flush=True)
            grads.zero_()


if __name__ == "__main__":
    main()
