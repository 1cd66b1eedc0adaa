#!/usr/bin/env python
"""The MLP phase of the fused training kernel on its own ("report tensor-pipe % for the MLP phase separately"):
k_nerf_train<2, 256, MLP_ONLY> = the 15 tcgen05 MMA groups of every 128-sample tile (5 forward, 5 data-gradient, 5 weight-gradient) and
their TMEM epilogues, without the hash-grid gather / scatter.  Prints the CUDA-event time next to the full kernel's; run under
`ncu --set full -k regex:k_nerf_train` for sm__pipe_tensor_cycles_active of both."""
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
    n = 1 << 18
    d, L = util.make_desc(n_levels=16, F=2, log2_T=19, aabb_scale=4)
    params = torch.from_numpy(util.random_params(L, seed=0, trained_like=True).astype(np.float16)).cuda()
    coords = torch.rand(n, 7, device="cuda")
    dl = (torch.randn(n, 4, device="cuda") * 0.1).half()
    grads = torch.zeros(d.n_params, dtype=torch.float16, device="cuda")
    scratch = torch.zeros(d.n_mlp_params, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def mlp_only():
        assert lib.ngp_profile_mlp_phase(C.byref(d), st, n, coords.data_ptr(), params.data_ptr(), dl.data_ptr(), grads.data_ptr(), scratch.data_ptr()) == 0, lib.ngp_last_error()

    def full():
        assert lib.ngp_nerf_forward_backward(C.byref(d), st, n, coords.data_ptr(), params.data_ptr(), dl.data_ptr(), grads.data_ptr(), None) == 0, lib.ngp_last_error()

    t_mlp, t_full = timed(mlp_only), timed(full)
    flops = 61440.0 * n     # MLP forward + backward per sample
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    print(json.dumps({"n": n, "mlp_phase_ms": t_mlp, "full_kernel_ms": t_full, "mlp_phase_tflops": flops / t_mlp / 1e9, "tiles": n // 128, "us_per_tile_per_cta": t_mlp * 1e3 / (n / 128 / (2 * sms)),
                      "note": "uniform random positions (no spatial coherence): the full-kernel time is an upper bound of what a training batch sees"}))


if __name__ == "__main__":
    main()
