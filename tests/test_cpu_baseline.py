"""CPU: the multi-threaded fp32 network port that bench.py times as `cpu_baseline` (oracle/ngp_net_cpu.c) computes the same function as
the parity oracle (oracle/net_oracle.py, fp16 storage like the reference) — to fp16 rounding, which is all a baseline has to promise."""
import numpy as np

import util
from oracle import net_cpu
from oracle import net_oracle as O


def test_cpu_baseline_network_matches_the_oracle():
    d, L = util.make_desc(n_levels=16, F=2, log2_T=14, aabb_scale=4)
    params = util.random_params(L, seed=3, trained_like=True).astype(np.float16)
    coords = util.random_coords(384, seed=5)
    dl = (np.random.default_rng(7).normal(0, 1, size=(384, 4)) * 0.25).astype(np.float16)
    net = net_cpu.NetCpu(L, params)
    want = O.nerf_forward(L, params, coords).astype(np.float32)
    got = net.forward(coords)
    assert np.abs(got - want).max() < 3e-2 * max(1.0, np.abs(want).max())
    out2, grads = net.forward_backward(coords, dl)
    assert np.array_equal(out2, got)
    g_want = O.nerf_backward(L, params, coords, dl)
    n_mlp = L.n_mlp_params
    # MLP weights: 3 %; hash grid: the oracle rounds dL/dencoding and every weight x gradient product to fp16 like kernel_grid_backward
    for a, b, tol in ((grads[:n_mlp], g_want[:n_mlp], 3e-2), (grads[n_mlp:], g_want[n_mlp:], 8e-2)):
        assert np.abs(a - b).max() <= tol * np.abs(b).max()
    assert net_cpu.threads() >= 1
