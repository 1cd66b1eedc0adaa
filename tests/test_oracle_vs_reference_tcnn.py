"""CPU: pin oracle/net_oracle.py against outputs of the REFERENCE network code run on a B200 — the reference's own
ngp::NerfNetwork<__half> (nerf_network.h) on its own tiny-cuda-nn kernels (kernel_grid, kernel_mlp_fused,
kernel_mlp_fused_backward, CUTLASS split-K weight gradients, kernel_grid_backward, adam_step, ema_step), compiled from
/root/reference by oracle/ref/Makefile and driven by oracle/ref/ref_tcnn_harness.cu.  Vectors: tests/golden/ref_tcnn_*.bin.gz
(generator: tests/golden/make_ref_tcnn_goldens.sh).

Tolerances: the reference accumulates in fp16 (wmma fp16 accumulators, fp16 atomics, fp16 split-K outputs) and uses fast-math;
the oracle accumulates in fp32/fp64.  The reference's own test-suite accepts 1e-2 relative between its two code paths
(tiny-cuda-nn/tests/test_common.h:177); the same bound is used here."""
import gzip
import struct
from pathlib import Path

import numpy as np
import pytest

from oracle import net_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
FILES = sorted(GOLD.glob("ref_tcnn_*.bin.gz"))


def load(path):
    d = gzip.open(path, "rb").read()
    magic, L, F, log2_T = struct.unpack_from("<4I", d, 0)
    assert magic == 0x5250474E
    (pls,) = struct.unpack_from("<f", d, 16)
    n_params, n = struct.unpack_from("<2I", d, 20)
    p = 28

    def take(dtype, count):
        nonlocal p
        a = np.frombuffer(d, dtype=dtype, count=count, offset=p).copy()
        p += a.nbytes
        return a

    g = dict(L=L, F=F, log2_T=log2_T, pls=pls, n_params=n_params, n=n)
    g["params"] = take(np.float16, n_params)
    g["coords"] = take(np.float32, n * 7).reshape(n, 7)
    g["inf"] = take(np.float16, n * 16).reshape(n, 16)
    g["fwd"] = take(np.float16, n * 16).reshape(n, 16)
    g["dl"] = take(np.float16, n * 16).reshape(n, 16)
    g["grads"] = take(np.float16, n_params)
    g["steps"] = []
    while p < len(d):
        g["steps"].append(dict(w32=take(np.float32, n_params), w16=take(np.float16, n_params), ema=take(np.float16, n_params), grads=take(np.float16, n_params)))
    assert p == len(d)
    return g


pytestmark = pytest.mark.skipif(not FILES, reason="reference GPU goldens not generated yet (tests/golden/make_ref_tcnn_goldens.sh)")


@pytest.fixture(scope="module", params=FILES, ids=lambda p: p.name)
def gold(request):
    g = load(request.param)
    og = O.grid_layout(g["L"], g["F"], g["log2_T"], 16, g["pls"])
    L = O.NerfLayout(og, 1, 2)
    assert L.n_params == g["n_params"], "oracle layout disagrees with the reference's n_params()"
    return g, L


def test_forward_matches_reference(gold):
    g, L = gold
    want = np.concatenate([g["inf"][:, 0:3], g["inf"][:, 3:4]], axis=1).astype(np.float32)
    got = O.nerf_forward(L, g["params"], g["coords"]).astype(np.float32)
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    print("forward: max abs err", err, "scale", scale)
    assert err <= 1e-2 * max(scale, 1.0)
    # the training-mode forward of the reference agrees with its inference pass
    assert np.abs(g["fwd"][:, :4].astype(np.float32) - g["inf"][:, :4].astype(np.float32)).max() <= 1e-2 * max(scale, 1.0)


def test_backward_matches_reference(gold):
    g, L = gold
    dl4 = g["dl"][:, :4]
    got = O.nerf_backward(L, g["params"], g["coords"], dl4)
    want = g["grads"].astype(np.float64)
    o = 0
    for (r, c) in L.density_shapes + L.rgb_shapes:
        a, b = got[o:o + r * c], want[o:o + r * c]
        scale = np.abs(b).max()
        err = np.abs(a - b).max() / scale
        mean_err = np.abs(a - b).mean() / np.abs(b).mean()
        print(f"layer {r}x{c}: max err / max {err:.3e}, mean err / mean {mean_err:.3e}")
        # the reference produces these with fp16 split-K GEMMs (fp16 partial sums, cutlass_matmul.h:479-531): its own test accepts a
        # MEAN relative error of 1.2e-2 between its own two code paths (tests/test_common.h:213-221); against exact fp32 sums the fp16-accumulated reference is 1-2 % off on average
        assert err < 6e-2 and mean_err < 3e-2
        o += r * c
    gg, wg = got[o:], want[o:]
    touched_w = np.abs(wg) > 0
    touched_g = np.abs(gg) > 0
    # same set of touched hash entries, up to fp16 underflow of individual contributions (tiny magnitudes only)
    scale = np.abs(wg).max()
    only_ref = touched_w & ~touched_g
    assert only_ref.sum() <= 1e-3 * touched_w.sum() and (np.abs(wg[only_ref]) <= 1e-3 * scale).all()
    # the reference sums up to thousands of fp16 contributions per coarse-level entry with fp16 atomics (grid.h:252-255):
    # individual heavily-hit entries are off by several percent, the bulk agrees
    mean_err = np.abs(gg - wg)[touched_w].mean() / np.abs(wg)[touched_w].mean()
    print("grid grads: max err / max", np.abs(gg - wg).max() / scale, "mean err / mean", mean_err)
    assert np.abs(gg - wg).max() <= 0.2 * scale and mean_err <= 2e-2
    assert abs(gg.sum() - wg.sum()) <= 2e-2 * np.abs(wg).sum()


def test_optimizer_matches_reference(gold):
    g, L = gold
    n = g["n_params"]
    w32 = g["params"].astype(np.float32)
    w16 = g["params"].copy()
    ema = g["params"].copy()
    m1, m2, steps = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
    grads_seq = [g["grads"]] + [s["grads"] for s in g["steps"][1:]]
    # steps[i]["grads"] holds the gradient that was consumed by optimizer step i (dumped after the step, not zeroed by the reference)
    for i, st in enumerate(g["steps"]):
        gr = (g["grads"] if i == 0 else st["grads"]).copy()
        O.adam_ema_step(L.n_mlp_params, w32, w16, ema, gr, m1, m2, steps, lr=1e-2, beta1=0.9, beta2=0.99, eps=1e-15, l2_reg=1e-6, loss_scale=128.0,
                        ema_decay=0.95, step=i + 1)
        d32 = np.abs(w32 - st["w32"]).max() / (np.abs(st["w32"]).max())
        d16 = (w16.view(np.uint16) != st["w16"].view(np.uint16)).mean()
        dema = np.abs(ema.astype(np.float32) - st["ema"].astype(np.float32)).max()
        print(f"step {i + 1}: w32 rel {d32:.2e}, w16 mismatching halves {d16:.2e}, ema abs {dema:.2e}")
        assert d32 < 1e-5 and d16 < 5e-3 and dema < 2e-3  # halves flip where the fp32 value sits on a rounding boundary
        # continue from the reference's state so that rounding differences do not accumulate across steps
        w32, w16, ema = st["w32"].copy(), st["w16"].copy(), st["ema"].copy()
