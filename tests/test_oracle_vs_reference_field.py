"""CPU: pin oracle/field_oracle.py against outputs of the REFERENCE's own NetworkWithInputEncoding<__half> + Loss objects
(tiny-cuda-nn network_with_input_encoding.h, losses/l2.h, losses/mape.h, kernel_grid over 2-D / 3-D positions, kernel_mlp_fused,
kernel_mlp_fused_backward, CUTLASS split-K weight gradients, kernel_grid_backward) run on a B200 — compiled from
/root/reference by oracle/ref/Makefile, driven by oracle/ref/ref_tcnn_harness.cu (run_field).  Vectors:
tests/golden/ref_field_*.bin.gz (generator: tests/golden/make_ref_tcnn_goldens.sh).

Tolerances as in tests/test_oracle_vs_reference_tcnn.py: the reference accumulates in fp16 and is built with --use_fast_math;
its own test-suite accepts 1e-2 relative between its code paths (tiny-cuda-nn/tests/test_common.h:177)."""
import gzip
import struct
from pathlib import Path

import numpy as np
import pytest

from oracle import field_oracle as FO
from oracle import net_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
FILES = sorted(GOLD.glob("ref_field_*.bin.gz"))
pytestmark = pytest.mark.skipif(not FILES, reason="reference GPU goldens not generated yet (tests/golden/make_ref_tcnn_goldens.sh)")


def load(path):
    d = gzip.open(path, "rb").read()
    magic, D, L, F, log2_T = struct.unpack_from("<5I", d, 0)
    assert magic == 0x4650474E
    (pls,) = struct.unpack_from("<f", d, 20)
    n_hidden, n_out, loss_type, n_params, n = struct.unpack_from("<5I", d, 24)
    p = 44

    def take(dtype, count):
        nonlocal p
        a = np.frombuffer(d, dtype=dtype, count=count, offset=p).copy()
        p += a.nbytes
        return a

    g = dict(D=D, L=L, F=F, log2_T=log2_T, pls=pls, n_hidden=n_hidden, n_out=n_out, loss=loss_type, n_params=n_params, n=n)
    g["params"] = take(np.float16, n_params)
    g["pos"] = take(np.float32, n * D).reshape(n, D)
    g["tgt"] = take(np.float32, n * n_out).reshape(n, n_out)
    g["inf"] = take(np.float16, n * 16).reshape(n, 16)
    g["fwd"] = take(np.float16, n * 16).reshape(n, 16)
    g["values"] = take(np.float32, n * 16).reshape(n, 16)
    g["dl"] = take(np.float16, n * 16).reshape(n, 16)
    g["grads"] = take(np.float16, n_params)
    assert p == len(d)
    return g


@pytest.fixture(scope="module", params=FILES, ids=lambda p: p.name)
def gold(request):
    g = load(request.param)
    og = O.grid_layout(g["L"], g["F"], g["log2_T"], 16, g["pls"], n_pos_dims=g["D"])
    L = FO.FieldLayout(og, g["n_hidden"], g["n_out"])
    assert L.n_params == g["n_params"], "oracle layout (MLP | grid) disagrees with the reference's n_params()"
    return g, L


def test_forward_matches_reference(gold):
    g, L = gold
    want = g["inf"].astype(np.float32)
    got = FO.field_forward(L, g["params"], g["pos"]).astype(np.float32)
    k = g["n_out"]
    scale = max(np.abs(want[:, :k]).max(), 1.0)
    err = np.abs(got[:, :k] - want[:, :k]).max()
    print("forward: max abs err", err, "scale", scale)
    assert err <= 1e-2 * scale
    assert np.abs(g["fwd"][:, :k].astype(np.float32) - want[:, :k]).max() <= 1e-2 * scale


def test_loss_matches_reference(gold):
    """the loss arithmetic on the REFERENCE's own fp16 predictions: same formula, IEEE here vs fast-math there"""
    g, L = gold
    k = g["n_out"]
    values, dl = FO.loss_evaluate(g["loss"], g["fwd"], g["tgt"], 128.0, k)
    assert (g["values"][:, k:] == 0).all() and (g["dl"][:, k:] == 0).all()      # padding carries no loss / gradient
    assert np.allclose(values, g["values"][:, :k], rtol=2e-5, atol=1e-9)
    a, b = dl[:, :k].astype(np.float32), g["dl"][:, :k].astype(np.float32)
    assert np.abs(a - b).max() <= 2.0 ** -10 * np.abs(b).max()                   # one fp16 ulp of the largest gradient
    assert (a.view(np.uint32) != b.view(np.uint32)).mean() < 0.02
    # Trainer::loss = sum of the per-element terms
    assert abs(values.sum() - g["values"].sum()) <= 1e-5 * abs(g["values"].sum())


def test_backward_matches_reference(gold):
    g, L = gold
    got = FO.field_backward(L, g["params"], g["pos"], g["dl"])
    want = g["grads"].astype(np.float64)
    o = 0
    for (r, c) in L.shapes:
        a, b = got[o:o + r * c], want[o:o + r * c]
        if r == 16:   # only the first n_out rows of the padded output layer receive gradient
            assert (np.abs(b.reshape(r, c)[g["n_out"]:]) == 0).all()
        scale = np.abs(b).max()
        err = np.abs(a - b).max() / scale
        mean_err = np.abs(a - b).mean() / np.abs(b).mean()
        print(f"layer {r}x{c}: max err / max {err:.3e}, mean err / mean {mean_err:.3e}")
        assert err < 6e-2 and mean_err < 3e-2
        o += r * c
    gg, wg = got[o:], want[o:]
    touched_w = np.abs(wg) > 0
    touched_g = np.abs(gg) > 0
    scale = np.abs(wg).max()
    only_ref = touched_w & ~touched_g
    assert only_ref.sum() <= 1e-3 * touched_w.sum() and (np.abs(wg[only_ref]) <= 1e-3 * scale).all()
    mean_err = np.abs(gg - wg)[touched_w].mean() / np.abs(wg)[touched_w].mean()
    print("grid grads: max err / max", np.abs(gg - wg).max() / scale, "mean err / mean", mean_err)
    assert np.abs(gg - wg).max() <= 0.2 * scale and mean_err <= 2e-2
    assert abs(gg.sum() - wg.sum()) <= 2e-2 * np.abs(wg).sum()
