"""CPU: pin the oracle against every known answer the reference's own tests hold for this path
(tiny-cuda-nn/tests/test_grid.cu:57-71) and against closed-form properties."""
import numpy as np

from oracle import net_oracle as O


def test_grid_layout_known_answers():
    g = O.grid_layout(20, 2, 16, 32, 1.5)
    assert g.offsets[0] == 0
    assert g.offsets[1] - g.offsets[0] == 32 ** 3
    assert g.offsets[2] - g.offsets[1] == 65536
    assert g.offsets[3] - g.offsets[2] == 65536
    assert g.n_params == 2555904


def test_survey_table_sizes():
    # encoding params of the three NeRF configurations
    for (L, F, aabb, want) in [(8, 4, 4, 12855296), (16, 2, 4, 13074912), (16, 2, 1, 12196240)]:
        pls = O.per_level_scale_for(aabb, 16, L)
        g = O.grid_layout(L, F, 19, 16, pls)
        assert g.n_params == want, (L, F, aabb, g.n_params)
        assert g.resolutions[0] == 16 and g.resolutions[-1] in (2048 * aabb, 2048 * aabb + 1)


def test_encode_interpolates_constant_and_linear_fields():
    g = O.grid_layout(4, 2, 19, 16, 1.5)
    grid = np.full(g.n_params, 0.25, dtype=np.float16)
    pos = np.random.default_rng(0).uniform(0, 1, size=(257, 3)).astype(np.float32)
    enc = O.grid_encode(g, grid, pos).astype(np.float32)
    assert np.allclose(enc, 0.25, atol=2e-3)  # partition of unity (up to fp16 rounding of the 8 fmas)


def test_backward_is_adjoint_of_forward():
    g = O.grid_layout(16, 2, 14, 16, 1.4)
    rng = np.random.default_rng(1)
    grid = rng.normal(0, 0.1, size=g.n_params).astype(np.float16)
    pos = rng.uniform(0, 1, size=(64, 3)).astype(np.float32)
    dy = rng.normal(0, 1, size=(64, 32)).astype(np.float16)
    enc = O.grid_encode(g, grid, pos).astype(np.float64)
    ggrid = O.grid_backward(g, pos, dy)
    lhs = float((enc * dy.astype(np.float64)).sum())
    rhs = float((ggrid * grid.astype(np.float64)).sum())
    assert abs(lhs - rhs) <= 2e-2 * (abs(lhs) + 1.0)


def test_pcg32_matches_published_reference_sequence():
    # PCG32 demo values for seed(42, 54) from the published pcg32-demo: 0xa15c02b7 0x7b47f409 0xba1d3330 ...
    r = O.Pcg32(42, 54)
    assert [r.next_uint() for _ in range(3)] == [0xA15C02B7, 0x7B47F409, 0xBA1D3330]
    a = O.Pcg32(1337)
    b = O.Pcg32(1337)
    for _ in range(1000):
        a.next_uint()
    b.advance(1000)
    assert a.state == b.state


def test_mlp_backward_matches_finite_differences():
    rng = np.random.default_rng(3)
    ws = [rng.normal(0, 0.3, size=s).astype(np.float16) for s in [(64, 32), (16, 64)]]
    x = rng.normal(0, 1, size=(8, 32)).astype(np.float16)
    out, acts = O.mlp_forward(ws, x, keep=True)
    dy = rng.normal(0, 1, size=(8, 16)).astype(np.float16)
    grads, dx = O.mlp_backward(ws, acts, dy)
    # analytic in float64 without fp16 rounding
    h = np.maximum(x.astype(np.float64) @ ws[0].astype(np.float64).T, 0)
    g1 = dy.astype(np.float64).T @ h
    assert np.allclose(grads[1], g1, rtol=2e-2, atol=2e-2)
