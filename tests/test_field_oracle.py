"""CPU checks of the image / SDF primitive path: the library's host-side descriptors and initialisation against the
independent oracle, and internal consistency of the oracle's losses and samplers (no GPU)."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import field_oracle as FO
from oracle import net_oracle as O


@pytest.mark.parametrize("kw", [
    dict(n_pos_dims=2, log2_T=24, per_level_scale=1.3819128274917603),   # configs/image/base.json on a 2048-wide image
    dict(n_pos_dims=2, log2_T=15, per_level_scale=2.0),
    dict(n_pos_dims=3, log2_T=19, per_level_scale=1.3819128274917603),   # configs/sdf/base.json
    dict(n_pos_dims=3, log2_T=19, n_levels=8, F=4, per_level_scale=2.0),
])
def test_library_descriptor_equals_oracle_layout_bit_for_bit(kw):
    d, L = util.make_field_desc(**kw)
    g = d.grid
    assert list(g.offsets[: g.n_levels + 1]) == L.grid.offsets
    assert list(g.resolutions[: g.n_levels]) == L.grid.resolutions
    assert np.array(list(g.scales[: g.n_levels]), dtype=np.float32).tobytes() == np.array(L.grid.scales, dtype=np.float32).tobytes()
    assert d.n_params == L.n_params and d.n_mlp_params == L.n_mlp_params and d.grid_offset == L.n_mlp_params


def test_reference_known_answer_for_a_2d_grid():
    """hand-checked against grid.h:699-730: base 16, scale 2 -> resolutions 16, 32 (ceil(31)+1), 64; 2-D level sizes res^2 until capped at 2^12"""
    g = O.grid_layout(3, 2, 12, 16, 2.0, n_pos_dims=2)
    assert g.resolutions == [16, 32, 64]
    assert g.offsets == [0, 256, 256 + 1024, 256 + 1024 + 4096]


def test_param_init_equals_oracle_bit_for_bit():
    lib = util.pkg().load_library()
    d, L = util.make_field_desc(n_pos_dims=2, log2_T=12, per_level_scale=1.5)
    p = np.zeros(d.n_params, dtype=np.float32)
    assert lib.ngp_field_init_params_host(C.byref(d), 1337, p.ctypes.data) == 0
    want = FO.field_init_params(L, 1337)
    assert p.tobytes() == want.tobytes()
    s0 = np.float32(np.sqrt(np.float32(6.0) / np.float32(96)))
    assert np.abs(p[:2048]).max() <= s0 and np.abs(p[L.n_mlp_params:]).max() <= 1e-4


@pytest.mark.parametrize("loss", [FO.LOSS_L2, FO.LOSS_L1, FO.LOSS_MAPE])
def test_loss_gradient_is_the_derivative_of_the_loss(loss):
    """(SMAPE and RelativeL2 treat their prediction-dependent normaliser as a constant in the reference, smape.h / relative_l2.h,
    so their gradient is deliberately not the derivative; they are covered by the bit-exact GPU-vs-oracle test.)"""
    rng = np.random.default_rng(loss)
    n, n_out = 64, 3
    pred = np.zeros((n, 16), dtype=np.float16)
    pred[:, :n_out] = rng.uniform(0.2, 1.0, size=(n, n_out))
    tgt = rng.uniform(0.2, 1.0, size=(n, n_out)).astype(np.float32) + 0.05
    values, g = FO.loss_evaluate(loss, pred, tgt, 128.0, n_out)
    assert (g[:, n_out:] == 0).all()
    eps = 2.0 ** -6
    p2, p1 = pred.copy(), pred.copy()
    p2[:, 0] = (pred[:, 0].astype(np.float32) + eps).astype(np.float16)
    p1[:, 0] = (pred[:, 0].astype(np.float32) - eps).astype(np.float16)
    step = (p2[:, 0].astype(np.float32) - p1[:, 0].astype(np.float32))
    v2, _ = FO.loss_evaluate(loss, p2, tgt, 128.0, n_out)
    v1, _ = FO.loss_evaluate(loss, p1, tgt, 128.0, n_out)
    num = (v2[:, 0] - v1[:, 0]) / step * 128.0
    ana = g[:, 0].astype(np.float32)
    smooth = np.abs(pred[:, 0].astype(np.float32) - tgt[:, 0]) > 2 * eps   # away from the kink of the absolute-value losses
    assert np.allclose(num[smooth], ana[smooth], rtol=0.05, atol=2e-3)


def test_tcnn_random_fill_pattern_and_stratification():
    rng = O.Pcg32(7)
    state0 = rng.state
    a = FO.tcnn_random_uniform(rng, 1000)
    # element i + n_threads*j is draw 4i + j of the stream; n_threads = 256 for 1000 elements
    r = O.Pcg32(7)
    r.state = state0
    seq = [np.float32(r.next_float()) for _ in range(1024)]
    assert a[0] == seq[0] and a[256] == seq[1] and a[512] == seq[2] and a[1] == seq[4] and a[257] == seq[5]
    r2 = O.Pcg32(7)
    r2.state = state0
    r2.advance(1000)
    assert rng.state == r2.state
    pos = FO.stratify2(np.full((256, 2), 0.5, dtype=np.float32), 8)
    cell = np.floor(pos * 16).astype(int)
    assert len({(x, y) for x, y in cell}) == 256    # one sample per cell of the 16 x 16 stratification


def test_eval_image_snap_and_bilinear():
    img = util.test_image(32, 16)
    pos = np.array([[0.0, 0.0], [0.999, 0.999], [0.5, 0.5], [(3 + 0.5) / 32, (5 + 0.5) / 16]], dtype=np.float32)
    p2, val = FO.eval_image_and_snap(img, pos, snap=True, linear_colors=True)
    assert np.allclose(val[3], img[5, 3, :3]) and np.allclose(p2[3], pos[3])
    assert np.allclose(p2[0], [0.5 / 32, 0.5 / 16])
    _, val_b = FO.eval_image_and_snap(img, pos, snap=False, linear_colors=True)
    assert np.allclose(val_b[3], img[5, 3, :3], atol=1e-6)         # pixel centres reproduce the pixel
    _, val_s = FO.eval_image_and_snap(img, pos, snap=False, linear_colors=False)
    assert (val_s[3] >= val_b[3] - 1e-6).all()                      # sRGB encoding brightens values in (0, 1)


def test_shuffle_is_the_reference_permutation():
    a = np.arange(30, dtype=np.float32).reshape(10, 3)
    s = FO.shuffle(a, 3, 4)
    for i in range(10):
        j = ((i + 4) * 1434869437 + 2097192037) % 10
        assert (s[i] == a[j]).all()


def test_module_handle_mirrors_tcnn_cpp_module():
    """tcnn::cpp::Module surface (cpp_api.h:92-125): create from the two JSON configs, sizes, pcg32{seed} initialisation"""
    lib = util.pkg().load_library()
    P = util.pkg()
    enc = b'{"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 12, "base_resolution": 16, "per_level_scale": 1.5}'
    net = b'{"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2}'
    m = lib.ngp_module_create_network_with_input_encoding(3, 1, enc, net)
    assert m, lib.ngp_last_error()
    d = P.FieldDesc()
    assert lib.ngp_module_get_desc(m, C.byref(d)) == 0
    og = O.grid_layout(16, 2, 12, 16, 1.5, n_pos_dims=3)
    L = FO.FieldLayout(og, 2, 1)
    assert lib.ngp_module_n_params(m) == L.n_params == d.n_params
    assert lib.ngp_module_n_input_dims(m) == 3 and lib.ngp_module_n_output_dims(m) == 16     # padded width, like Module::n_output_dims
    p = np.zeros(L.n_params, dtype=np.float32)
    assert lib.ngp_module_initialize_params(m, 42, p.ctypes.data, 1.0) == 0
    # same generator as the oracle's initialiser, seeded directly with pcg32{seed} (cpp_api.cu:141-144) instead of through seed_seq
    rng = O.Pcg32(42)
    want = np.empty(L.n_params, dtype=np.float32)
    o = 0
    for (r, c) in L.shapes:
        scale = np.float32(np.sqrt(np.float32(6.0) / np.float32(r + c)))
        for i in range(r * c):
            want[o + i] = np.float32(rng.next_float()) * np.float32(2.0) * scale - scale
        o += r * c
    want[o:] = FO.tcnn_random_uniform(rng, L.grid.n_params, np.float32(-1e-4), np.float32(1e-4), advance=False)
    assert p.tobytes() == want.tobytes()
    lib.ngp_module_free(m)
    # unsupported configurations fail loudly with the reason
    bad = lib.ngp_module_create_network_with_input_encoding(3, 1, b'{"otype": "Frequency"}', net)
    assert not bad and b"HashGrid" in lib.ngp_last_error()
    bad = lib.ngp_module_create_network_with_input_encoding(3, 1, enc, b'{"otype": "FullyFusedMLP", "n_neurons": 128}')
    assert not bad and b"64" in lib.ngp_last_error()
