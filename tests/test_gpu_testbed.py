"""GPU: the Testbed surface end to end — train / render / snapshot / counters — on a procedural scene."""
import importlib

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
S = importlib.import_module("instant-ngp_b200.synthetic")


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10.0 * np.log10(1.0 / max(mse, 1e-12))


def srgb_to_linear(x):
    return np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


@pytest.fixture(scope="module")
def trained():
    P = util.pkg()
    tb = P.Testbed()
    imgs, cams, focal = S.make_dataset(n_images=24, width=160, height=160)
    S.load_into_testbed(tb, imgs, cams, focal, aabb_scale=1)
    tb.reload_network_from_json(S.base_config(16, 2, 17))
    losses = []
    hist = []
    for i in range(400):
        tb.train(1 << 16)
        if tb.training_step % 16 == 1:
            losses.append(tb.loss)
        hist.append(tb.counters())
    return tb, imgs, cams, focal, losses, hist


def test_training_converges_and_counters_behave(trained):
    tb, imgs, cams, focal, losses, hist = trained
    assert tb.training_step == 400
    assert all(np.isfinite(losses)) and losses[-1] < 0.35 * losses[0], losses
    last = hist[-1]
    # the rays_per_batch controller steers the compacted batch towards the target (testbed_nerf.cu:2698-2699)
    assert 0.6 * (1 << 16) <= last["measured_batch_size"] <= 1.10 * (1 << 16), last
    assert last["rays_per_batch"] % 256 == 0 and last["rays_per_batch"] <= 1 << 18
    assert last["measured_batch_size_before_compaction"] >= last["measured_batch_size"]
    grid, bits = tb.get_density_grid()
    occ = np.unpackbits(bits[: 128 ** 3 // 8]).mean()
    assert 0.005 < occ < 0.5, occ  # the ball fills ~9 % of the unit cube
    assert np.isfinite(grid).all()


def test_render_matches_training_view(trained):
    tb, imgs, cams, focal, _, _ = trained
    h, w = imgs.shape[1:3]
    got, depth = tb.render(w, h, cams[5], focal, return_depth=True)
    want = imgs[5]
    # ground truth is linear premultiplied; render output is linear premultiplied as well (shade_kernel_nerf)
    p = psnr(np.clip(got[..., :3], 0, 1), want[..., :3])
    cov_err = np.abs((got[..., 3] > 0.5).mean() - (want[..., 3] > 0.5).mean())
    print("PSNR", p, "coverage err", cov_err, "steps", tb.last_render_steps)
    assert p > 20.0 and cov_err < 0.03
    inside = want[..., 3] > 0.5
    assert np.isfinite(depth).all() and 0.5 < np.median(depth[inside & (got[..., 3] > 0.5)]) < 2.0
    # tile sharding: rows rendered separately equal the full frame bit for bit
    top = tb.render(w, h, cams[5], focal, rows=(0, h // 2))
    bot = tb.render(w, h, cams[5], focal, rows=(h // 2, h))
    assert np.array_equal(top[: h // 2], got[: h // 2]) and np.array_equal(bot[h // 2:], got[h // 2:])


def test_snapshot_round_trip(trained, tmp_path):
    tb, imgs, cams, focal, _, _ = trained
    P = util.pkg()
    path = tmp_path / "model.ngpb"
    tb.save_snapshot(str(path))
    a = tb.render(96, 96, cams[2], focal * 96 / imgs.shape[2])
    tb2 = P.Testbed()
    tb2.load_snapshot(str(path))
    b = tb2.render(96, 96, cams[2], focal * 96 / imgs.shape[2])
    assert np.array_equal(a, b)
    assert tb2.training_step == tb.training_step and tb2.n_params == tb.n_params
    assert np.array_equal(tb2.get_params(inference=True), tb.get_params(inference=True))


def test_errors_are_reported(trained):
    P = util.pkg()
    tb = P.Testbed()
    with pytest.raises(P.NgpError):
        tb.train(1 << 16)  # no network
    tb.create_empty_nerf_dataset(2, aabb_scale=1)
    with pytest.raises(P.NgpError):
        tb.reload_network_from_json({"encoding": {"otype": "Frequency"}})
    with pytest.raises(P.NgpError):
        tb.create_empty_nerf_dataset(2, aabb_scale=3)
    tb.reload_network_from_json(S.base_config(8, 4, 15))
    with pytest.raises(P.NgpError):
        tb.train(1 << 16)  # n_images_for_training == 0
    with pytest.raises(P.NgpError):
        tb.train(1000)     # not a multiple of 256
