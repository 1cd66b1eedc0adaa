"""GPU: the Testbed surface end to end — train / render / snapshot / counters — on a procedural scene."""
import importlib
from pathlib import Path

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


def test_render_spp_and_srgb_epilogue(trained):
    """Testbed.render(width, height, spp, linear): spp frames accumulated, sRGB output for linear=False (python_api.cu:507-519)"""
    from oracle import march_oracle as M

    tb, imgs, cams, focal, _, _ = trained
    res = 80
    f = focal * res / imgs.shape[2]
    one = tb.render(res, res, cams[4], f)
    four = tb.render(res, res, cams[4], f, spp=4)
    # samples differ only in the jitter of each ray's first step: the average stays close to one sample and is not identical
    assert not np.array_equal(one, four) and psnr(np.clip(one[..., :3], 0, 1), np.clip(four[..., :3], 0, 1)) > 30.0
    tb.background_color = [0.0, 0.0, 0.0, 0.0]   # transparent background: the epilogue then leaves alpha alone (default is opaque black, testbed.h:1031)
    srgb = tb.render(res, res, cams[4], f, spp=1, linear=False)
    want = np.empty_like(one[..., :3])
    import ctypes as C

    src = np.ascontiguousarray(one[..., :3])
    M.lib().orc_linear_to_srgb_n(src.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), src.size)
    diff = np.abs(srgb[..., :3].astype(np.float64) - want.astype(np.float64))
    print("srgb epilogue: max abs diff", diff.max(), "mismatching", int((srgb[..., :3] != want).sum()), "of", want.size)
    big = want > 1e-30   # (sub-denormal dust of empty pixels aside)
    assert np.array_equal(srgb[..., :3][big], want[big]) and diff.max() < 1e-30 and np.array_equal(srgb[..., 3], one[..., 3])
    tb.background_color = [0.0, 0.0, 0.0, 1.0]
    opaque = tb.render(res, res, cams[4], f, spp=1, linear=False)
    assert np.allclose(opaque[..., 3], 1.0, atol=1e-6) and np.array_equal(opaque[..., :3], srgb[..., :3])


def test_reference_shaped_camera_api(trained):
    """render(width, height, spp, linear) from the Testbed's own camera: set_camera_to_training_view / set_nerf_camera_matrix /
    fov / zoom / screen_center (python_api.cu:507-519, 653-654; src/testbed.cu:440, 486-505, 4081-4087, 4649-4657)"""
    import math
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import loader_scenes

    tb, imgs, cams, focal, _, _ = trained
    P = util.pkg()
    h, w = imgs.shape[1:3]
    v = tb.training_view(5)
    assert v["resolution"] == (w, h) and np.allclose(v["xform"], np.asarray(cams[5])[:3, :4], atol=1e-6)
    assert abs(v["focal_length"][0] - focal) < 1e-3 and v["principal_point"] == (0.5, 0.5)
    tb.set_camera_to_training_view(5)
    assert np.array_equal(tb.render(w, h, 1, True), tb.render(w, h, cams[5], focal))
    assert np.array_equal(tb.render(w, h), tb.render(w, h, cams[5], focal))
    # another frame size: the focal length follows the resolution along fov_axis
    half = tb.render(w // 2, h // 2, spp=1, linear=True)
    assert np.array_equal(half, tb.render(w // 2, h // 2, cams[5], focal / 2))
    rgba, depth = tb.render_with_depth(w // 2, h // 2, 1, True)
    assert np.array_equal(rgba, half) and depth.shape == (h // 2, w // 2) and np.isfinite(depth).all()
    # fov in degrees along fov_axis
    tb.fov = 40.0
    assert abs(tb.fov - 40.0) < 1e-4 and tb.fov_axis == 1
    assert np.array_equal(tb.render(64, 64), tb.render(64, 64, cams[5], 0.5 * 64 / math.tan(math.radians(20.0))))
    tb.zoom = 2.0
    assert np.array_equal(tb.render(64, 64), tb.render(64, 64, cams[5], 64 / math.tan(math.radians(20.0))))
    tb.zoom = 1.0
    # a camera in the original NeRF convention goes through the dataset's scale / offset
    tb.set_nerf_camera_matrix(loader_scenes.ngp_to_nerf(cams[3]))
    assert np.allclose(tb.camera_matrix, np.asarray(cams[3])[:3, :4], atol=1e-5)
    # exposure (stops) acts in the sRGB epilogue only
    tb.set_camera_to_training_view(4)
    lin = tb.render(64, 64, 1, True)
    base = tb.render(64, 64, 1, False)
    tb.exposure = 1.0
    assert tb.exposure == 1.0 and np.array_equal(tb.render(64, 64, 1, True), lin)
    brighter = tb.render(64, 64, 1, False)
    tb.exposure = 0.0
    lit = lin[..., :3].max(axis=-1) > 0.05
    assert lit.any() and (brighter[..., :3][lit] >= base[..., :3][lit]).all() and brighter[..., :3][lit].mean() > 1.2 * base[..., :3][lit].mean()
    assert tb.render_mode == P.RenderMode.Shade and not tb.snap_to_pixel_centers and tb.jit_fusion   # m_snap_to_pixel_centers defaults to false (testbed.h)
    with pytest.raises(P.NgpError):
        tb.render_mode = P.RenderMode.Normals    # needs network input gradients: not built
    # Depth / Cost / AO / Positions through the Testbed: alpha agrees with the shaded frame, Cost counts the frame's network evaluations
    tb.render_mode = P.RenderMode.Depth
    dep = tb.render(64, 64, 1, True)
    assert np.allclose(dep[..., 3], lin[..., 3], atol=1e-5) and np.allclose(dep[..., 0], dep[..., 1]) and (dep[..., 0][lit] > 0).all()
    tb.render_mode = P.RenderMode.Cost
    cost = tb.render(64, 64, 1, True)
    assert (cost[..., 3] == 1.0).all() and int(round(float(cost[..., 0].sum()) * 128)) == tb.last_render_steps
    tb.render_mode = P.RenderMode.Shade
    with pytest.raises(P.NgpError):
        tb.render(64, 64, 1, True, 0.0, 1.0)     # camera paths


def test_reset_and_load_file(tmp_path):
    """Testbed.reset (python_api.cu:534) and load_file (python_api.cu:573, src/testbed.cu:353-410)"""
    import json

    P = util.pkg()
    tb = P.Testbed()
    imgs, cams, focal = S.make_dataset(n_images=8, width=64, height=64)
    S.load_into_testbed(tb, imgs, cams, focal, aabb_scale=1)
    cfg_path = tmp_path / "tiny.json"
    cfg_path.write_text(json.dumps(S.base_config(16, 2, 14)))
    tb.load_file(cfg_path)                      # has "encoding" / "network": a network config
    n = tb.n_params
    assert n > 0 and tb.training_step == 0
    w0 = tb.get_params()
    for _ in range(20):
        tb.train(1 << 14)
    grid_before, _ = tb.get_density_grid()
    assert tb.training_step == 20 and not np.array_equal(tb.get_params(), w0) and grid_before.max() > 0
    tb.reset(reset_density_grid=False)
    grid_kept, _ = tb.get_density_grid()
    assert tb.training_step == 0 and np.array_equal(tb.get_params(), w0) and np.array_equal(grid_kept, grid_before)
    tb.reset()
    grid_cleared, _ = tb.get_density_grid()
    assert np.array_equal(tb.get_params(), w0) and grid_cleared.max() == 0
    for _ in range(3):
        tb.train(1 << 14)
    snap = tmp_path / "m.ingp"
    tb.save_snapshot(str(snap))
    tb2 = P.Testbed()
    tb2.load_file(snap)                         # by extension: a snapshot
    assert tb2.training_step == 3 and tb2.n_params == n
    with pytest.raises(P.NgpError, match="does not exist"):
        tb2.load_file(tmp_path / "nothing.ingp")


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


def _read_ingp(path):
    import gzip

    import msgpack

    raw = open(path, "rb").read()
    if str(path).endswith(".ingp"):
        raw = gzip.decompress(raw)
    return msgpack.unpackb(raw, raw=False)


def test_ingp_snapshot_has_the_reference_schema_and_round_trips(trained, tmp_path):
    """.ingp / .msgpack: gzip(msgpack) of {network config, "snapshot": {...}} as Testbed::save_snapshot writes it
    (src/testbed.cu:5288-5355), read here with the independent msgpack / gzip modules"""
    tb, imgs, cams, focal, _, _ = trained
    P = util.pkg()
    path = tmp_path / "model.ingp"
    tb.save_snapshot(str(path), include_optimizer_state=True)
    d = _read_ingp(path)
    n = tb.n_params
    assert {"encoding", "network", "rgb_network", "dir_encoding", "loss", "optimizer", "snapshot"} <= set(d)
    assert d["encoding"]["n_levels"] == 16 and d["encoding"]["log2_hashmap_size"] == 17 and d["optimizer"]["nested"]["nested"]["otype"] == "Adam"
    s = d["snapshot"]
    assert s["version"] == 1 and s["mode"] == "nerf" and s["density_grid_size"] == 128 and s["params_type"] == "__half" and s["n_params"] == n
    assert len(s["params_binary"]) == 2 * n and len(s["density_grid_binary"]) == 2 * 128 ** 3
    assert np.array_equal(np.frombuffer(s["params_binary"], dtype=np.float16), tb.get_params(inference=True))
    grid, _ = tb.get_density_grid()
    assert np.array_equal(np.frombuffer(s["density_grid_binary"], dtype=np.float16), grid.astype(np.float16).reshape(-1))
    assert s["training_step"] == tb.training_step and s["nerf"]["aabb_scale"] == 1
    assert set(s["nerf"]["rgb"]) == {"rays_per_batch", "measured_batch_size", "measured_batch_size_before_compaction"}
    ds = s["nerf"]["dataset"]
    assert ds["n_images"] == 24 and len(ds["xforms"]) == 24 and np.array(ds["xforms"][3]["start"]).shape == (3, 4)
    assert ds["metadata"][0]["resolution"] == [160, 160] and abs(ds["metadata"][0]["focal_length"][0] - focal) < 1e-3
    assert np.allclose(np.array(ds["xforms"][3]["start"]), np.asarray(cams[3])[:3, :4], atol=1e-6)
    o = s["optimizer"]    # Ema { ExponentialDecay { Adam } } nesting of the config
    assert len(o["weights_ema_binary"]) == 2 * n and "learning_rate_factor" in o["nested"]
    adam = o["nested"]["nested"]
    assert adam["current_step"] == tb.training_step and len(adam["first_moments_binary"]) == 4 * n and len(adam["param_steps_binary"]) == 4 * n

    # a fresh Testbed without a dataset takes everything from the file and renders the same picture
    res = 96
    a = tb.render(res, res, cams[2], focal * res / imgs.shape[2])
    tb2 = P.Testbed()
    tb2.load_snapshot(str(path))
    assert tb2.training_step == tb.training_step and tb2.n_params == n
    assert np.array_equal(tb2.get_params(inference=True), tb.get_params(inference=True))
    b = tb2.render(res, res, cams[2], focal * res / imgs.shape[2])
    # the occupancy grid travels as fp16: cells within half an fp16 ulp of the threshold may flip
    assert psnr(np.clip(a[..., :3], 0, 1), np.clip(b[..., :3], 0, 1)) > 45.0
    # training resumes from the stored optimizer state once the images are back
    S.load_into_testbed(tb2, imgs, cams, focal, aabb_scale=1)
    tb2.load_snapshot(str(path))
    for _ in range(17):
        tb2.train(1 << 16)
    assert tb2.training_step == tb.training_step + 17 and np.isfinite(tb2.loss) and tb2.loss < 2.0 * max(tb.loss, 1e-4)

    # .msgpack (no gzip), without optimizer state
    p2 = tmp_path / "model.msgpack"
    tb.save_snapshot(str(p2))
    d2 = _read_ingp(p2)
    assert "optimizer" not in d2["snapshot"] and d2["snapshot"]["params_binary"] == s["params_binary"]


def test_loads_a_snapshot_written_the_way_the_reference_writes_it(trained, tmp_path):
    """a file assembled with the independent msgpack module in the reference's schema, float32 parameters
    (Trainer::deserialize's "float" branch, trainer.h:459-462) and a never-populated density grid"""
    import gzip

    import msgpack

    tb, imgs, cams, focal, _, _ = trained
    P = util.pkg()
    n = tb.n_params
    w = np.random.default_rng(0).normal(0, 0.05, size=n).astype(np.float32)
    cfg = S.base_config(16, 2, 17)
    cfg["snapshot"] = {
        "version": 1, "mode": "nerf", "density_grid_size": 128, "density_grid_binary": b"", "n_params": n, "params_type": "float",
        "params_binary": w.tobytes(), "training_step": 1234, "loss": 0.5, "aabb": {"min": [0, 0, 0], "max": [1, 1, 1]}, "bounding_radius": 1.0,
        "nerf": {"aabb_scale": 1, "rgb": {"rays_per_batch": 4096, "measured_batch_size": 0, "measured_batch_size_before_compaction": 0}},
    }
    path = tmp_path / "ref_style.ingp"
    path.write_bytes(gzip.compress(msgpack.packb(cfg, use_single_float=False)))
    tb2 = P.Testbed()
    tb2.load_snapshot(str(path))
    assert tb2.training_step == 1234 and tb2.n_params == n
    assert np.array_equal(tb2.get_params(inference=True), w.astype(np.float16))
    assert np.array_equal(tb2.get_params(inference=False), w.astype(np.float16))
    bad = dict(cfg)
    bad["snapshot"] = dict(cfg["snapshot"], version=0)
    path.write_bytes(gzip.compress(msgpack.packb(bad)))
    with pytest.raises(P.NgpError, match="old format"):
        tb2.load_snapshot(str(path))


def test_loads_a_snapshot_written_by_the_reference_serializer_stack():
    """tests/golden/ref_snapshot.ingp was written by oracle/ref/ref_snapshot_harness.cu: nlohmann::json::to_msgpack through zstr's
    gzip stream, every structured value through the reference's own to_json (json_binding.h, vec_json.h), keys as
    Testbed::save_snapshot / Trainer::serialize assign them (src/testbed.cu:5288-5355, trainer.h:442-455)"""
    P = util.pkg()
    golden = Path(__file__).resolve().parent / "golden" / "ref_snapshot.ingp"
    tb = P.Testbed()
    tb.load_snapshot(str(golden))
    n = tb.n_params
    assert n == 141312 and tb.training_step == 1234 and abs(tb.loss - 0.00123) < 1e-8
    i = np.arange(n, dtype=np.uint64)
    pattern = ((((i * 37) % 1001).astype(np.float32) - 500.0) / 4000.0).astype(np.float16)
    assert np.array_equal(tb.get_params(inference=True), pattern)
    assert np.array_equal(tb.get_params(inference=False), pattern)
    grid, bits = tb.get_density_grid()
    k = np.arange(128 ** 3)
    want_grid = np.where(k % 7 == 0, np.float32(1.5), np.float32(0.001) * (k % 5).astype(np.float32)).astype(np.float16).astype(np.float32)
    assert np.array_equal(grid.reshape(-1)[: 128 ** 3], want_grid)
    # bitfield: cells above min(mean, NERF_MIN_OPTICAL_THICKNESS()) — here exactly the 1.5 cells
    # (update_density_grid_mean_and_bitfield, testbed_nerf.cu:3357-3377; mean ~ 0.2157 > 0.01)
    occ = np.unpackbits(bits[: 128 ** 3 // 8], bitorder="little")
    assert np.array_equal(occ.astype(bool), want_grid > 0.01)
    c = tb.counters()
    assert c["rays_per_batch"] == 4096 and c["measured_batch_size"] == 250000 and c["measured_batch_size_before_compaction"] == 300000
    # dataset metadata came from the file: render through view 1's OpenCV lens works and is finite
    cam = np.array([[0.0, 1.0, 0.0, 0.5], [0.6, 0.0, 0.8, 0.5], [0.8, 0.0, -0.6, 0.0]], dtype=np.float32)   # xforms[0].start of the file
    img = tb.render(64, 48, cam, 50.0)
    assert img.shape == (48, 64, 4) and np.isfinite(img).all()


REF_SNAPSHOT = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "ref_snapshot"


@pytest.mark.skipif(not REF_SNAPSHOT.exists(), reason="oracle/_ref/ref_snapshot not built (make -C oracle/ref snapshot; needs /root/reference)")
def test_the_reference_serializer_stack_reads_our_snapshot(trained, tmp_path):
    """the other direction: a file written by ngp_testbed_save_snapshot_ex goes through zstr + nlohmann::json::from_msgpack and the
    reference's from_json for BoundingBox / NerfDataset / Lens / TrainingXForm (the typed reads of Testbed::load_snapshot,
    src/testbed.cu:5357-5460); binary payloads are compared by checksum"""
    import json
    import subprocess

    tb, imgs, cams, focal, _, _ = trained
    path = tmp_path / "ours.ingp"
    tb.save_snapshot(str(path), include_optimizer_state=True)
    d = json.loads(subprocess.check_output([str(REF_SNAPSHOT), "dump", str(path)]))
    s = d["snapshot"]
    n = tb.n_params
    assert s["n_params"] == n and s["params_type"] == "__half" and s["version"] == 1 and s["mode"] == "nerf"
    params = tb.get_params(inference=True).tobytes()
    assert s["params_binary"]["bytes"] == 2 * n and s["params_binary"]["wsum64"] == util.wsum64(params)
    grid, _ = tb.get_density_grid()
    assert s["density_grid_binary"]["wsum64"] == util.wsum64(grid.astype(np.float16).tobytes())
    ds = s["nerf"]["dataset"]    # re-emitted by the reference's to_json(NerfDataset) after its from_json accepted ours
    assert ds["n_images"] == 24 and len(ds["xforms"]) == 24 and ds["aabb_scale"] == 1 and ds["wants_importance_sampling"] is True
    assert np.allclose(np.array(ds["xforms"][7]["start"]), np.asarray(cams[7])[:3, :4], atol=1e-6)
    assert ds["metadata"][0]["resolution"] == [160, 160] and abs(ds["metadata"][0]["focal_length"][1] - focal) < 1e-3
    assert s["aabb"] == {"min": [0.0, 0.0, 0.0], "max": [1.0, 1.0, 1.0]}
    adam = s["optimizer"]["nested"]["nested"]
    assert adam["current_step"] == tb.training_step and adam["first_moments_binary"]["bytes"] == 4 * n


def test_optimize_exposure_finds_the_darkened_views():
    """Nerf::Training::optimize_exposure (testbed_nerf.cu:2962-3000).  A view's colour is multiplied by 2^exposure before it becomes the target.
    A model trained on the clean views is frozen, three views are then replaced by half-bright copies: their exposures must rise towards one
    stop above the others' (with the network free to move, its view dependence simply absorbs a darker view and there is nothing left for the
    exposure to explain).  The exposures stay zero-mean, and nothing moves while the option is off."""
    P = util.pkg()
    tb = P.Testbed()
    # a narrow field of view: the ball fills the frame, every pixel carries colour
    imgs, cams, focal = S.make_dataset(n_images=16, width=96, height=96, fov_deg=16.0)
    assert imgs[..., 3].min() > 0.99
    S.load_into_testbed(tb, imgs, cams, focal, aabb_scale=1)
    tb.reload_network_from_json(S.base_config(16, 2, 16))
    for _ in range(300):
        tb.train(1 << 15)
    assert not any(tb.nerf.training.camera_exposure(i).any() for i in range(16))
    dark = [3, 8, 12]
    others = [i for i in range(16) if i not in dark]
    for i in dark:
        half = imgs[i].copy()
        half[..., :3] *= 0.5
        tb.nerf.training.set_image(i, half)
    tb._set("train_network", 0)
    tb._set("train_encoding", 0)
    tb.nerf.training.optimize_exposure = True
    tb.nerf.training.n_steps_between_cam_updates = 2     # one Adam step of at most ~1e-2 stops per update: 16 (the default) would need thousands of steps
    assert tb.nerf.training.optimize_exposure is True
    for _ in range(400):
        tb.train(1 << 15)
    e = np.stack([tb.nerf.training.camera_exposure(i) for i in range(16)])
    assert np.isfinite(e).all() and np.abs(e.mean(0)).max() < 1e-5
    gap = e[dark].mean() - e[others].mean()
    print("exposures (mean over channels):", np.round(e.mean(1), 3), "gap", gap)
    assert e[dark].mean(1).min() > e[others].mean(1).max()
    assert 0.4 < gap < 1.3          # one stop (measured 0.73 after 400 steps)
    # a set value restarts that view's optimizer and is applied
    tb.nerf.training.set_camera_exposure(0, [0.25, 0.25, 0.25])
    assert np.allclose(tb.nerf.training.camera_exposure(0), 0.25)
    tb.nerf.training.optimize_exposure = False
    before = np.stack([tb.nerf.training.camera_exposure(i) for i in range(16)])
    for _ in range(20):
        tb.train(1 << 15)
    assert np.array_equal(before, np.stack([tb.nerf.training.camera_exposure(i) for i in range(16)]))
    assert np.isfinite(tb.loss)


def test_update_image_async_lands_before_the_loss_kernel():
    """ngp_testbed_update_image_async: a replacement frame of a view without masked pixels is uploaded on a copy stream into a staging buffer
    and moved into the view's pixel buffer right before the next step's loss kernel (or by sync()); frames land in call order, also when more
    of them are queued than there are staging buffers"""
    import ctypes as C

    import torch

    P = util.pkg()
    B = importlib.import_module("instant-ngp_b200.binding")
    cudart = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    tb = P.Testbed()
    imgs, cams, focal = S.make_dataset(n_images=6, width=64, height=64)
    S.load_into_testbed(tb, imgs, cams, focal, aabb_scale=1)
    tb.reload_network_from_json(S.base_config(16, 2, 15))

    def device_pixels(i):
        v = B.TrainView()
        B.check(B.lib().ngp_testbed_get_view(tb._h, i, C.byref(v)))
        out = np.empty((64, 64, 4), dtype=np.float32)
        assert cudart.cudaMemcpy(out.ctypes.data, C.c_void_p(v.pixels), out.nbytes, 2) == 0
        return out

    for _ in range(3):
        tb.train(1 << 14)
    rng = np.random.default_rng(3)
    frames = [torch.from_numpy(np.ascontiguousarray(rng.uniform(0.0, 1.0, size=(64, 64, 4)).astype(np.float32))).pin_memory() for _ in range(5)]
    # one frame, one training step
    tb.update_image_async(2, frames[0].data_ptr())
    tb.train(1 << 14)
    tb.sync()
    assert np.array_equal(device_pixels(2), frames[0].numpy())
    # a frame per step, views in turn
    last = {2: 0}
    for k in range(25):
        f, view = k % 5, (k * 5) % 6
        tb.update_image_async(view, frames[f].data_ptr())
        last[view] = f
        tb.train(1 << 14)
    tb.sync()
    for view, f in last.items():
        assert np.array_equal(device_pixels(view), frames[f].numpy()), view
    # four frames queued without a step in between (two staging buffers), the same view twice: the later frame wins
    for k, (view, f) in enumerate([(0, 1), (1, 2), (0, 3), (4, 4)]):
        tb.update_image_async(view, frames[f].data_ptr())
    tb.sync()
    assert np.array_equal(device_pixels(0), frames[3].numpy()) and np.array_equal(device_pixels(1), frames[2].numpy()) and np.array_equal(device_pixels(4), frames[4].numpy())
    tb.train(1 << 14)
    assert np.isfinite(tb.loss)


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


def test_load_training_data_from_transforms_json(tmp_path):
    """Testbed.load_training_data (python_api.cu:452): a transforms.json scene on disk trains like the same scene set through the API"""
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_nerf_loader import write_scene

    imgs, cams, focal = write_scene(tmp_path, n=16, w=96, h=96)
    P = util.pkg()
    tb = P.Testbed()
    tb.load_training_data(str(tmp_path))
    assert len(tb.dataset["images"]) == 16 and tb.dataset["aabb_scale"] == 1
    tb.reload_network_from_json(S.base_config(16, 2, 15))
    before = psnr(np.clip(tb.render(96, 96, cams[3], focal)[..., :3], 0, 1), imgs[3][..., :3])
    for i in range(300):
        tb.train(1 << 16)
    assert np.isfinite(tb.loss)
    got = tb.render(96, 96, cams[3], focal)
    p = psnr(np.clip(got[..., :3], 0, 1), imgs[3][..., :3])
    print("psnr from disk: before", before, "after 300 steps", p)
    assert p > before + 3.0
