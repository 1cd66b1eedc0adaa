"""CPU: the host-side camera state behind Testbed.render(width, height, spp, linear) (instant-ngp_b200/camera.py ≙
src/testbed.cu:486-528, 4081-4087, 4649-4657; fov helpers common_device.cuh:645-659) and the pyngp surface a script touches."""
import importlib
import math

import numpy as np
import pytest

CAM = importlib.import_module("instant-ngp_b200.camera")
P = importlib.import_module("instant-ngp_b200")


def test_fov_and_focal_length_are_inverse_and_match_the_reference_formulas():
    for deg in (10.0, 39.6, 50.625, 90.0, 120.0):
        f = CAM.fov_to_focal_length(800, deg)
        assert f == pytest.approx(0.5 * 800 / math.tan(0.5 * deg * math.pi / 180.0))
        assert CAM.focal_length_to_fov(800, f) == pytest.approx(deg, rel=1e-12)
    assert CAM.fov_to_focal_length(1, 90.0) == pytest.approx(0.5)


def test_reset_camera_state():
    c = CAM.CameraState()
    assert c.fov_axis == 1 and c.zoom == 1.0 and c.screen_center == (0.5, 0.5) and c.scale == 1.5
    assert c.fov() == pytest.approx(50.625) and c.fov_xy() == pytest.approx((50.625, 50.625))
    # m_default_camera moved back by m_scale along the view direction (column 2)
    assert np.array_equal(c.matrix, np.array([[1, 0, 0, 0.5], [0, -1, 0, 0.5], [0, 0, -1, 2.0]], dtype=np.float32))


def test_render_arguments():
    c = CAM.CameraState()
    c.set_fov(90.0)
    cam, focal, center = c.render_args(640, 480)
    assert focal == pytest.approx((240.0, 240.0)) and center == (0.5, 0.5)          # fov_axis = 1: relative to the height
    c.fov_axis = 0
    assert c.render_args(640, 480)[1] == pytest.approx((320.0, 320.0))
    c.zoom = 2.0
    c.screen_center = (0.25, 0.5)
    cam, focal, center = c.render_args(640, 480)
    assert focal == pytest.approx((640.0, 640.0)) and center == pytest.approx((1.0, 0.5))   # (0.5 - sc) * zoom + 0.5
    c.set_fov_xy((90.0, 2 * math.degrees(math.atan(0.25))))
    assert c.relative_focal_length == pytest.approx((0.5, 2.0))


def test_training_view_round_trip():
    c = CAM.CameraState()
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    c.to_training_view(x, (1375.52, 1374.49), (1080, 1920), (554.558 / 1080.0, 965.268 / 1920.0))
    cam, focal, center = c.render_args(1080, 1920)
    assert np.array_equal(cam, x) and focal == pytest.approx((1375.52, 1374.49)) and center == pytest.approx((554.558 / 1080.0, 965.268 / 1920.0))
    # a quarter-size frame keeps the field of view
    assert c.render_args(270, 480)[1] == pytest.approx((1375.52 / 4, 1374.49 / 4))
    x[0, 0] = 99.0
    assert c.matrix[0, 0] == 0.0                                                      # the state owns its matrix


def test_pyngp_surface_names():
    """the members scripts/run.py and the BASELINE configs touch, present on the class without a device"""
    for name in ("load_training_data", "reload_network_from_file", "reload_network_from_json", "load_file", "train", "frame", "render",
                 "render_with_depth", "save_snapshot", "load_snapshot", "reset", "reset_camera", "training_step", "loss", "n_params", "shall_train",
                 "background_color", "exposure", "snap_to_pixel_centers", "render_mode", "color_space", "fov", "fov_axis", "fov_xy", "zoom",
                 "screen_center", "set_nerf_camera_matrix", "set_camera_to_training_view", "create_empty_nerf_dataset", "set_seed"):
        assert hasattr(P.Testbed, name), name
    for name in ("compute_image_mse", "override_sdf_training_data", "set_image", "reload_network_from_json", "train", "render", "loss", "training_step"):
        assert hasattr(P.FieldTestbed, name), name
    assert {m.name for m in P.RenderMode} >= {"Shade", "Depth", "Normals"} and {m.name for m in P.TrainMode} == {"Nerf", "Rfl", "RflRelax"}


def test_render_dispatches_between_the_reference_form_and_the_explicit_camera(monkeypatch):
    """no device: _render_explicit is replaced by a recorder"""
    tb = object.__new__(P.Testbed)
    tb._h = None
    tb._camera = CAM.CameraState()
    calls = []
    monkeypatch.setattr(P.Testbed, "_render_explicit", lambda self, *a, **k: calls.append((a, k)) or "img")
    cam = np.eye(4, dtype=np.float32)[:3]
    assert tb.render(64, 48, cam, 50.0) == "img" and calls[-1] == ((64, 48, cam, 50.0), {})
    tb.render(64, 48, cam, 50.0, spp=4, linear=False, rows=(0, 8))
    assert calls[-1][1] == {"spp": 4, "linear": False, "rows": (0, 8)}
    tb.render(64, 48, camera_matrix=cam, focal_length=(50.0, 51.0))
    assert calls[-1][1]["focal_length"] == (50.0, 51.0)
    tb._camera.set_fov(90.0)
    tb.render(64, 48, 2, False)                                                       # the reference's positional form: spp, linear
    a, k = calls[-1]
    assert a[:2] == (64, 48) and np.array_equal(a[2], tb._camera.matrix) and a[3] == pytest.approx((24.0, 24.0)) and a[4] == (0.5, 0.5)
    assert k == {"spp": 2, "linear": False, "return_depth": False}
    tb.render(64, 48)
    assert calls[-1][1] == {"spp": 1, "linear": True, "return_depth": False}
    tb.render_with_depth(32, 32, 1, True)
    assert calls[-1][1]["return_depth"] is True
    with pytest.raises(TypeError):
        tb.render(64, 48, 1, spp=2)
    with pytest.raises(TypeError):
        tb.render(64, 48, bogus=1)
    with pytest.raises(P.NgpError):
        tb.render(64, 48, 1, True, 0.0, 1.0)
    tb._h = None
