"""CPU: transforms.json semantics of the dataset loader (instant-ngp_b200/nerf_loader.py ≙ src/nerf_loader.cu:121-735) — on the
reference's own fox scene when /root/reference is present, and on a small scene written to disk."""
import importlib
import json
import math
from pathlib import Path

import numpy as np
import pytest

NL = importlib.import_module("instant-ngp_b200.nerf_loader")
FOX = Path("/root/reference/data/nerf/fox")


def ngp_to_nerf(m, scale=0.33, offset=(0.5, 0.5, 0.5)):
    """inverse of nerf_matrix_to_ngp, to author test scenes"""
    m = np.asarray(m, dtype=np.float64)[:3, :4]
    n = m[[2, 0, 1], :].copy()
    n[:, 3] = (n[:, 3] - np.asarray(offset)) / scale
    n[:, 1] *= -1
    n[:, 2] *= -1
    return np.vstack([n, [0, 0, 0, 1]])


def test_nerf_matrix_to_ngp_matches_the_reference_formula():
    # nerf_loader.h:101-120 written out by hand for one matrix
    m = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], dtype=np.float32)
    got = NL.nerf_matrix_to_ngp(m, 0.33, [0.5, 0.5, 0.5])
    want = np.array([[5, -6, -7, 8 * 0.33 + 0.5], [9, -10, -11, 12 * 0.33 + 0.5], [1, -2, -3, 4 * 0.33 + 0.5]], dtype=np.float32)
    assert np.allclose(got, want, atol=1e-6)
    assert np.allclose(NL.nerf_matrix_to_ngp(ngp_to_nerf(want), 0.33, [0.5, 0.5, 0.5]), want, atol=1e-5)


def test_natural_order_and_focal_length_precedence():
    names = ["img10.png", "img2.png", "img1.png", "a/img3.png"]
    assert sorted(names, key=NL.natural_key) == ["a/img3.png", "img1.png", "img2.png", "img10.png"]
    assert NL.read_focal_length({"camera_angle_x": math.radians(40.0)}, (800, 600)) == pytest.approx((0.5 * 800 / math.tan(math.radians(20)),) * 2)
    assert NL.read_focal_length({"camera_angle_x": 1.0, "fl_x": 123.0, "fl_y": 77.0}, (800, 600)) == (123.0, 77.0)
    assert NL.read_focal_length({"x_fov": 90.0, "fl_x": 123.0}, (800, 600)) == pytest.approx((400.0, 400.0))
    assert NL.read_focal_length({"camera_angle_y": math.radians(90.0)}, (800, 600)) == pytest.approx((300.0, 300.0))
    assert NL.read_focal_length({}, (800, 600)) is None


@pytest.mark.skipif(not FOX.exists(), reason="reference data not present (GPU box)")
def test_fox_metadata():
    ds = NL.load_metadata([FOX / "transforms.json"])
    raw = json.loads((FOX / "transforms.json").read_text())
    assert len(raw["frames"]) == 67 and len(ds["images"]) == 50          # frames without an image on disk are dropped (:386)
    assert ds["aabb_scale"] == 4 and ds["scale"] == 0.33 and ds["offset"] == [0.5, 0.5, 0.5]
    paths = [im["path"].name for im in ds["images"]]
    assert paths == sorted(paths, key=NL.natural_key) and paths[0] == "0001.jpg"
    lens = ds["images"][0]["lens"]
    assert lens["opencv"] and lens["params"] == [0.0578421, -0.0805099, -0.000980296, 0.00015575]
    assert lens["principal"] == pytest.approx((554.558 / 1080.0, 965.268 / 1920.0))
    assert NL.read_focal_length(ds["images"][0]["globals"], (1080, 1920)) == (1375.52, 1374.49)   # fl_x beats camera_angle_x
    f0 = next(f for f in raw["frames"] if f["file_path"].endswith("0001.jpg"))
    want = NL.nerf_matrix_to_ngp(np.asarray(f0["transform_matrix"], dtype=np.float32), 0.33, [0.5] * 3)
    assert np.array_equal(ds["images"][0]["xform"], want)
    # all cameras end up inside the aabb_scale-4 training box around (0.5, 0.5, 0.5)
    o = np.stack([im["xform"][:, 3] for im in ds["images"]])
    assert (np.abs(o - 0.5) < 2.0).all()


def write_scene(tmp_path, n=6, w=40, h=30, extra=None, frame_extra=None):
    from PIL import Image

    S = importlib.import_module("instant-ngp_b200.synthetic")
    imgs, cams, focal = S.make_dataset(n_images=n, width=w, height=h)
    frames = []
    (tmp_path / "images").mkdir()
    for i in range(n):
        a = imgs[i]
        straight = np.where(a[..., 3:4] > 0, a[..., :3] / np.maximum(a[..., 3:4], 1e-6), 0.0)
        srgb = np.where(straight <= 0.0031308, 12.92 * straight, 1.055 * np.power(np.maximum(straight, 1e-12), 1 / 2.4) - 0.055)
        px = np.concatenate([np.clip(srgb, 0, 1), a[..., 3:4]], axis=-1)
        Image.fromarray((px * 255 + 0.5).astype(np.uint8), "RGBA").save(tmp_path / "images" / f"frame_{i + 1}.png")
        fr = {"file_path": f"images/frame_{i + 1}", "transform_matrix": ngp_to_nerf(cams[i]).tolist()}   # extension left to resolve_path
        fr.update((frame_extra or {}).get(i, {}))
        frames.append(fr)
    t = {"camera_angle_x": 2 * math.atan(0.5 * w / focal), "aabb_scale": 1, "frames": frames[::-1]}   # written out of order on purpose
    t.update(extra or {})
    (tmp_path / "transforms.json").write_text(json.dumps(t))
    return imgs, cams, focal


def test_scene_on_disk(tmp_path):
    imgs, cams, focal = write_scene(tmp_path, n=12, frame_extra={3: {"fl_x": 55.0}})
    ds = NL.load_metadata(NL.find_transforms(tmp_path))
    assert [im["path"].name for im in ds["images"]] == [f"frame_{i + 1}.png" for i in range(12)]     # natural order: 2 before 10
    for i in range(12):
        assert np.allclose(ds["images"][i]["xform"], np.asarray(cams[i])[:3, :4], atol=2e-5)
    rgba = NL.read_image_linear_rgba(ds["images"][0]["path"])
    assert rgba.shape == (30, 40, 4) and rgba.dtype == np.float32
    want = imgs[0]
    straight = np.where(want[..., 3:4] > 0, want[..., :3] / np.maximum(want[..., 3:4], 1e-6), 0.0)
    inside = want[..., 3] > 0.99
    assert np.abs(rgba[..., :3][inside] - straight[inside]).max() < 0.02 and np.abs(rgba[..., 3] - want[..., 3]).max() < 0.01
    assert NL.read_focal_length(ds["images"][3]["frame"], (40, 30)) == (55.0, 55.0)                  # per-frame override
    with pytest.raises(ValueError):
        NL.load_metadata([])


def test_sharpness_culling_and_aabb_fitting(tmp_path):
    write_scene(tmp_path, n=8, extra={"sharpness_discard_threshold": 0.9, "aabb": [[-2, -1, -1], [2, 1, 1]]},
                frame_extra={i: {"sharpness": (10.0 if i != 4 else 1.0)} for i in range(8)})
    ds = NL.load_metadata([tmp_path / "transforms.json"])
    assert len(ds["images"]) == 7 and "frame_5.png" not in [im["path"].name for im in ds["images"]]   # the blurry frame is dropped
    assert ds["scale"] == pytest.approx(0.25) and ds["offset"] == pytest.approx([0.5, 0.5, 0.5])      # longest aabb side 4 -> unit cube
