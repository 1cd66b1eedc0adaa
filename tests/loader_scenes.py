"""Deterministic NeRF scenes on disk (transforms.json + PNG images) that exercise the reference loader's semantics
(ngp::load_nerf, src/nerf_loader.cu:271-735).  Written once on the GPU box for oracle/_ref/ref_loader (the reference's own loader,
tools/make_ref_loader_golden.sh -> tests/golden/ref_loader.json) and again on the CPU for the product's loader
(tests/test_nerf_loader.py): PNG is lossless, so both sides decode the same pixels.

    python tests/loader_scenes.py <root>        # writes <root>/<scene>/...
"""
import importlib
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def ngp_to_nerf(m, scale=0.33, offset=(0.5, 0.5, 0.5)):
    """inverse of nerf_matrix_to_ngp (nerf_loader.h:101-120), to author transform_matrix entries"""
    m = np.asarray(m, dtype=np.float64)[:3, :4]
    n = m[[2, 0, 1], :].copy()
    n[:, 3] = (n[:, 3] - np.asarray(offset)) / scale
    n[:, 1] *= -1
    n[:, 2] *= -1
    return np.vstack([n, [0, 0, 0, 1]])


def _srgb_bytes(img):
    """linear premultiplied float RGBA -> straight-alpha sRGB bytes"""
    a = img[..., 3:4]
    straight = np.where(a > 0, img[..., :3] / np.maximum(a, 1e-6), 0.0)
    srgb = np.where(straight <= 0.0031308, 12.92 * straight, 1.055 * np.power(np.maximum(straight, 1e-12), 1 / 2.4) - 0.055)
    px = np.concatenate([np.clip(srgb, 0, 1), a], axis=-1)
    return (px * 255 + 0.5).astype(np.uint8)


def _dataset(n, w, h):
    S = importlib.import_module("instant-ngp_b200.synthetic")
    return S.make_dataset(n_images=n, width=w, height=h)


def _write(root, name, n=6, w=40, h=30, extra=None, frame_extra=None, mode="RGBA", ext_in_path=False, json_name="transforms.json", first=0,
           reverse=True, drop_angle=False, subdir="images"):
    from PIL import Image

    d = root / name
    (d / subdir).mkdir(parents=True, exist_ok=True)
    imgs, cams, focal = _dataset(first + n, w, h)
    frames = []
    for i in range(first, first + n):
        px = _srgb_bytes(imgs[i])
        im = Image.fromarray(px, "RGBA")
        if mode == "RGB":
            # composite over white: exercises white_transparent
            a = px[..., 3:4].astype(np.float32) / 255.0
            rgb = (px[..., :3].astype(np.float32) * a + 255.0 * (1.0 - a) + 0.5).astype(np.uint8)
            im = Image.fromarray(rgb, "RGB")
        stem = f"frame_{i + 1}"
        im.save(d / subdir / f"{stem}.png")
        fr = {"file_path": f"{subdir}/{stem}" + (".png" if ext_in_path else ""), "transform_matrix": ngp_to_nerf(cams[i]).tolist()}
        fr.update((frame_extra or {}).get(i, {}))
        frames.append(fr)
    t = {"aabb_scale": 1, "frames": frames[::-1] if reverse else frames}
    if not drop_angle:
        t["camera_angle_x"] = 2 * math.atan(0.5 * w / focal)
    t.update(extra or {})
    (d / json_name).write_text(json.dumps(t, indent=1))
    return d


def write_all(root):
    root = Path(root)
    root.mkdir(parents=True, exist_ok=True)
    from PIL import Image

    # 1. plain: frames out of order, extensionless paths resolved against the supported formats, camera_angle_x
    _write(root, "basic", n=12)
    # 2. global OpenCV intrinsics, one per-frame focal override, one per-frame distortion override, scale / offset / up / render_aabb
    _write(root, "opencv", n=6, drop_angle=True, ext_in_path=True,
           extra={"fl_x": 61.5, "fl_y": 60.25, "cx": 19.5, "cy": 15.25, "w": 40, "h": 30, "k1": 0.05, "k2": -0.02, "p1": 0.001, "p2": -0.0005, "aabb_scale": 4,
                  "scale": 0.5, "offset": [0.4, 0.5, 0.6], "up": [0.0, 0.0, 1.0], "render_aabb": [[-0.5, -0.25, 0.0], [1.5, 1.25, 1.0]],
                  "camera_angle_x": 1.0},
           frame_extra={3: {"fl_x": 55.0}, 2: {"k1": 0.1, "cx": 21.0, "w": 40}, 4: {"fl_y": 70.0}})
    # 3. sharpness culling (half-open neighbourhood mean), n_frames, aabb fitted into the unit cube
    _write(root, "culled", n=9, extra={"sharpness_discard_threshold": 0.9, "aabb": [[-2, -1, -1], [2, 1, 1]], "n_frames": 8},
           frame_extra={i: {"sharpness": (10.0 if i not in (4, 6) else (1.0 if i == 4 else 8.9))} for i in range(9)})
    # 4. RGB images over white + white_transparent; scalar offset; x_fov in degrees beats fl_x
    _write(root, "white", n=4, mode="RGB", drop_angle=True, extra={"white_transparent": True, "offset": 0.25, "x_fov": 40.0, "fl_x": 1.0, "aabb_scale": 2})
    # 5. black_transparent on RGBA, from_mitsuba scale / offset, camera_angle_y only, importance_sampling off
    _write(root, "mitsuba", n=4, drop_angle=True, extra={"black_transparent": True, "from_mitsuba": True, "camera_angle_y": 0.7, "importance_sampling": False,
                                                       "n_extra_learnable_dims": 0})
    # 6. a separate alpha image: <file_path>.alpha.<ext>, red channel, sRGB -> linear
    d = _write(root, "alpha_file", n=3, ext_in_path=False)
    for i in range(3):
        g = np.tile(np.linspace(0, 255, 40, dtype=np.uint8)[None, :, None], (30, 1, 4))
        g[..., 3] = 255
        Image.fromarray(np.ascontiguousarray(g), "RGBA").save(d / "images" / f"frame_{i + 1}.alpha.png")
    # 7. two transform files in one directory (later keys win; scale / offset are left alone: the reference converts the matrices on
    # pool threads that read result.scale while the main thread may already be parsing the next file, nerf_loader.cu:536-706)
    d = _write(root, "two_files", n=4, json_name="transforms_a.json", subdir="a")
    _write(root, "two_files", n=3, first=4, json_name="transforms_b.json", subdir="b", extra={"aabb_scale": 2, "k1": 0.03})
    # 8. what the product does not load (kept in the golden as the specification): fisheye + rolling shutter, dynamic masks, 16-bit depth
    _write(root, "fisheye_rs", n=3, extra={"is_fisheye": True, "k1": 0.01, "k2": 0.02, "k3": 0.03, "k4": 0.04, "rolling_shutter": [0.0, 0.0, 0.5, 0.25]})
    d = _write(root, "masked", n=3, ext_in_path=True)
    m = np.zeros((30, 40, 3), dtype=np.uint8)
    m[5:12, 8:20] = 255
    Image.fromarray(m, "RGB").save(d / "images" / "dynamic_mask_frame_2.png")
    d = _write(root, "depth", n=3, ext_in_path=True, extra={"integer_depth_scale": 0.001, "enable_depth_loading": True},
               frame_extra={i: {"depth_path": f"images/depth_{i + 1}.png"} for i in range(3)})
    for i in range(3):
        z = (np.arange(30 * 40, dtype=np.uint32).reshape(30, 40) * 7 + 1000 * (i + 1)).astype(np.uint16)
        Image.fromarray(z).save(d / "images" / f"depth_{i + 1}.png")
    return sorted(p.name for p in root.iterdir() if p.is_dir())


if __name__ == "__main__":
    print(" ".join(write_all(Path(sys.argv[1]))))
