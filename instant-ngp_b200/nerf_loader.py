"""NeRF dataset loading: `transforms.json` semantics of the reference's loader (src/nerf_loader.cu:121-735, read_lens :23-89,
read_focal_length :91-119) restated for the `Testbed.load_training_data` entry point (python_api.cu:452).

The metadata logic (frame ordering, culling by sharpness, scale / offset / aabb, lens and focal-length precedence, the
nerf -> ngp coordinate change) lives here; pixels are decoded with PIL and handed to the C++ Testbed through the same
`set_image` / `set_camera_*` calls a user would make (python_api.cu:809-853).  8-bit images are uploaded as bytes (sRGB colour, straight alpha:
`ngp_testbed_set_image_bytes` ≙ set_training_image with EImageDataType::Byte) and converted to linear premultiplied colour on
every read, as the reference does (common_device.cuh:698-735).

Pinned against the reference's own `ngp::load_nerf`, compiled from /root/reference and run over the scenes of
tests/loader_scenes.py (oracle/ref/ref_loader_harness.cu -> tests/golden/ref_loader.json, tests/test_nerf_loader.py): frame order
and culling, scale / offset / up / render_aabb, per-image focal length, principal point, lens, transform and the stored pixel bytes.

Not supported (raises): EXR images, depth supervision, per-pixel ray files, rolling shutter, fisheye /
f-theta / lat-long lenses, environment maps.  Several transform files with different `scale` / `offset` use the final values for
every frame (the reference converts matrices on pool threads racing the parse of the next file, nerf_loader.cu:536-706)."""
from __future__ import annotations

import json
import math
import os
import re
from pathlib import Path

import numpy as np

NERF_SCALE = 0.33
SUPPORTED_FORMATS = ("png", "jpg", "jpeg", "bmp", "gif", "tga", "pnm")


def _strip_json_comments(text: str) -> str:
    return re.sub(r"//[^\n]*|/\*.*?\*/", "", text, flags=re.S)


def natural_key(s: str):
    """SI::natural::compare (dependencies/NaturalSort): digit runs compare as numbers"""
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def fov_to_focal_length(resolution: int, degrees: float) -> float:
    return 0.5 * float(resolution) / math.tan(0.5 * degrees * math.pi / 180.0)


def read_focal_length(j: dict, res):
    """read_focal_length (nerf_loader.cu:91-119): x_fov (degrees) > fl_x > camera_angle_x (radians), per axis"""
    def one(resolution, axis):
        if axis + "_fov" in j:
            return fov_to_focal_length(resolution, float(j[axis + "_fov"]))
        if "fl_" + axis in j:
            return float(j["fl_" + axis])
        if "camera_angle_" + axis in j:
            return fov_to_focal_length(resolution, float(j["camera_angle_" + axis]) * 180.0 / math.pi)
        return 0.0

    x, y = one(res[0], "x"), one(res[1], "y")
    if x != 0:
        return (x, y if y != 0 else x)
    if y != 0:
        return (y, y)
    return None


def read_lens(j: dict, lens: dict) -> dict:
    """read_lens (nerf_loader.cu:23-89) for the perspective / OpenCV models; `lens` carries the inherited values"""
    out = dict(lens)
    if j.get("is_fisheye", False) or any(k in j for k in ("ftheta_p0", "latlong", "equirectangular", "orthographic")):
        raise ValueError("only perspective and OpenCV lenses are supported")
    # k3 / k4 land in the slots p1 / p2 overwrite afterwards, exactly as in the reference (:41-47)
    for name, idx in (("k1", 0), ("k2", 1), ("k3", 2), ("k4", 3), ("p1", 2), ("p2", 3)):
        if name in j:
            out["params"] = list(out["params"])
            out["params"][idx] = float(j[name])
            if out["params"][idx] != 0.0:
                out["opencv"] = True
    if "cx" in j:
        out["principal"] = (float(j["cx"]) / float(j["w"]), out["principal"][1])
    if "cy" in j:
        out["principal"] = (out["principal"][0], float(j["cy"]) / float(j["h"]))
    if "rolling_shutter" in j and any(float(v) != 0 for v in j["rolling_shutter"]):
        raise ValueError("rolling shutter is not supported")
    return out


def nerf_matrix_to_ngp(m: np.ndarray, scale: float, offset, from_mitsuba: bool = False) -> np.ndarray:
    """NerfDataset::nerf_matrix_to_ngp (nerf_loader.h:101-120): flip the y and z columns, scale + offset the origin, then cycle the
    axes xyz <- yzx — or, for Mitsuba scenes, negate the x and z columns instead"""
    m = np.asarray(m, dtype=np.float32)[:3, :4].copy()
    m[:, 1] *= -1
    m[:, 2] *= -1
    m[:, 3] = m[:, 3] * np.float32(scale) + np.asarray(offset, dtype=np.float32)
    if from_mitsuba:
        m[:, 0] *= -1
        m[:, 2] *= -1
        return m
    return m[[1, 2, 0], :]


def resolve_path(base: Path, local: str) -> Path:
    p = Path(local.replace("\\", "/"))
    p = p if p.is_absolute() else base / p
    if p.suffix == "" and not p.exists():
        for fmt in SUPPORTED_FORMATS + ("exr",):
            if p.with_suffix("." + fmt).exists():
                return p.with_suffix("." + fmt)
    return p


def load_metadata(json_paths) -> dict:
    """Everything of load_nerf that does not touch pixels.  Returns the dataset description: per image `path`, `xform`
    (3x4, ngp convention), `focal_length`, `principal_point`, `lens`; and `scale`, `offset`, `aabb_scale`, `up`, `render_aabb`."""
    json_paths = [Path(p) for p in json_paths]
    if not json_paths:
        raise ValueError("Cannot load NeRF data from an empty set of paths.")
    ds = dict(images=[], scale=NERF_SCALE, offset=[0.5, 0.5, 0.5], aabb_scale=1, up=[0.0, 1.0, 0.0], render_aabb=None, from_mitsuba=False,
              white_transparent=False, black_transparent=False, wants_importance_sampling=True, n_extra_learnable_dims=0)
    for jp in json_paths:
        j = json.loads(_strip_json_comments(jp.read_text()))
        if isinstance(j.get("camera"), list):
            raise ValueError("hdf5 is no longer supported. please use the hdf52nerf.py conversion script")
        if not isinstance(j.get("frames"), list):
            continue
        base = jp.parent
        frames = sorted(j["frames"], key=lambda f: natural_key(f["file_path"]))
        if "n_frames" in j:
            frames = frames[: min(len(frames), int(j["n_frames"]))]
        if frames and "sharpness" in frames[0]:
            thr = float(j.get("sharpness_discard_threshold", 0.0))
            kept, nb = [], 3
            for i, f in enumerate(frames):
                lo, hi = max(0, i - nb), min(i + nb, len(frames) - 1)
                mean = sum(float(frames[k].get("sharpness", 1.0)) for k in range(lo, hi)) / (hi - lo)   # (the reference's half-open window)
                if resolve_path(base, f["file_path"]).exists() and float(f.get("sharpness", 1.0)) > thr * mean:
                    kept.append(f)
            frames = kept
        if "normal_mts_args" in j:
            ds["from_mitsuba"] = True
        if "from_mitsuba" in j:
            ds["from_mitsuba"] = bool(j["from_mitsuba"])
        if ds["from_mitsuba"]:
            ds["scale"] = 0.66
            ds["offset"] = [0.25 * 0.66] * 3
        if "render_aabb" in j:
            ds["render_aabb"] = [[float(v) for v in j["render_aabb"][0]], [float(v) for v in j["render_aabb"][1]]]
        for k in ("white_transparent", "black_transparent"):
            if k in j:
                ds[k] = bool(j[k])
        if "scale" in j:
            ds["scale"] = float(j["scale"])
        if "importance_sampling" in j:
            ds["wants_importance_sampling"] = bool(j["importance_sampling"])
        if "n_extra_learnable_dims" in j:
            ds["n_extra_learnable_dims"] = int(j["n_extra_learnable_dims"])
            if ds["n_extra_learnable_dims"]:
                raise ValueError("'n_extra_learnable_dims' is not supported")
        for k in ("enable_depth_loading", "integer_depth_scale", "envmap"):
            if j.get(k):
                raise ValueError(f"'{k}' is not supported")
        lens = read_lens(j, dict(params=[0.0, 0.0, 0.0, 0.0], opencv=False, principal=(0.5, 0.5)))
        if "aabb_scale" in j:
            ds["aabb_scale"] = int(j["aabb_scale"])
        if "offset" in j:
            o = j["offset"]
            ds["offset"] = [float(v) for v in o] if isinstance(o, list) else [float(o)] * 3
        if "aabb" in j:
            a = j["aabb"]
            length = max(1e-6, max(abs(float(a[1][k]) - float(a[0][k])) for k in range(3)))
            ds["scale"] = 1.0 / length
            ds["offset"] = [((float(a[1][k]) + float(a[0][k])) * 0.5) * -ds["scale"] + 0.5 for k in range(3)]
        if "up" in j:
            ds["up"] = [float(j["up"][1]), float(j["up"][2]), float(j["up"][0])]
        for f in frames:
            path = resolve_path(base, f["file_path"])
            if not path.exists():
                raise FileNotFoundError(f"Could not find image file '{path}'.")
            if "transform_matrix_start" in f or "transform_matrix_end" in f:
                raise ValueError("per-frame start / end transforms (motion blur) are not supported")
            if (path.parent / f"rays_{path.stem}.dat").exists() and j.get("enable_ray_loading", True):
                raise ValueError("per-pixel ray files are not supported")
            if "driver_parameters" in f:
                raise ValueError("light directions (driver_parameters) are not supported")
            ds["images"].append(dict(path=path, json_path=str(f["file_path"]), frame=f, globals=j, lens=read_lens(f, lens)))
    if not ds["images"]:
        raise ValueError("No training images were found for NeRF training!")
    # the transform needs the final scale / offset (the reference applies them as it walks the files; one file = same thing)
    for im in ds["images"]:
        im["xform"] = nerf_matrix_to_ngp(np.asarray(im["frame"]["transform_matrix"], dtype=np.float32), ds["scale"], ds["offset"], ds["from_mitsuba"])
        # focal length: the file's, overridden by the frame's (nerf_loader.cu:676-680); needs the image size for the fov forms
        im["resolution"] = image_size(im["path"])
        fl = read_focal_length(im["frame"], im["resolution"]) or read_focal_length(im["globals"], im["resolution"])
        if fl is None:
            raise ValueError("Couldn't read fov.")
        im["focal_length"] = fl
    return ds


def image_size(path: Path):
    """(width, height) from the file header"""
    from PIL import Image

    if Path(path).suffix.lower() == ".exr":
        raise ValueError("EXR images are not supported")
    with Image.open(path) as im:
        return im.size


def read_image_bytes_rgba(path: Path, white_transparent=False, black_transparent=False) -> np.ndarray:
    """[H, W, 4] uint8, sRGB colour + straight alpha: the Byte image as the reference stores it — load_stbi(..., 4), the optional
    `<name>.alpha.<ext>` companion (nerf_loader.cu:581-599) and convert_rgba32's white / black -> transparent (:41-63)"""
    from PIL import Image

    path = Path(path)
    if path.suffix.lower() == ".exr":
        raise ValueError("EXR images are not supported")
    img = np.array(Image.open(path).convert("RGBA"), dtype=np.uint8)
    alpha_path = Path(f"{path.with_suffix('')}.alpha{path.suffix}")
    if alpha_path.exists():
        a = np.asarray(Image.open(alpha_path).convert("RGBA"), dtype=np.uint8)
        if a.shape != img.shape:
            raise ValueError(f"Alpha image {alpha_path} has wrong resolution.")
        r = a[..., 0].astype(np.float32) * np.float32(1.0 / 255.0)
        lin = np.where(r <= np.float32(0.04045), r / np.float32(12.92), np.power((r + np.float32(0.055)) / np.float32(1.055), np.float32(2.4)))
        img[..., 3] = (np.float32(255.0) * lin.astype(np.float32)).astype(np.uint8)
    # dynamic masks (nerf_loader.cu:600-619): `dynamic_mask_<stem>.png` beside the frame; every pixel where the mask is not black becomes
    # MASK_COLOR 0x00FF00FF (R 255, G 0, B 255, A 0) — what read_rgba reports as "masked away" (negative colour): the sample generator
    # draws no ray through it (testbed_nerf.cu:732-736)
    mask_path = path.parent / f"dynamic_mask_{path.stem}.png"
    if mask_path.exists():
        m = np.asarray(Image.open(mask_path).convert("RGBA"), dtype=np.uint8)
        if m.shape != img.shape:
            raise ValueError(f"Dynamic mask {mask_path} has wrong resolution.")
        img[(m[..., :3] != 0).any(axis=-1)] = (255, 0, 255, 0)
    if white_transparent:
        img[(img[..., :3] == 255).all(axis=-1), 3] = 0
    if black_transparent:
        img[(img[..., :3] == 0).all(axis=-1), 3] = 0
    return img


def read_image_linear_rgba(path: Path, white_transparent=False, black_transparent=False) -> np.ndarray:
    """decode to [H, W, 4] float32, linear colour, straight alpha (the Byte branch of read_rgba, common_device.cuh:698-735)"""
    img = read_image_bytes_rgba(path, white_transparent, black_transparent)
    s = img[..., :3].astype(np.float32) / 255.0
    out = np.empty(img.shape, dtype=np.float32)
    out[..., :3] = np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)
    out[..., 3] = img[..., 3].astype(np.float32) / 255.0
    return out


def find_transforms(path) -> list:
    """Testbed::load_training_data accepts a transforms .json, or a directory holding transforms*.json (testbed_nerf.cu:2375-2410)"""
    p = Path(path)
    if p.is_dir():
        found = sorted(q for q in p.iterdir() if q.suffix == ".json")
        if not found:
            raise FileNotFoundError(f"no transforms .json under '{p}'")
        return found
    if not p.exists():
        raise FileNotFoundError(f"Data path '{p}' does not exist.")
    return [p]


def load_into_testbed(tb, path) -> dict:
    """≙ Testbed::load_training_data(path) for NeRF scenes (python_api.cu:452, testbed_nerf.cu:2375-2440)"""
    ds = load_metadata(find_transforms(path))
    n = len(ds["images"])
    tb.create_empty_nerf_dataset(n, aabb_scale=ds["aabb_scale"])
    tb._set("nerf.training.dataset.scale", ds["scale"])
    for k, ax in enumerate("xyz"):
        tb._set(f"nerf.training.dataset.offset.{ax}", ds["offset"][k])
    for i, im in enumerate(ds["images"]):
        # 8-bit files stay bytes (EImageDataType::Byte, nerf_loader.cu:606-612): read_rgba converts sRGB -> linear and premultiplies
        # by alpha on every read (common_device.cuh:698-735), exactly what the loss kernel's Byte branch does here
        rgba8 = read_image_bytes_rgba(im["path"], ds["white_transparent"], ds["black_transparent"])
        fl = im["focal_length"]
        tb.nerf.training.set_image_bytes(i, rgba8)
        tb.nerf.training.set_camera_extrinsics(i, im["xform"], convert_to_ngp=False)
        lens = im["lens"]
        k = lens["params"] if lens["opencv"] else [0.0, 0.0, 0.0, 0.0]
        tb.nerf.training.set_camera_intrinsics(i, fx=fl[0], fy=fl[1], cx=-lens["principal"][0], cy=-lens["principal"][1], k1=k[0], k2=k[1], p1=k[2], p2=k[3])
    tb.nerf.training.n_images_for_training = n
    return ds
