"""Procedural NeRF datasets (there is no network for NeRF-synthetic, and the fox JPEGs cannot travel to the GPU box).

Cameras sit on a sphere around the scene centre (0.5, 0.5, 0.5) in ngp coordinates and look at it; images are analytic
renders of a shaded, textured ball (premultiplied linear RGBA), which gives the occupancy grid and the network something
real to converge on.  Used by bench.py and the tests; deterministic given the arguments.
"""
from __future__ import annotations

import numpy as np


def look_at_camera(pos, target=(0.5, 0.5, 0.5), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """ngp-convention camera-to-world 3x4 (columns: right, down, forward, origin)."""
    pos, target, up = (np.asarray(v, dtype=np.float64) for v in (pos, target, up))
    fwd = target - pos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, up)
    if np.linalg.norm(right) < 1e-6:
        right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.stack([right, down, fwd, pos], axis=1)
    return m.astype(np.float32)


def sphere_cameras(n: int, radius: float = 1.3, seed: int = 0) -> list:
    """n cameras on a Fibonacci sphere (upper 80 % of the sphere), deterministic."""
    cams = []
    ga = np.pi * (3.0 - np.sqrt(5.0))
    for i in range(n):
        z = 0.9 - 1.7 * (i + 0.5) / n  # avoid the exact poles
        r = np.sqrt(max(0.0, 1.0 - z * z))
        th = ga * i + 0.1 * seed
        p = np.array([0.5 + radius * r * np.cos(th), 0.5 + radius * r * np.sin(th), 0.5 + radius * z])
        cams.append(look_at_camera(p))
    return cams


def render_ball(cam: np.ndarray, width: int, height: int, focal: float, ball_radius: float = 0.28) -> np.ndarray:
    """Analytic image of a textured ball centred in the cube: float32 [H, W, 4], linear, premultiplied."""
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    u = (xs + 0.5) / width
    v = (ys + 0.5) / height
    d_cam = np.stack([(u - 0.5) * width / focal, (v - 0.5) * height / focal, np.ones_like(u)], axis=-1)
    R = cam[:, :3].astype(np.float64)
    o = cam[:, 3].astype(np.float64)
    d = d_cam @ R.T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    oc = o - 0.5
    b = (d * oc).sum(-1)
    c = (oc * oc).sum() - ball_radius ** 2
    disc = b * b - c
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0.0))
    hit &= t > 0
    p = o + d * t[..., None]
    nrm = (p - 0.5) / ball_radius
    light = np.array([0.4, 0.3, 0.85])
    light /= np.linalg.norm(light)
    shade = 0.25 + 0.75 * np.clip((nrm * light).sum(-1), 0, 1)
    # latitude / longitude checker so that views disagree unless the geometry is right
    lon = np.arctan2(nrm[..., 1], nrm[..., 0])
    lat = np.arcsin(np.clip(nrm[..., 2], -1, 1))
    check = ((np.floor(lon / (np.pi / 6)) + np.floor(lat / (np.pi / 8))) % 2).astype(np.float64)
    base = np.stack([0.85 - 0.5 * check, 0.35 + 0.4 * check, 0.25 + 0.5 * (nrm[..., 2] * 0.5 + 0.5)], axis=-1)
    rgb = base * shade[..., None]
    a = hit.astype(np.float64)
    img = np.concatenate([rgb * a[..., None], a[..., None]], axis=-1)
    return img.astype(np.float32)


def make_dataset(n_images: int = 100, width: int = 800, height: int = 800, fov_deg: float = 40.0, radius: float = 1.3, seed: int = 0):
    """returns (images [n,H,W,4] float32, cameras list of 3x4 ngp matrices, focal length in pixels)"""
    focal = 0.5 * width / np.tan(0.5 * np.deg2rad(fov_deg))
    cams = sphere_cameras(n_images, radius=radius, seed=seed)
    imgs = np.stack([render_ball(c, width, height, focal) for c in cams])
    return imgs, cams, float(focal)


def load_into_testbed(tb, imgs, cams, focal, aabb_scale: int = 1) -> None:
    """the pyngp calls a user would make: create_empty_nerf_dataset + set_image / set_camera_* (python_api.cu:444-451, 809-853)"""
    n, h, w, _ = imgs.shape
    tb.create_empty_nerf_dataset(n, aabb_scale=aabb_scale)
    for i in range(n):
        tb.nerf.training.set_image(i, imgs[i])
        tb.nerf.training.set_camera_extrinsics(i, cams[i], convert_to_ngp=False)
        tb.nerf.training.set_camera_intrinsics(i, fx=focal, fy=focal, cx=0.5 * w, cy=0.5 * h)
    tb.nerf.training.n_images_for_training = n


BASE_CONFIG_L16F2 = {
    "loss": {"otype": "Huber"},
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20000, "decay_interval": 10000, "decay_base": 0.33,
                  "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
    "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
    "dir_encoding": {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4}, {"otype": "Identity"}]},
    "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}


def base_config(n_levels: int = 16, n_features_per_level: int = 2, log2_hashmap_size: int = 19) -> dict:
    """configs/nerf/base.json with the hash-grid shape overridden (the shipped file is L=8,F=4; BASELINE quotes L=16,F=2)."""
    import copy

    c = copy.deepcopy(BASE_CONFIG_L16F2)
    c["encoding"].update(n_levels=n_levels, n_features_per_level=n_features_per_level, log2_hashmap_size=log2_hashmap_size)
    return c
