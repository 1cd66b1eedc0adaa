"""Shared helpers for the parity tests (seeded inputs, descriptors, device buffers)."""
from __future__ import annotations

import ctypes as C
import importlib

import numpy as np

from oracle import net_oracle as O


def pkg():
    return importlib.import_module("instant-ngp_b200")


def make_desc(n_levels=16, F=2, log2_T=19, base_res=16, aabb_scale=4, nhd=1, nhr=2):
    """(C descriptor from the library, independent oracle layout) for the same config."""
    P = pkg()
    lib = P.load_library()
    g = P.GridDesc()
    assert lib.ngp_grid_desc_init(C.byref(g), n_levels, F, log2_T, base_res, 0.0, aabb_scale) == 0, lib.ngp_last_error()
    d = P.NerfDesc()
    assert lib.ngp_nerf_desc_init(C.byref(d), C.byref(g), nhd, nhr) == 0, lib.ngp_last_error()
    pls = O.per_level_scale_for(aabb_scale, base_res, n_levels)
    og = O.grid_layout(n_levels, F, log2_T, base_res, pls)
    return d, O.NerfLayout(og, nhd, nhr)


def random_params(L: O.NerfLayout, seed=0, grid_scale=1e-4, trained_like=False):
    """fp32 params in the flat reference layout: Xavier-uniform MLPs, small uniform hash grid (trainer.h:69-87)."""
    rng = np.random.default_rng(seed)
    parts = []
    for (r, c) in L.density_shapes + L.rgb_shapes:
        s = np.sqrt(6.0 / (r + c))
        parts.append(rng.uniform(-s, s, size=r * c))
    if trained_like:
        parts.append(np.clip(rng.normal(0.0, 0.1, size=L.grid.n_params), -1, 1))
    else:
        parts.append(rng.uniform(-grid_scale, grid_scale, size=L.grid.n_params))
    return np.concatenate(parts).astype(np.float32)


def random_coords(n, seed=1):
    rng = np.random.default_rng(seed)
    c = np.zeros((n, 7), dtype=np.float32)
    c[:, 0:3] = rng.uniform(0, 1, size=(n, 3))
    c[:, 3] = rng.uniform(0, 1, size=n)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    c[:, 4:7] = (d + 1) * 0.5
    return c


def rel_err(a, b, eps=1e-3):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / (np.abs(b) + eps)


# ---------------------------------------------------------------------------------------------------------------------
# march / loss helpers
# ---------------------------------------------------------------------------------------------------------------------
def sphere_bitfield(radius=0.25, max_cascade=0, full=False):
    """occupancy bitfield (128^3 * 8 / 8 bytes incl. all mips): solid sphere of the given radius around the cube centre
    in every cascade <= max_cascade (SURVEY §8d synthetic occupancy), or everything occupied."""
    idx = np.arange(128 ** 3, dtype=np.uint32)

    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xC30C30C3
        x = (x | (x >> 4)) & 0x0F00F00F
        x = (x | (x >> 8)) & 0xFF0000FF
        x = (x | (x >> 16)) & 0x0000FFFF
        return x

    X, Y, Z = compact(idx), compact(idx >> 1), compact(idx >> 2)
    bf = np.zeros(128 ** 3 * 8 // 8, dtype=np.uint8)
    for mip in range(max_cascade + 1):
        if full:
            occ = np.ones(128 ** 3, dtype=bool)
        else:
            s = np.float32(2.0 ** mip)
            p = [((c.astype(np.float32) + np.float32(0.5)) / np.float32(128.0) - np.float32(0.5)) * s + np.float32(0.5) for c in (X, Y, Z)]
            dist = np.sqrt((p[0] - 0.5) ** 2 + (p[1] - 0.5) ** 2 + (p[2] - 0.5) ** 2)
            occ = dist < radius
        bf[mip * 128 ** 3 // 8:(mip + 1) * 128 ** 3 // 8] = np.packbits(occ.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    return bf


def make_train_cfg(aabb_scale=1, **kw):
    P = pkg()
    lib = P.load_library()
    c = P.NerfTrainCfg()
    half = 0.5 * min(128, aabb_scale)
    for k in range(3):
        c.aabb_min[k] = 0.5 - half
        c.aabb_max[k] = 0.5 + half
    mc = 0
    while (1 << mc) < aabb_scale:
        mc += 1
    c.max_cascade = mc
    assert lib.ngp_march_consts_init(C.byref(c.march), 0.0 if aabb_scale <= 1 else 1.0 / 256.0) == 0
    c.snap_to_pixel_centers = 1
    c.random_bg_color = 1
    c.linear_colors = 0
    c.color_space = 0
    c.loss_type = 4  # Huber, as configs/nerf/base.json
    c.rgb_activation = 2
    c.density_activation = 3
    c.near_distance = 0.1
    c.loss_scale = 128.0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def make_views(imgs, cams, focal, lens=None, image_type=3):
    """(host view array whose pixel pointers are numpy buffers, list of buffers to keep alive)"""
    P = pkg()
    n, h, w, _ = imgs.shape
    arr = (P.TrainView * n)()
    keep = []
    for i in range(n):
        v = arr[i]
        buf = np.ascontiguousarray(imgs[i])
        keep.append(buf)
        v.pixels = buf.ctypes.data
        v.image_type = image_type
        v.width, v.height = w, h
        v.focal_x = v.focal_y = focal
        v.principal_x = v.principal_y = 0.5
        if lens is not None:
            v.lens_mode = 1
            for k in range(4):
                v.lens_params[k] = lens[k]
        m = np.asarray(cams[i], dtype=np.float32)
        for c in range(4):
            for r in range(3):
                v.xform[c * 3 + r] = m[r, c]
    return arr, keep


def views_to_device(arr, keep):
    """copy of the view array whose pixel pointers are CUDA buffers; returns (torch uint8 tensor holding the structs, keepalive)"""
    import torch

    P = pkg()
    n = len(arr)
    dev_arr = (P.TrainView * n)()
    C.memmove(dev_arr, arr, C.sizeof(arr))
    tens = []
    for i in range(n):
        t = torch.from_numpy(keep[i]).cuda()
        tens.append(t)
        dev_arr[i].pixels = t.data_ptr()
    raw = np.frombuffer(bytes(dev_arr), dtype=np.uint8).copy()
    t_views = torch.from_numpy(raw).cuda()
    return t_views, tens


# ---------------------------------------------------------------------------------------------------------------------
# field (image / SDF primitive) helpers
# ---------------------------------------------------------------------------------------------------------------------
def make_field_desc(n_pos_dims=2, n_levels=16, F=2, log2_T=19, base_res=16, per_level_scale=1.5, n_hidden=2, n_out=3):
    """(C descriptor from the library, independent oracle layout) of a NetworkWithInputEncoding."""
    from oracle import field_oracle as FO

    P = pkg()
    lib = P.load_library()
    g = P.GridDesc()
    assert lib.ngp_grid_desc_init_nd(C.byref(g), n_pos_dims, n_levels, F, log2_T, base_res, per_level_scale) == 0, lib.ngp_last_error()
    d = P.FieldDesc()
    assert lib.ngp_field_desc_init(C.byref(d), C.byref(g), n_pos_dims, n_hidden, n_out) == 0, lib.ngp_last_error()
    og = O.grid_layout(n_levels, F, log2_T, base_res, per_level_scale, n_pos_dims=n_pos_dims)
    return d, FO.FieldLayout(og, n_hidden, n_out)


def random_field_params(L, seed=0, trained_like=True):
    rng = np.random.default_rng(seed)
    parts = []
    for (r, c) in L.shapes:
        s = np.sqrt(6.0 / (r + c))
        parts.append(rng.uniform(-s, s, size=r * c))
    if trained_like:
        parts.append(np.clip(rng.normal(0.0, 0.1, size=L.grid.n_params), -1, 1))
    else:
        parts.append(rng.uniform(-1e-4, 1e-4, size=L.grid.n_params))
    return np.concatenate(parts).astype(np.float32)


def test_image(w=96, h=64, seed=0):
    """smooth + sharp synthetic RGBA float image, linear colour"""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    u, v = x / w, y / h
    img = np.zeros((h, w, 4), dtype=np.float32)
    img[..., 0] = 0.5 + 0.5 * np.sin(9 * u + 3 * v)
    img[..., 1] = (np.hypot(u - 0.5, v - 0.5) < 0.3).astype(np.float32) * 0.8 + 0.1
    img[..., 2] = u * v
    img[..., 3] = 1.0
    rng = np.random.default_rng(seed)
    img[..., :3] += rng.uniform(0, 0.02, size=(h, w, 3)).astype(np.float32)
    img[..., :3] = np.clip(img[..., :3], 0.0, 1.0)
    return img


def wsum64(b: bytes) -> str:
    """sum of byte[i] * (i + 1) mod 2^64, as oracle/ref/ref_snapshot_harness.cu prints it for binary values"""
    a = np.frombuffer(b, dtype=np.uint8).astype(np.uint64)
    with np.errstate(over="ignore"):
        return str(int((a * (np.arange(a.size, dtype=np.uint64) + np.uint64(1))).sum(dtype=np.uint64)))
