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
