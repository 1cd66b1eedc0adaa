"""ctypes wrapper of oracle/libngp_oracle.so (oracle/ngp_oracle.c) — TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
from __future__ import annotations

import ctypes as C
import importlib
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_lib = None


def _binding():
    return importlib.import_module("instant-ngp_b200.binding")  # struct layouts of the boundary header only


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        p = HERE / "libngp_oracle.so"
        if not p.exists():
            importlib.import_module("instant-ngp_b200.build").build_oracle()
        l = C.CDLL(str(p))
        B = _binding()
        f32, u32, u64, vp = C.c_float, C.c_uint32, C.c_uint64, C.c_void_p
        MC, TC, RC = C.POINTER(B.MarchConsts), C.POINTER(B.NerfTrainCfg), C.POINTER(B.RenderCfg)
        sig = {
            "orc_march_consts": (None, [f32, MC]),
            "orc_to_stepping_space": (f32, [f32, MC]),
            "orc_from_stepping_space": (f32, [f32, MC]),
            "orc_calc_dt": (f32, [f32, MC]),
            "orc_mip_from_dt": (u32, [f32, vp, u32]),
            "orc_cascaded_grid_idx_at": (u32, [vp, u32]),
            "orc_advance_to_next_voxel": (f32, [f32, MC, vp, vp, u32]),
            "orc_uv_to_ray": (None, [f32, f32, C.POINTER(B.TrainView), vp]),
            "orc_detmath_n": (None, [C.c_int, vp, vp, vp, u32]),
            "orc_srgb_to_linear": (f32, [f32]),
            "orc_linear_to_srgb": (f32, [f32]),
            "orc_linear_to_srgb_n": (None, [vp, vp, u32]),
            "orc_srgb_to_linear_n": (None, [vp, vp, u32]),
            "orc_half_to_float": (f32, [C.c_uint16]),
            "orc_float_to_half": (C.c_uint16, [f32]),
            "orc_morton3d": (u32, [u32, u32, u32]),
            "orc_ld_random_val": (f32, [u32, u32]),
            "orc_pcg32_seed": (None, [u64, u64, C.POINTER(u64), C.POINTER(u64)]),
            "orc_pcg32_next_uint": (u32, [C.POINTER(u64), u64]),
            "orc_pcg32_advance": (None, [C.POINTER(u64), u64, u64]),
            "orc_generate_training_samples": (u32, [u32, u32, u32, u64, u64, TC, vp, u32, vp, u32, C.POINTER(u32), vp, vp, vp, vp, vp]),
            "orc_compute_loss": (u32, [u32, u32, u64, u64, TC, vp, u32, vp, u32, vp, vp, vp, vp, vp, vp, vp, f32]),
            "orc_fill_rollover": (None, [u32, u32, vp, vp]),
            "orc_mark_untrained_density_grid": (None, [u32, vp, u32, vp, C.c_int]),
            "orc_generate_grid_samples": (None, [u32, u64, u64, u32, TC, vp, vp, vp, u32, f32]),
            "orc_splat_and_ema": (None, [u32, vp, vp, u32, u32, f32, vp, vp]),
            "orc_density_mean": (f32, [vp]),
            "orc_update_bitfield": (None, [u32, vp, f32, vp]),
            "orc_render_march": (None, [RC, C.c_int32, C.c_int32, vp, u32, vp, vp]),
            "orc_render_composite": (None, [RC, C.c_int32, C.c_int32, u32, vp, vp, vp, vp, vp, vp]),
            "orc_accumulate": (None, [u32, vp, vp, f32, u32]),
            "orc_tonemap": (None, [u32, C.POINTER(B.TonemapCfg), vp, vp]),
            "orc_version": (C.c_int, []),
        }
        for k, (r, a) in sig.items():
            fn = getattr(l, k)
            fn.restype, fn.argtypes = r, a
        _lib = l
    return _lib


def pcg32_seed(initstate: int, initseq: int = 1):
    s, i = C.c_uint64(), C.c_uint64()
    lib().orc_pcg32_seed(initstate, initseq, C.byref(s), C.byref(i))
    return s.value, i.value


def march_consts(cone_angle: float):
    m = _binding().MarchConsts()
    lib().orc_march_consts(cone_angle, C.byref(m))
    return m


def ptr(a: np.ndarray):
    return a.ctypes.data if a is not None else None


def generate_training_samples(n_rays, ray_offset, n_rays_global, rng, cfg, views_arr, n_views, bitfield, max_samples):
    """returns dict(n_kept, n_samples, per_ray_numsteps, ray_indices, rays, numsteps, coords)"""
    per_ray = np.zeros(n_rays, dtype=np.uint32)
    ray_indices = np.zeros(n_rays, dtype=np.uint32)
    rays = np.zeros((n_rays, 6), dtype=np.float32)
    numsteps = np.zeros((n_rays, 2), dtype=np.uint32)
    coords = np.zeros((max_samples, 7), dtype=np.float32)
    ns = C.c_uint32(0)
    kept = lib().orc_generate_training_samples(n_rays, ray_offset, n_rays_global, rng[0], rng[1], C.byref(cfg), C.addressof(views_arr), n_views, ptr(bitfield),
                                               max_samples, C.byref(ns), ptr(per_ray), ptr(ray_indices), ptr(rays), ptr(numsteps), ptr(coords))
    return dict(n_kept=kept, n_samples=ns.value, per_ray_numsteps=per_ray, ray_indices=ray_indices[:kept], rays=rays[:kept], numsteps=numsteps[:kept], coords=coords)
