"""ctypes binding of include/ngp_b200.h — structures and function prototypes, one to one."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libngp_b200.so"

NGP_MAX_LEVELS = 32


class NgpError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32),
        ("n_features_per_level", C.c_uint32),
        ("log2_hashmap_size", C.c_uint32),
        ("base_resolution", C.c_uint32),
        ("per_level_scale", C.c_float),
        ("n_params", C.c_uint32),
        ("offsets", C.c_uint32 * (NGP_MAX_LEVELS + 1)),
        ("resolutions", C.c_uint32 * NGP_MAX_LEVELS),
        ("scales", C.c_float * NGP_MAX_LEVELS),
    ]


class NerfDesc(C.Structure):
    _fields_ = [
        ("grid", GridDesc),
        ("n_hidden_density", C.c_uint32),
        ("n_hidden_rgb", C.c_uint32),
        ("density_mlp_offset", C.c_uint32),
        ("rgb_mlp_offset", C.c_uint32),
        ("grid_offset", C.c_uint32),
        ("n_mlp_params", C.c_uint32),
        ("n_params", C.c_uint32),
    ]


class FieldDesc(C.Structure):
    _fields_ = [
        ("grid", GridDesc),
        ("n_pos_dims", C.c_uint32),
        ("n_hidden", C.c_uint32),
        ("n_output_dims", C.c_uint32),
        ("mlp_offset", C.c_uint32),
        ("grid_offset", C.c_uint32),
        ("n_mlp_params", C.c_uint32),
        ("n_params", C.c_uint32),
    ]


class TonemapCfg(C.Structure):
    _fields_ = [
        ("exposure", C.c_float),
        ("background_color", C.c_float * 4),
        ("color_space", C.c_uint32),
        ("output_color_space", C.c_uint32),
        ("tonemap_curve", C.c_uint32),
        ("clamp_output_color", C.c_uint32),
        ("unmultiply_alpha", C.c_uint32),
    ]


class TrainView(C.Structure):
    _fields_ = [
        ("pixels", C.c_void_p),
        ("image_type", C.c_uint32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("focal_x", C.c_float),
        ("focal_y", C.c_float),
        ("principal_x", C.c_float),
        ("principal_y", C.c_float),
        ("lens_mode", C.c_uint32),
        ("lens_params", C.c_float * 4),
        ("xform", C.c_float * 12),
        ("no_mask", C.c_uint32),
    ]


class MarchConsts(C.Structure):
    _fields_ = [("cone_angle", C.c_float), ("log1p_c", C.c_float), ("a", C.c_float), ("b", C.c_float), ("at", C.c_float), ("bt", C.c_float)]


class NerfTrainCfg(C.Structure):
    _fields_ = [
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
        ("max_cascade", C.c_uint32),
        ("march", MarchConsts),
        ("snap_to_pixel_centers", C.c_uint32),
        ("random_bg_color", C.c_uint32),
        ("linear_colors", C.c_uint32),
        ("color_space", C.c_uint32),
        ("background_color", C.c_float * 3),
        ("loss_type", C.c_uint32),
        ("rgb_activation", C.c_uint32),
        ("density_activation", C.c_uint32),
        ("near_distance", C.c_float),
        ("loss_scale", C.c_float),
        ("train_mode", C.c_uint32),
        ("ray_stride", C.c_uint32),
        ("math_mode", C.c_uint32),
        ("gen_lanes_per_ray", C.c_uint32),
        ("gen_walk_empty", C.c_uint32),
        ("gen_speculation", C.c_uint32),
        ("compaction_order", C.c_uint32),
        ("cam_exposure", C.c_void_p),
        ("cam_exposure_gradient", C.c_void_p),
    ]


class NerfCounters(C.Structure):
    _fields_ = [("n_rays", C.c_uint32), ("n_samples", C.c_uint32), ("n_samples_compacted", C.c_uint32), ("pad", C.c_uint32)]


class AdamCfg(C.Structure):
    _fields_ = [
        ("learning_rate", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("epsilon", C.c_float),
        ("l2_reg", C.c_float),
        ("loss_scale", C.c_float),
        ("ema_decay", C.c_float),
        ("ema_step", C.c_uint32),
        ("optimize_matrix_params", C.c_uint32),
        ("optimize_non_matrix_params", C.c_uint32),
    ]


class RenderCfg(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("focal_x", C.c_float),
        ("focal_y", C.c_float),
        ("screen_x", C.c_float),
        ("screen_y", C.c_float),
        ("camera", C.c_float * 12),
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
        ("render_aabb_min", C.c_float * 3),
        ("render_aabb_max", C.c_float * 3),
        ("max_cascade", C.c_uint32),
        ("march", MarchConsts),
        ("rgb_activation", C.c_uint32),
        ("density_activation", C.c_uint32),
        ("min_transmittance", C.c_float),
        ("spp_index", C.c_uint32),
        ("near_distance", C.c_float),
        ("pixel_offset", C.c_float * 2),
        ("lens_mode", C.c_uint32),
        ("lens_params", C.c_float * 4),
        ("skips_per_tile", C.c_uint32),
        ("render_mode", C.c_uint32),
        ("depth_scale", C.c_float),
        ("math_mode", C.c_uint32),
    ]


u32, u64, i32, f32, vp, cp = C.c_uint32, C.c_uint64, C.c_int32, C.c_float, C.c_void_p, C.c_char_p
P = C.POINTER

# name -> (restype, argtypes); every symbol declared in include/ngp_b200.h
PROTOTYPES = {
    "ngp_last_error": (cp, []),
    "ngp_version": (C.c_int, []),
    "ngp_device_count": (C.c_int, []),
    "ngp_launch_count": (u64, []),
    "ngp_grid_desc_init": (C.c_int, [P(GridDesc), u32, u32, u32, u32, f32, u32]),
    "ngp_nerf_desc_init": (C.c_int, [P(NerfDesc), P(GridDesc), u32, u32]),
    "ngp_march_consts_init": (C.c_int, [P(MarchConsts), f32]),
    "ngp_nerf_init_params_host": (C.c_int, [P(NerfDesc), u64, vp]),
    "ngp_nerf_inference": (C.c_int, [P(NerfDesc), vp, u32, vp, vp, vp, u32]),
    "ngp_nerf_inference_rays": (C.c_int, [P(NerfDesc), vp, u32, vp, vp, vp, vp, vp, u32, vp, u32]),
    "ngp_nerf_density": (C.c_int, [P(NerfDesc), vp, u32, vp, u32, vp, vp]),
    "ngp_nerf_forward_backward": (C.c_int, [P(NerfDesc), vp, u32, vp, vp, vp, vp, vp]),
    "ngp_grid_encode": (C.c_int, [P(GridDesc), vp, u32, vp, u32, vp, vp]),
    "ngp_optimizer_step": (C.c_int, [P(NerfDesc), vp, P(AdamCfg), vp, vp, vp, vp, vp, vp, vp]),
    "ngp_nerf_generate_training_samples": (C.c_int, [vp, u32, u32, u32, u64, u64, P(NerfTrainCfg), vp, u32, vp, u32, vp, vp, vp, vp, vp]),
    "ngp_nerf_compute_loss": (C.c_int, [vp, u32, u32, u64, u64, P(NerfTrainCfg), vp, u32, vp, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "ngp_nerf_fill_rollover": (C.c_int, [vp, u32, vp, vp, vp]),
    "ngp_nerf_density_grid_scratch_bytes": (C.c_size_t, [u32]),
    "ngp_nerf_update_density_grid": (C.c_int, [P(NerfDesc), vp, P(NerfTrainCfg), vp, P(u64), u64, u32, u32, f32, vp, u32, vp, vp, vp, vp]),
    "ngp_nerf_update_bitfield": (C.c_int, [vp, u32, vp, vp, vp]),
    "ngp_nerf_render_scratch_bytes": (C.c_size_t, [i32, i32]),
    "ngp_nerf_render": (C.c_int, [P(NerfDesc), vp, P(RenderCfg), i32, i32, vp, vp, vp, vp, vp, vp]),
    "ngp_testbed_create": (vp, [C.c_int, vp]),
    "ngp_testbed_destroy": (None, [vp]),
    "ngp_testbed_create_empty_nerf_dataset": (C.c_int, [vp, u32, u32]),
    "ngp_testbed_set_image": (C.c_int, [vp, u32, vp, i32, i32]),
    "ngp_testbed_set_image_bytes": (C.c_int, [vp, u32, vp, i32, i32]),
    "ngp_render_pixel_offset": (None, [u32, P(f32)]),
    "ngp_testbed_set_camera_extrinsics": (C.c_int, [vp, u32, vp, C.c_int]),
    "ngp_testbed_set_camera_intrinsics": (C.c_int, [vp, u32, f32, f32, f32, f32, f32, f32, f32, f32]),
    "ngp_testbed_reload_network_from_json": (C.c_int, [vp, cp]),
    "ngp_testbed_reload_network_from_file": (C.c_int, [vp, cp]),
    "ngp_testbed_set_seed": (C.c_int, [vp, u64]),
    "ngp_testbed_reset": (C.c_int, [vp, C.c_int]),
    "ngp_testbed_get_view": (C.c_int, [vp, u32, P(TrainView)]),
    "ngp_testbed_set_option": (C.c_int, [vp, cp, C.c_double]),
    "ngp_testbed_get_camera_exposure": (C.c_int, [vp, u32, vp]),
    "ngp_testbed_set_camera_exposure": (C.c_int, [vp, u32, vp]),
    "ngp_testbed_get_option": (C.c_double, [vp, cp]),
    "ngp_testbed_train": (C.c_int, [vp, u32]),
    "ngp_set_scatter_aggregation": (None, [C.c_int]),
    "ngp_profile_mlp_phase": (C.c_int, [P(NerfDesc), vp, u32, vp, vp, vp, vp, vp]),
    "ngp_testbed_set_dp": (C.c_int, [vp, u32, u32]),
    "ngp_dp_unique_id_bytes": (C.c_size_t, []),
    "ngp_dp_unique_id": (C.c_int, [vp, C.c_size_t]),
    "ngp_testbed_init_dp": (C.c_int, [vp, u32, u32, vp, C.c_size_t]),
    "ngp_dp_rows": (None, [u32, u32, C.c_int32, P(C.c_int32), P(C.c_int32)]),
    "ngp_testbed_gather_rows": (C.c_int, [vp, C.c_int32, C.c_int32, vp, vp]),
    "ngp_testbed_train_compute_grads": (C.c_int, [vp, u32]),
    "ngp_testbed_train_front": (C.c_int, [vp, u32]),
    "ngp_testbed_train_back": (C.c_int, [vp]),
    "ngp_testbed_train_apply_grads": (C.c_int, [vp]),
    "ngp_testbed_grads": (vp, [vp]),
    "ngp_testbed_params": (vp, [vp]),
    "ngp_testbed_params_inference": (vp, [vp]),
    "ngp_testbed_params_fp32": (vp, [vp]),
    "ngp_testbed_dp_counters": (vp, [vp]),
    "ngp_testbed_n_params": (u32, [vp]),
    "ngp_testbed_training_step": (u32, [vp]),
    "ngp_testbed_loss": (f32, [vp]),
    "ngp_testbed_get_counters": (C.c_int, [vp, P(u32), P(u32), P(u32)]),
    "ngp_testbed_get_desc": (C.c_int, [vp, P(NerfDesc)]),
    "ngp_testbed_set_params_fp32": (C.c_int, [vp, vp, u32]),
    "ngp_testbed_get_params_fp16": (C.c_int, [vp, vp, u32, C.c_int]),
    "ngp_testbed_get_density_grid": (C.c_int, [vp, vp, u32, vp, u32]),
    "ngp_testbed_set_density_grid": (C.c_int, [vp, vp, u32]),
    "ngp_testbed_render": (C.c_int, [vp, i32, i32, vp, f32, f32, f32, f32, i32, i32, vp, vp, P(u32)]),
    "ngp_testbed_render_device": (C.c_int, [vp, i32, i32, vp, f32, f32, f32, f32, i32, i32, vp, vp]),
    "ngp_render_accumulate": (C.c_int, [vp, i32, i32, vp, vp, f32, u32]),
    "ngp_render_tonemap": (C.c_int, [vp, i32, i32, P(TonemapCfg), vp, vp]),
    "ngp_testbed_render_ex": (C.c_int, [vp, i32, i32, vp, f32, f32, f32, f32, u32, C.c_int, vp, vp]),
    "ngp_testbed_save_snapshot": (C.c_int, [vp, cp]),
    "ngp_testbed_save_snapshot_ex": (C.c_int, [vp, cp, C.c_int, C.c_int]),
    "ngp_json_to_msgpack": (C.c_int, [cp, C.c_int, vp, C.c_size_t, P(C.c_size_t)]),
    "ngp_load_network_config": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t, P(C.c_size_t)]),
    "ngp_msgpack_to_json": (C.c_int, [vp, C.c_size_t, C.c_int, vp, C.c_size_t, P(C.c_size_t)]),
    "ngp_testbed_load_snapshot": (C.c_int, [vp, cp]),
    "ngp_testbed_sync": (C.c_int, [vp]),
    "ngp_testbed_set_profiling": (C.c_int, [vp, C.c_int]),
    "ngp_testbed_get_phase_ms": (C.c_int, [vp, P(f32), P(u32)]),
    "ngp_testbed_update_image_async": (C.c_int, [vp, u32, vp]),
    # fields (image / SDF primitives)
    "ngp_grid_desc_init_nd": (C.c_int, [P(GridDesc), u32, u32, u32, u32, u32, f32]),
    "ngp_field_desc_init": (C.c_int, [P(FieldDesc), P(GridDesc), u32, u32, u32]),
    "ngp_field_init_params_host": (C.c_int, [P(FieldDesc), u64, vp]),
    "ngp_field_inference": (C.c_int, [P(FieldDesc), vp, u32, vp, vp, vp, u32]),
    "ngp_field_train_step": (C.c_int, [P(FieldDesc), vp, u32, vp, vp, u32, f32, vp, vp, vp, vp, vp]),
    "ngp_loss_evaluate": (C.c_int, [vp, u32, u32, u32, u32, f32, vp, vp, vp, vp]),
    "ngp_optimizer_step_flat": (C.c_int, [u32, u32, vp, P(AdamCfg), vp, vp, vp, vp, vp, vp, vp]),
    "ngp_image_generate_training_data": (C.c_int, [vp, u32, u64, u64, C.c_int, vp, u32, i32, i32, C.c_int, C.c_int, vp, vp]),
    "ngp_shuffle": (C.c_int, [vp, u32, u32, u32, vp, vp]),
    "ngp_module_create_network_with_input_encoding": (vp, [u32, u32, cp, cp]),
    "ngp_module_free": (None, [vp]),
    "ngp_module_n_input_dims": (u32, [vp]),
    "ngp_module_n_output_dims": (u32, [vp]),
    "ngp_module_n_params": (C.c_size_t, [vp]),
    "ngp_module_get_desc": (C.c_int, [vp, P(FieldDesc)]),
    "ngp_module_initialize_params": (C.c_int, [vp, C.c_size_t, vp, f32]),
    "ngp_module_inference": (C.c_int, [vp, vp, u32, vp, vp, vp]),
    "ngp_module_forward": (C.c_int, [vp, vp, u32, vp, vp, vp]),
    "ngp_module_backward": (C.c_int, [vp, vp, u32, vp, vp, vp, vp, vp, vp]),
    "ngp_field_testbed_create": (vp, [u32, C.c_int, vp]),
    "ngp_field_testbed_destroy": (None, [vp]),
    "ngp_field_testbed_set_image": (C.c_int, [vp, vp, i32, i32]),
    "ngp_field_testbed_set_sdf_training_data": (C.c_int, [vp, vp, vp, u32]),
    "ngp_field_testbed_reload_network_from_json": (C.c_int, [vp, cp]),
    "ngp_field_testbed_save_snapshot": (C.c_int, [vp, cp, C.c_int, C.c_int]),
    "ngp_field_testbed_load_snapshot": (C.c_int, [vp, cp]),
    "ngp_field_testbed_set_seed": (C.c_int, [vp, u64]),
    "ngp_field_testbed_set_option": (C.c_int, [vp, cp, C.c_double]),
    "ngp_field_testbed_train": (C.c_int, [vp, u32]),
    "ngp_field_testbed_loss": (f32, [vp]),
    "ngp_field_testbed_training_step": (u32, [vp]),
    "ngp_field_testbed_n_params": (C.c_size_t, [vp]),
    "ngp_field_testbed_get_desc": (C.c_int, [vp, P(FieldDesc)]),
    "ngp_field_testbed_params": (vp, [vp]),
    "ngp_field_testbed_params_inference": (vp, [vp]),
    "ngp_field_testbed_grads": (vp, [vp]),
    "ngp_field_testbed_set_params_fp32": (C.c_int, [vp, vp, C.c_size_t]),
    "ngp_field_testbed_get_params_fp16": (C.c_int, [vp, vp, C.c_size_t, C.c_int]),
    "ngp_field_testbed_evaluate": (C.c_int, [vp, vp, u32, vp]),
    "ngp_field_testbed_render_image": (C.c_int, [vp, i32, i32, vp]),
    "ngp_field_testbed_sync": (C.c_int, [vp]),
}

_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libngp_b200.so and attach prototypes.  Fails loudly if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise NgpError(f"{p} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    l = C.CDLL(str(p))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(l, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def lib() -> C.CDLL:
    return load_library()


def check(rc: int) -> None:
    if rc != 0:
        raise NgpError(lib().ngp_last_error().decode("utf-8", "replace"))
