"""`pyngp.Testbed`-compatible surface over libngp_b200.so for the NeRF path.

Mirrors src/python_api.cu:439-853 name for name where the hot path needs it: ``create_empty_nerf_dataset``,
``nerf.training.set_image / set_camera_extrinsics / set_camera_intrinsics / n_images_for_training``,
``reload_network_from_file / _from_json``, ``train``, ``render``, ``loss``, ``training_step``, ``n_params``,
``save_snapshot / load_snapshot``, and the camera a script sets before ``render(width, height, spp, linear)`` (``fov``,
``fov_axis``, ``set_nerf_camera_matrix``, ``set_camera_to_training_view``, ``screen_center``, ``zoom``, ``exposure``), ``reset``,
``load_file``.  Everything heavy happens inside the shared library; the camera state is host arithmetic (camera.py).
"""
from __future__ import annotations

import ctypes as C
import enum
import json
from pathlib import Path

import numpy as np

from . import binding as B
from .camera import CameraState


class TestbedMode(enum.IntEnum):
    Nerf = 0
    Sdf = 1
    Image = 2
    Volume = 3
    None_ = 4


class TrainMode(enum.IntEnum):  # ETrainMode (common.h:47-51)
    Nerf = 0
    Rfl = 1
    RflRelax = 2


class LossType(enum.IntEnum):  # ELossType
    L2 = 0
    L1 = 1
    Mape = 2
    Smape = 3
    Huber = 4
    LogL1 = 5
    RelativeL2 = 6


class NerfActivation(enum.IntEnum):  # ENerfActivation
    None_ = 0
    ReLU = 1
    Logistic = 2
    Exponential = 3


class RenderMode(enum.IntEnum):  # ERenderMode (common.h:68-79); rendered here: AO, Shade, Positions, Depth, Cost
    AO = 0
    Shade = 1
    Normals = 2
    Positions = 3
    Depth = 4
    Distortion = 5
    Cost = 6
    Slice = 7


class ColorSpace(enum.IntEnum):
    Linear = 0
    SRGB = 1


def _opt_property(name: str, cast=float):
    def getter(self):
        return cast(self._tb._get(name))

    def setter(self, value):
        self._tb._set(name, float(value))

    return property(getter, setter)


class _Training:
    """``testbed.nerf.training`` (python_api.cu:781-853)."""

    def __init__(self, tb: "Testbed"):
        self._tb = tb

    n_images_for_training = _opt_property("nerf.training.n_images_for_training", int)
    random_bg_color = _opt_property("nerf.training.random_bg_color", bool)
    linear_colors = _opt_property("nerf.training.linear_colors", bool)
    snap_to_pixel_centers = _opt_property("nerf.training.snap_to_pixel_centers", bool)
    near_distance = _opt_property("nerf.training.near_distance", float)
    density_grid_decay = _opt_property("nerf.training.density_grid_decay", float)
    optimize_exposure = _opt_property("nerf.training.optimize_exposure", bool)       # python_api.cu:791
    exposure_l2_reg = _opt_property("nerf.training.exposure_l2_reg", float)          # python_api.cu:806
    n_steps_between_cam_updates = _opt_property("nerf.training.n_steps_between_cam_updates", int)

    def camera_exposure(self, frame_idx: int) -> np.ndarray:
        """the image's exposure in stops per colour channel (Nerf::Training::cam_exposure[i].variable(), zero-mean over the training views)"""
        out = np.zeros(3, dtype=np.float32)
        B.check(B.lib().ngp_testbed_get_camera_exposure(self._tb._h, int(frame_idx), out.ctypes.data))
        return out

    def set_camera_exposure(self, frame_idx: int, rgb) -> None:
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(rgb, dtype=np.float32), (3,)))
        B.check(B.lib().ngp_testbed_set_camera_exposure(self._tb._h, int(frame_idx), v.ctypes.data))

    @property
    def loss_type(self) -> LossType:
        return LossType(int(self._tb._get("nerf.training.loss_type")))

    @loss_type.setter
    def loss_type(self, v) -> None:
        self._tb._set("nerf.training.loss_type", float(int(v)))

    @property
    def train_mode(self) -> TrainMode:
        """python_api.cu:781.  Default here is Nerf — what the reference runs without JIT fusion (testbed_nerf.cu:3091-3093);
        its JIT default is RflRelax."""
        return TrainMode(int(self._tb._get("nerf.training.train_mode")))

    @train_mode.setter
    def train_mode(self, v) -> None:
        self._tb._set("nerf.training.train_mode", float(int(v)))

    def set_image(self, frame_idx: int, img: np.ndarray, depth_img=None, depth_scale: float = 1.0) -> None:
        img = np.ascontiguousarray(img, dtype=np.float32)
        if img.ndim != 3 or img.shape[2] != 4:
            raise ValueError("image should be (H,W,C) where C=4")
        B.check(B.lib().ngp_testbed_set_image(self._tb._h, frame_idx, img.ctypes.data, img.shape[1], img.shape[0]))

    def set_image_bytes(self, frame_idx: int, img: np.ndarray) -> None:
        """NerfDataset::set_training_image(EImageDataType::Byte) (src/nerf_loader.cu:749-850): [H, W, 4] uint8, sRGB + straight alpha,
        as load_nerf keeps 8-bit files; converted to linear premultiplied colour on every read like the reference"""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 4:
            raise ValueError("image should be (H,W,4) uint8")
        B.check(B.lib().ngp_testbed_set_image_bytes(self._tb._h, frame_idx, img.ctypes.data, img.shape[1], img.shape[0]))

    def set_camera_extrinsics(self, frame_idx: int, camera_to_world: np.ndarray, convert_to_ngp: bool = True) -> None:
        m = np.ascontiguousarray(np.asarray(camera_to_world, dtype=np.float32)[:3, :4])
        B.check(B.lib().ngp_testbed_set_camera_extrinsics(self._tb._h, frame_idx, m.ctypes.data, int(convert_to_ngp)))

    def set_camera_intrinsics(self, frame_idx: int, fx: float = 0.0, fy: float = 0.0, cx: float = -0.5, cy: float = -0.5, k1: float = 0.0,
                              k2: float = 0.0, p1: float = 0.0, p2: float = 0.0) -> None:
        B.check(B.lib().ngp_testbed_set_camera_intrinsics(self._tb._h, frame_idx, fx, fy, cx, cy, k1, k2, p1, p2))


class _Nerf:
    """``testbed.nerf`` (python_api.cu:714-742)."""

    def __init__(self, tb: "Testbed"):
        self._tb = tb
        self.training = _Training(tb)

    cone_angle_constant = _opt_property("nerf.cone_angle_constant", float)
    render_min_transmittance = _opt_property("nerf.render_min_transmittance", float)

    @property
    def rgb_activation(self) -> NerfActivation:
        return NerfActivation(int(self._tb._get("nerf.rgb_activation")))

    @rgb_activation.setter
    def rgb_activation(self, v) -> None:
        self._tb._set("nerf.rgb_activation", float(int(v)))

    @property
    def density_activation(self) -> NerfActivation:
        return NerfActivation(int(self._tb._get("nerf.density_activation")))

    @density_activation.setter
    def density_activation(self, v) -> None:
        self._tb._set("nerf.density_activation", float(int(v)))


def load_network_config(path) -> dict:
    """Testbed::load_network_config for a .json file (src/testbed.cu:280-309): comments allowed, "parent" inheritance resolved
    (merge_parent_network_config :86-97).  Host only; the C library does the work (ngp_load_network_config)."""
    cap = 1 << 20
    out = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    B.check(B.lib().ngp_load_network_config(str(path).encode(), out, cap, C.byref(n)))
    return json.loads(out.value.decode())


def _default_stream(device: int) -> int:
    try:  # run on torch's current stream when torch drives the process (bench / DP); plain default stream otherwise
        import torch

        return torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
    except Exception:  # pragma: no cover
        return 0


class _ImageTraining:
    """``testbed.image.training`` (python_api.cu:755-765)."""

    def __init__(self, tb: "FieldTestbed"):
        self._tb = tb
        self._snap, self._linear = False, False

    @property
    def snap_to_pixel_centers(self) -> bool:
        return self._snap

    @snap_to_pixel_centers.setter
    def snap_to_pixel_centers(self, v) -> None:
        self._snap = bool(v)
        self._tb._set("image.training.snap_to_pixel_centers", float(self._snap))

    @property
    def linear_colors(self) -> bool:
        return self._linear

    @linear_colors.setter
    def linear_colors(self, v) -> None:
        self._linear = bool(v)
        self._tb._set("image.training.linear_colors", float(self._linear))


class _Image:
    def __init__(self, tb: "FieldTestbed"):
        self.training = _ImageTraining(tb)


class FieldTestbed:
    """``pyngp.Testbed(TestbedMode.Image)`` / ``(TestbedMode.Sdf)``: the image and SDF primitives (train_image,
    src/testbed_image.cu:231-302; train_sdf on caller-provided records, src/testbed_sdf.cu:1578-1619)."""

    def __init__(self, mode: TestbedMode, device: int = 0, stream: int | None = None):
        if mode not in (TestbedMode.Image, TestbedMode.Sdf):
            raise B.NgpError("FieldTestbed: mode must be Image or Sdf")
        self.mode = TestbedMode(mode)
        if stream is None:
            stream = _default_stream(device)
        self._h = B.lib().ngp_field_testbed_create(int(mode), device, C.c_void_p(stream))
        if not self._h:
            raise B.NgpError(B.lib().ngp_last_error().decode())
        self.image = _Image(self)
        self.training_batch_size = 1 << 18  # testbed.h:1089
        self.shall_train = True
        self.root_dir = ""

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                B.lib().ngp_field_testbed_destroy(h)
            except Exception:  # pragma: no cover
                pass

    def _set(self, name: str, value: float) -> None:
        B.check(B.lib().ngp_field_testbed_set_option(self._h, name.encode(), float(value)))

    # -- data ------------------------------------------------------------------------------------------------------
    def set_image(self, img: np.ndarray) -> None:
        """[H, W, 4] (or [H, W, 3]) float32 linear colour — what load_image leaves in m_image.data for a float image."""
        a = np.asarray(img, dtype=np.float32)
        if a.ndim != 3 or a.shape[2] not in (3, 4):
            raise B.NgpError("set_image: expected [H, W, 3|4]")
        if a.shape[2] == 3:
            a = np.concatenate([a, np.ones(a.shape[:2] + (1,), np.float32)], axis=2)
        a = np.ascontiguousarray(a)
        B.check(B.lib().ngp_field_testbed_set_image(self._h, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0]))
        self._image = a

    def load_training_data(self, path) -> None:
        """python_api.cu:452 for the image mode — Testbed::load_image (src/testbed_image.cu:393-437): .exr, .bin or an 8-bit image"""
        if self.mode != TestbedMode.Image:
            raise B.NgpError("load_training_data: the SDF mode trains on override_sdf_training_data (no mesh loader in this build)")
        from . import image_io

        try:
            img = image_io.load_image(path)
        except (OSError, ValueError) as e:
            raise B.NgpError(str(e)) from e
        self.set_image(img)
        self.data_path = str(path)

    def load_file(self, path) -> None:
        """python_api.cu:573: a snapshot (.ingp / .msgpack), a network config (.json with network / encoding / loss / optimizer keys) or the
        training image"""
        p = Path(path)
        if p.suffix.lower() in (".ingp", ".msgpack"):
            return self.load_snapshot(p)
        if p.suffix.lower() == ".json":
            return self.reload_network_from_file(p)
        self.load_training_data(p)

    def save_snapshot(self, path: str, include_optimizer_state: bool = False, compress: bool = True) -> None:
        """python_api.cu:563-570 for the image / SDF modes: the reference's .ingp / .msgpack container (network config + inference weights)"""
        B.check(B.lib().ngp_field_testbed_save_snapshot(self._h, str(path).encode(), int(include_optimizer_state), int(compress)))

    def load_snapshot(self, path: str) -> None:
        """python_api.cu:571: rebuilds the network from the config stored in the snapshot and takes its weights"""
        B.check(B.lib().ngp_field_testbed_load_snapshot(self._h, str(Path(path)).encode()))

    def override_sdf_training_data(self, points: np.ndarray, distances: np.ndarray) -> None:
        """python_api.cu:74-113.  Points are taken as unit-cube coordinates (no mesh is loaded, so there is no raw AABB to
        normalise by)."""
        p = np.ascontiguousarray(points, dtype=np.float32)
        d = np.ascontiguousarray(distances, dtype=np.float32)
        if p.ndim != 2 or d.ndim != 1 or p.shape[0] != d.shape[0] or p.shape[1] != 3:
            raise B.NgpError("Invalid Points<->Distances data")
        B.check(B.lib().ngp_field_testbed_set_sdf_training_data(self._h, p.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), p.shape[0]))

    # -- network ---------------------------------------------------------------------------------------------------
    def find_network_config(self, path) -> Path:
        """src/testbed.cu:254-270: as given, else <root_dir>/configs/<mode>/<path>"""
        p = Path(path)
        if p.exists() or p.is_absolute():
            return p
        cand = Path(self.root_dir) / "configs" / ("image" if self.mode == TestbedMode.Image else "sdf") / p
        return cand if cand.exists() else p

    def reload_network_from_file(self, path: str = "") -> None:
        self.reload_network_from_json(load_network_config(self.find_network_config(path)))

    def reload_network_from_json(self, config, config_base_path: str = "") -> None:
        text = config if isinstance(config, str) else json.dumps(config)
        B.check(B.lib().ngp_field_testbed_reload_network_from_json(self._h, text.encode()))

    def set_seed(self, seed: int) -> None:
        B.check(B.lib().ngp_field_testbed_set_seed(self._h, seed))

    @property
    def n_params(self) -> int:
        return int(B.lib().ngp_field_testbed_n_params(self._h))

    @property
    def training_step(self) -> int:
        return int(B.lib().ngp_field_testbed_training_step(self._h))

    @property
    def loss(self) -> float:
        return float(B.lib().ngp_field_testbed_loss(self._h))

    def desc(self) -> B.FieldDesc:
        d = B.FieldDesc()
        B.check(B.lib().ngp_field_testbed_get_desc(self._h, C.byref(d)))
        return d

    def train(self, batch_size: int | None = None) -> None:
        B.check(B.lib().ngp_field_testbed_train(self._h, int(batch_size or self.training_batch_size)))

    def frame(self) -> bool:
        if self.shall_train:
            self.train(self.training_batch_size)
        return True

    def params_ptr(self, inference: bool = False) -> int:
        f = B.lib().ngp_field_testbed_params_inference if inference else B.lib().ngp_field_testbed_params
        return int(f(self._h) or 0)

    def grads_ptr(self) -> int:
        return int(B.lib().ngp_field_testbed_grads(self._h) or 0)

    def set_params(self, params_fp32: np.ndarray) -> None:
        a = np.ascontiguousarray(params_fp32, dtype=np.float32)
        B.check(B.lib().ngp_field_testbed_set_params_fp32(self._h, a.ctypes.data_as(C.c_void_p), a.size))

    def get_params(self, inference: bool = False) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float16)
        B.check(B.lib().ngp_field_testbed_get_params_fp16(self._h, out.ctypes.data_as(C.c_void_p), out.size, int(inference)))
        return out

    def evaluate(self, positions: np.ndarray) -> np.ndarray:
        """network output (inference weights) at [n, D] positions -> [n, n_output_dims] float32"""
        d = self.desc()
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, d.n_pos_dims)
        out = np.empty((p.shape[0], d.n_output_dims), dtype=np.float32)
        B.check(B.lib().ngp_field_testbed_evaluate(self._h, p.ctypes.data_as(C.c_void_p), p.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def compute_image_mse(self, quantize: bool = False) -> float:
        """python_api.cu:659 / src/testbed_image.cu:490-560: the network (inference weights) at every pixel centre against the image
        itself — snapped read, sRGB-encoded unless `image.training.linear_colors` — mean of |diff|^2 / 3; `quantize` rounds the
        prediction to 8 bits first.  The reduction runs on the host in float64 (the reference sums floats on the device)."""
        img = getattr(self, "_image", None)
        if self.mode != TestbedMode.Image or img is None:
            raise B.NgpError("compute_image_mse: image mode with an image set")
        h, w = img.shape[:2]
        xs = (np.arange(w, dtype=np.float32) + np.float32(0.5)) / np.float32(w)
        ys = (np.arange(h, dtype=np.float32) + np.float32(0.5)) / np.float32(h)
        pos = np.stack(np.meshgrid(xs, ys, indexing="xy"), axis=-1).reshape(-1, 2)
        pred = self.evaluate(pos)[:, :3].astype(np.float32)
        target = img[..., :3].reshape(-1, 3).astype(np.float32)
        if not self.image.training.linear_colors:
            target = np.where(target < 0.0031308, 12.92 * target, 1.055 * np.power(np.maximum(target, 0.0), 0.41666) - 0.055).astype(np.float32)
        if quantize:
            pred = np.clip((pred * np.float32(255.0) + np.float32(0.5)).astype(np.int32), 0, 255).astype(np.float32) / np.float32(255.0)
        d = (target - pred).astype(np.float64)
        return float(np.mean(np.sum(d * d, axis=1) / 3.0))

    def render(self, width: int, height: int, spp: int = 1, linear: bool = True) -> np.ndarray:
        """Image mode: the learned image at every pixel centre, float32 [H, W, 4] (render_image, default full-frame view)."""
        out = np.empty((height, width, 4), dtype=np.float32)
        B.check(B.lib().ngp_field_testbed_render_image(self._h, width, height, out.ctypes.data_as(C.c_void_p)))
        return out

    def sync(self) -> None:
        B.check(B.lib().ngp_field_testbed_sync(self._h))


class Testbed:
    """Drop-in for ``pyngp.Testbed``: the NeRF path; ``Testbed(TestbedMode.Image | TestbedMode.Sdf)`` gives a FieldTestbed."""

    def __new__(cls, mode: TestbedMode = TestbedMode.Nerf, device: int = 0, stream: int | None = None):
        if cls is Testbed and mode in (TestbedMode.Image, TestbedMode.Sdf):
            return FieldTestbed(mode, device, stream)
        return super().__new__(cls)

    def __init__(self, mode: TestbedMode = TestbedMode.Nerf, device: int = 0, stream: int | None = None):
        if mode != TestbedMode.Nerf:
            raise B.NgpError("ngp_b200 implements the NeRF, Image and Sdf modes only")
        if stream is None:
            stream = _default_stream(device)
        self._h = B.lib().ngp_testbed_create(device, C.c_void_p(stream))
        if not self._h:
            raise B.NgpError(B.lib().ngp_last_error().decode())
        self.nerf = _Nerf(self)
        self.training_batch_size = 1 << 18
        self.root_dir = ""
        self.jit_fusion = True        # accepted for script compatibility: every kernel here is the fused form
        self._camera = CameraState()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                B.lib().ngp_testbed_destroy(h)
            except Exception:  # pragma: no cover
                pass

    # -- options ---------------------------------------------------------------------------------------------------
    def _set(self, name: str, value: float) -> None:
        B.check(B.lib().ngp_testbed_set_option(self._h, name.encode(), value))

    def _get(self, name: str) -> float:
        v = B.lib().ngp_testbed_get_option(self._h, name.encode())
        if v != v:
            raise B.NgpError(B.lib().ngp_last_error().decode())
        return v

    shall_train = _opt_property("shall_train", bool)

    @property
    def _tb(self):
        return self

    @property
    def color_space(self) -> ColorSpace:
        return ColorSpace(int(self._get("color_space")))

    @color_space.setter
    def color_space(self, v) -> None:
        self._set("color_space", float(int(v)))

    exposure = _opt_property("exposure", float)

    @property
    def snap_to_pixel_centers(self) -> bool:
        """m_snap_to_pixel_centers (python_api.cu:699): False (the reference's default) jitters the pixel position per sample index
        (ld_random_pixel_offset), True sends every sample through the pixel centre (what scripts/run.py sets for evaluation)"""
        return bool(self._get("snap_to_pixel_centers"))

    @snap_to_pixel_centers.setter
    def snap_to_pixel_centers(self, v) -> None:
        self._set("snap_to_pixel_centers", float(bool(v)))

    @property
    def render_with_lens_distortion(self) -> bool:
        """m_render_with_lens_distortion (python_api.cu:690): apply `render_lens` (set by set_camera_to_training_view) to the render rays"""
        return bool(self._get("render_with_lens_distortion"))

    @render_with_lens_distortion.setter
    def render_with_lens_distortion(self, v) -> None:
        self._set("render_with_lens_distortion", float(bool(v)))

    @property
    def render_lens(self) -> dict:
        return dict(mode=int(self._get("render_lens.mode")), params=[self._get(f"render_lens.params.{k}") for k in range(4)])

    @render_lens.setter
    def render_lens(self, lens) -> None:
        self._set("render_lens.mode", float(int(lens["mode"])))
        for k in range(4):
            self._set(f"render_lens.params.{k}", float(lens["params"][k]))

    @property
    def render_mode(self) -> RenderMode:
        """m_render_mode (python_api.cu:683): Shade, AO, Positions, Depth and Cost are rendered; the others are refused by the library"""
        return RenderMode(int(self._get("render_mode")))

    @render_mode.setter
    def render_mode(self, v) -> None:
        self._set("render_mode", float(int(RenderMode(int(v)))))

    # -- camera (src/testbed.cu:440, 486-528, 4081-4087, 4649-4657) -------------------------------------------------------
    @property
    def fov(self) -> float:
        return self._camera.fov()

    @fov.setter
    def fov(self, degrees: float) -> None:
        self._camera.set_fov(float(degrees))

    @property
    def fov_xy(self):
        return self._camera.fov_xy()

    @fov_xy.setter
    def fov_xy(self, degrees_xy) -> None:
        self._camera.set_fov_xy(degrees_xy)

    @property
    def fov_axis(self) -> int:
        return self._camera.fov_axis

    @fov_axis.setter
    def fov_axis(self, axis: int) -> None:
        if int(axis) not in (0, 1):
            raise ValueError("fov_axis must be 0 or 1")
        self._camera.fov_axis = int(axis)

    @property
    def relative_focal_length(self):
        return self._camera.relative_focal_length

    @relative_focal_length.setter
    def relative_focal_length(self, v) -> None:
        self._camera.relative_focal_length = (float(v[0]), float(v[1]))

    @property
    def screen_center(self):
        return self._camera.screen_center

    @screen_center.setter
    def screen_center(self, v) -> None:
        self._camera.screen_center = (float(v[0]), float(v[1]))

    @property
    def zoom(self) -> float:
        return self._camera.zoom

    @zoom.setter
    def zoom(self, v: float) -> None:
        self._camera.zoom = float(v)

    @property
    def camera_matrix(self) -> np.ndarray:
        """3x4 camera-to-world in ngp convention (m_camera)"""
        return self._camera.matrix.copy()

    @camera_matrix.setter
    def camera_matrix(self, m) -> None:
        self._camera.matrix = np.ascontiguousarray(np.asarray(m, dtype=np.float32)[:3, :4]).copy()

    def reset_camera(self) -> None:
        self._camera.reset()

    def set_nerf_camera_matrix(self, cam) -> None:
        """python_api.cu:653: a camera-to-world matrix in the original NeRF convention, through the dataset's scale / offset"""
        from .nerf_loader import nerf_matrix_to_ngp

        off = [self._get(f"nerf.training.dataset.offset.{ax}") for ax in "xyz"]
        self._camera.matrix = nerf_matrix_to_ngp(np.asarray(cam, dtype=np.float32), self._get("nerf.training.dataset.scale"), off)

    def training_view(self, idx: int) -> dict:
        """intrinsics and transform of one training view as the Testbed holds them"""
        v = B.TrainView()
        B.check(B.lib().ngp_testbed_get_view(self._h, int(idx), C.byref(v)))
        xform = np.array(list(v.xform), dtype=np.float32).reshape(4, 3).T.copy()     # stored column major [c * 3 + r]
        return dict(resolution=(v.width, v.height), focal_length=(v.focal_x, v.focal_y), principal_point=(v.principal_x, v.principal_y),
                    lens_mode=int(v.lens_mode), lens_params=[float(x) for x in v.lens_params], xform=xform)

    def set_camera_to_training_view(self, trainview: int) -> None:
        """python_api.cu:654 / src/testbed.cu:486-505: camera, focal length, screen centre AND lens of the view
        (m_render_with_lens_distortion = true, m_render_lens = the view's lens)"""
        v = self.training_view(trainview)
        self._camera.to_training_view(v["xform"], v["focal_length"], v["resolution"], v["principal_point"])
        self.render_lens = dict(mode=v["lens_mode"], params=v["lens_params"])
        self.render_with_lens_distortion = True

    @property
    def background_color(self):
        return [self._get(f"background_color.{k}") for k in "rgba"]

    @background_color.setter
    def background_color(self, rgba) -> None:
        self._set("background_color.r", float(rgba[0]))
        self._set("background_color.g", float(rgba[1]))
        self._set("background_color.b", float(rgba[2]))
        if len(rgba) > 3:
            self._set("background_color.a", float(rgba[3]))

    # -- data ------------------------------------------------------------------------------------------------------
    def create_empty_nerf_dataset(self, n_images: int, aabb_scale: int = 1, is_hdr: bool = False) -> None:
        if is_hdr:
            raise B.NgpError("is_hdr datasets are not supported")
        B.check(B.lib().ngp_testbed_create_empty_nerf_dataset(self._h, n_images, aabb_scale))

    def load_training_data(self, path: str) -> None:
        """python_api.cu:452 — a transforms.json (or a directory of them) with 8-bit images (nerf_loader.py)"""
        from . import nerf_loader

        self.dataset = nerf_loader.load_into_testbed(self, path)

    # -- network ---------------------------------------------------------------------------------------------------
    def reload_network_from_file(self, path: str = "") -> None:
        """python_api.cu:543: a config file, looked up under <root_dir>/configs/nerf/ when not found as given; "parent" chains resolve
        inside the library (ngp_testbed_reload_network_from_file)"""
        B.check(B.lib().ngp_testbed_reload_network_from_file(self._h, str(self.find_network_config(path)).encode()))

    def reload_network_from_json(self, config, config_base_path: str = "") -> None:
        text = config if isinstance(config, str) else json.dumps(config)
        B.check(B.lib().ngp_testbed_reload_network_from_json(self._h, text.encode()))

    def set_seed(self, seed: int) -> None:
        B.check(B.lib().ngp_testbed_set_seed(self._h, seed))

    def reset(self, reset_density_grid: bool = True) -> None:
        """python_api.cu:534 — Testbed::reset_network: parameters re-initialised from the seed, optimizer / counters / step cleared"""
        B.check(B.lib().ngp_testbed_reset(self._h, int(reset_density_grid)))

    def find_network_config(self, path) -> Path:
        """src/testbed.cu:254-270: as given, else <root_dir>/configs/nerf/<path>"""
        p = Path(path)
        if p.exists() or p.is_absolute():
            return p
        cand = Path(self.root_dir) / "configs" / "nerf" / p
        return cand if cand.exists() else p

    def load_file(self, path) -> None:
        """python_api.cu:573 / src/testbed.cu:353-410: snapshot, network config or training data, told apart like the reference does"""
        p = Path(path)
        if not p.exists():
            if p.suffix.lower() == ".json" and self.find_network_config(p).exists():
                return self.reload_network_from_file(self.find_network_config(p))
            raise B.NgpError(f"File '{p}' does not exist.")
        if p.suffix.lower() in (".ingp", ".msgpack", ".ngpb"):
            return self.load_snapshot(p)
        if p.suffix.lower() == ".json":
            from .nerf_loader import _strip_json_comments

            doc = json.loads(_strip_json_comments(p.read_text()))
            if "snapshot" in doc:
                raise B.NgpError("snapshots in JSON text are not supported: use .ingp / .msgpack")
            if any(k in doc for k in ("parent", "network", "encoding", "loss", "optimizer")):
                return self.reload_network_from_file(p)
            if "path" in doc:
                raise B.NgpError("camera paths are not implemented")
        self.load_training_data(p)
        self.shall_train = True

    @property
    def n_params(self) -> int:
        return int(B.lib().ngp_testbed_n_params(self._h))

    @property
    def training_step(self) -> int:
        return int(B.lib().ngp_testbed_training_step(self._h))

    @property
    def loss(self) -> float:
        return float(B.lib().ngp_testbed_loss(self._h))

    def desc(self) -> B.NerfDesc:
        d = B.NerfDesc()
        B.check(B.lib().ngp_testbed_get_desc(self._h, C.byref(d)))
        return d

    def counters(self) -> dict:
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        B.lib().ngp_testbed_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"rays_per_batch": a.value, "measured_batch_size": b.value, "measured_batch_size_before_compaction": c.value}

    # -- training --------------------------------------------------------------------------------------------------
    def train(self, batch_size: int | None = None) -> None:
        B.check(B.lib().ngp_testbed_train(self._h, int(batch_size or self.training_batch_size)))

    def frame(self) -> bool:
        if self.shall_train:
            self.train(self.training_batch_size)
        return True

    def set_data_parallel(self, rank: int, world: int) -> None:
        """ray partition only: the caller runs the collectives (train_front / train_back / train_apply_grads)"""
        B.check(B.lib().ngp_testbed_set_dp(self._h, rank, world))

    @staticmethod
    def dp_unique_id() -> bytes:
        """rank 0: the id every rank passes to init_data_parallel (hand it over with torch.distributed / MPI / a file)"""
        n = B.lib().ngp_dp_unique_id_bytes()
        buf = (C.c_uint8 * n)()
        B.check(B.lib().ngp_dp_unique_id(buf, n))
        return bytes(buf)

    def init_data_parallel(self, rank: int, world: int, unique_id: bytes) -> None:
        """one process per GPU: from here on train() is the whole data-parallel step (NCCL all-reduce of counters and gradients inside
        the library, ngp_testbed_init_dp)"""
        buf = (C.c_uint8 * len(unique_id)).from_buffer_copy(unique_id)
        B.check(B.lib().ngp_testbed_init_dp(self._h, rank, world, buf, len(unique_id)))
        self._dp = (rank, world)

    def dp_rows(self, height: int, rank: int | None = None):
        r, w = getattr(self, "_dp", (0, 1))
        y0, y1 = C.c_int32(), C.c_int32()
        B.lib().ngp_dp_rows(r if rank is None else rank, w, height, C.byref(y0), C.byref(y1))
        return y0.value, y1.value

    def render_device_sharded(self, width: int, height: int, camera_matrix, focal_length, rgba_ptr: int, depth_ptr: int, screen_center=(0.5, 0.5)) -> None:
        """this rank's row tile of the frame, then the exchange of tiles: every rank ends up with the whole frame in its device buffers"""
        y0, y1 = self.dp_rows(height)
        if y1 > y0:
            self.render_device(width, height, camera_matrix, focal_length, rgba_ptr, depth_ptr, screen_center, rows=(y0, y1))
        B.check(B.lib().ngp_testbed_gather_rows(self._h, width, height, C.c_void_p(rgba_ptr), C.c_void_p(depth_ptr)))

    def train_compute_grads(self, batch_size: int | None = None) -> None:
        B.check(B.lib().ngp_testbed_train_compute_grads(self._h, int(batch_size or self.training_batch_size)))

    def train_front(self, batch_size: int | None = None) -> None:
        B.check(B.lib().ngp_testbed_train_front(self._h, int(batch_size or self.training_batch_size)))

    def train_back(self) -> None:
        B.check(B.lib().ngp_testbed_train_back(self._h))

    def train_apply_grads(self) -> None:
        B.check(B.lib().ngp_testbed_train_apply_grads(self._h))

    def grads_ptr(self) -> int:
        return int(B.lib().ngp_testbed_grads(self._h) or 0)

    def dp_counters_ptr(self) -> int:
        return int(B.lib().ngp_testbed_dp_counters(self._h) or 0)

    def params_ptr(self, inference: bool = False) -> int:
        f = B.lib().ngp_testbed_params_inference if inference else B.lib().ngp_testbed_params
        return int(f(self._h) or 0)

    # -- params / grid exchange ------------------------------------------------------------------------------------
    def set_params(self, params_fp32: np.ndarray) -> None:
        p = np.ascontiguousarray(params_fp32, dtype=np.float32)
        B.check(B.lib().ngp_testbed_set_params_fp32(self._h, p.ctypes.data, p.size))

    def get_params(self, inference: bool = False) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float16)
        B.check(B.lib().ngp_testbed_get_params_fp16(self._h, out.ctypes.data, out.size, int(inference)))
        return out

    def get_density_grid(self):
        n_casc = int(self._get("nerf.max_cascade")) + 1
        grid = np.empty(128 ** 3 * n_casc, dtype=np.float32)
        bits = np.empty(128 ** 3 * 8 // 8, dtype=np.uint8)
        B.check(B.lib().ngp_testbed_get_density_grid(self._h, grid.ctypes.data, grid.size, bits.ctypes.data, bits.size))
        return grid, bits

    def set_density_grid(self, grid: np.ndarray) -> None:
        g = np.ascontiguousarray(grid, dtype=np.float32)
        B.check(B.lib().ngp_testbed_set_density_grid(self._h, g.ctypes.data, g.size))

    # -- render ----------------------------------------------------------------------------------------------------
    def render(self, width: int = 1920, height: int = 1080, *args, **kw):
        """Two call forms.  ``render(width, height, spp=1, linear=True)`` is the reference's (python_api.cu:507-519): the frame of the
        Testbed's own camera (``set_nerf_camera_matrix`` / ``set_camera_to_training_view`` / ``camera_matrix``, ``fov``, ``zoom``,
        ``screen_center``).  ``render(width, height, camera_matrix, focal_length, screen_center=(0.5, 0.5), spp=1, linear=True,
        rows=None, return_depth=False)`` takes the camera explicitly (ngp convention, focal length in pixels).  Both return
        float32 [H, W, 4] premultiplied RGBA, linear — or sRGB after exposure / tonemap for ``linear=False``."""
        explicit = "camera_matrix" in kw or (len(args) >= 1 and not isinstance(args[0], (int, np.integer, bool)))
        if explicit:
            return self._render_explicit(width, height, *args, **kw)
        names = ("spp", "linear", "start_t", "end_t", "fps", "shutter_fraction")
        opts = dict(zip(names, args))
        for k, v in kw.items():
            if k in opts:
                raise TypeError(f"render() got multiple values for argument '{k}'")
            if k not in names + ("return_depth",):
                raise TypeError(f"render() got an unexpected keyword argument '{k}'")
            opts[k] = v
        if opts.get("start_t", -1.0) >= 0 or opts.get("end_t", -1.0) >= 0:
            raise B.NgpError("render: camera paths (start_t / end_t) are not implemented")
        cam, focal, center = self._camera.render_args(width, height)
        # Testbed::render_to_cpu hands render_frame an EMPTY lens (python_api.cu:199-213, the `{}, // lens` argument): the Python
        # render() of the reference never applies m_render_lens, whatever render_with_lens_distortion says (only the windowed frame()
        # path does, src/testbed.cu:3217).  Measured on the B200 (profiles/r2a): reference render() of fox's test views = 26.0 dB
        # against the distorted photographs, 29.4 dB once the lens is applied.  The reference-shaped call reproduces the reference;
        # the explicit-camera form below honours render_with_lens_distortion / render_lens.
        with_lens = self._h is not None and self.render_with_lens_distortion
        if with_lens:
            self.render_with_lens_distortion = False
        try:
            return self._render_explicit(width, height, cam, focal, center, spp=int(opts.get("spp", 1)), linear=bool(opts.get("linear", True)),
                                         return_depth=bool(opts.get("return_depth", False)))
        finally:
            if with_lens:
                self.render_with_lens_distortion = True

    def render_with_depth(self, width: int = 1920, height: int = 1080, spp: int = 1, linear: bool = True):
        """python_api.cu:520-532: (rgba, depth)"""
        return self.render(width, height, spp, linear, return_depth=True)

    def _render_explicit(self, width: int, height: int, camera_matrix: np.ndarray, focal_length, screen_center=(0.5, 0.5), spp: int = 1,
                         linear: bool = True, rows=None, return_depth: bool = False):
        """≙ Testbed.render(width, height, spp, linear) with an explicit ngp-convention 3x4 camera-to-world matrix
        (the reference takes it from testbed.set_nerf_camera_matrix).  Returns float32 [H, W, 4] linear premultiplied RGBA."""
        cam = np.ascontiguousarray(np.asarray(camera_matrix, dtype=np.float32)[:3, :4])
        fx, fy = (focal_length, focal_length) if np.isscalar(focal_length) else focal_length
        y0, y1 = rows if rows is not None else (0, height)
        # a full frame is written completely by the library: no need to zero 41 MB first (a 1920x1080 frame: ~8 ms of page faults)
        alloc = np.zeros if rows is not None else np.empty
        rgba = alloc((height, width, 4), dtype=np.float32)
        depth = alloc((height, width), dtype=np.float32)
        if spp != 1 or not linear:
            # the accumulate + tonemap epilogue (render_buffer.cu): spp frames averaged, sRGB output for linear=False
            if rows is not None:
                raise B.NgpError("render: row tiles are rendered with spp=1, linear=True (accumulate per tile on the caller's side)")
            B.check(B.lib().ngp_testbed_render_ex(self._h, width, height, cam.ctypes.data, fx, fy, screen_center[0], screen_center[1], int(spp), int(linear),
                                                  rgba.ctypes.data, depth.ctypes.data))
            return (rgba, depth) if return_depth else rgba
        steps = C.c_uint32(0)
        B.check(B.lib().ngp_testbed_render(self._h, width, height, cam.ctypes.data, fx, fy, screen_center[0], screen_center[1], y0, y1, rgba.ctypes.data,
                                           depth.ctypes.data, C.byref(steps)))
        self.last_render_steps = steps.value
        return (rgba, depth) if return_depth else rgba

    def render_device(self, width: int, height: int, camera_matrix: np.ndarray, focal_length, rgba_ptr: int, depth_ptr: int, screen_center=(0.5, 0.5),
                      rows=None) -> None:
        cam = np.ascontiguousarray(np.asarray(camera_matrix, dtype=np.float32)[:3, :4])
        fx, fy = (focal_length, focal_length) if np.isscalar(focal_length) else focal_length
        y0, y1 = rows if rows is not None else (0, height)
        B.check(B.lib().ngp_testbed_render_device(self._h, width, height, cam.ctypes.data, fx, fy, screen_center[0], screen_center[1], y0, y1,
                                                  C.c_void_p(rgba_ptr), C.c_void_p(depth_ptr)))

    # -- snapshots -------------------------------------------------------------------------------------------------
    def save_snapshot(self, path: str, include_optimizer_state: bool = False, compress: bool = True) -> None:
        p = str(path)
        if p.lower().endswith((".ingp", ".msgpack")):
            B.check(B.lib().ngp_testbed_save_snapshot_ex(self._h, p.encode(), int(include_optimizer_state), int(compress)))
        else:
            B.check(B.lib().ngp_testbed_save_snapshot(self._h, p.encode()))

    def load_snapshot(self, path: str) -> None:
        B.check(B.lib().ngp_testbed_load_snapshot(self._h, str(Path(path)).encode()))

    # -- profiling / streaming data --------------------------------------------------------------------------------------
    PHASES = ("occupancy_grid", "sample_generation", "inference", "loss_compaction", "forward_backward", "optimizer", "allreduce")

    def set_profiling(self, enable: bool) -> None:
        B.check(B.lib().ngp_testbed_set_profiling(self._h, int(enable)))

    def phase_ms(self) -> dict:
        ms = (C.c_float * len(self.PHASES))()
        n = C.c_uint32(0)
        B.check(B.lib().ngp_testbed_get_phase_ms(self._h, ms, C.byref(n)))
        return {"steps": n.value, **{k: float(ms[i]) for i, k in enumerate(self.PHASES)}}

    def update_image_async(self, frame_idx: int, host_ptr: int) -> None:
        B.check(B.lib().ngp_testbed_update_image_async(self._h, frame_idx, C.c_void_p(host_ptr)))

    def sync(self) -> None:
        B.check(B.lib().ngp_testbed_sync(self._h))
