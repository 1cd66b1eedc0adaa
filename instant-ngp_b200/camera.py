"""Host-side camera state of the Testbed (the part of ``pyngp.Testbed`` a script touches before ``render``): pure functions,
no device.  Follows src/testbed.cu:4081-4087 (fov <-> relative focal length), :4649-4657 (calc_focal_length,
render_screen_center), :486-505 (set_camera_to_training_view), :507-540 (reset_camera) and common.h's fov helpers."""
from __future__ import annotations

import math

import numpy as np


def fov_to_focal_length(resolution: float, degrees: float) -> float:
    """common_host.h / nerf_loader.cu: 0.5 * resolution / tan(0.5 * fov)"""
    return 0.5 * float(resolution) / math.tan(0.5 * float(degrees) * math.pi / 180.0)


def focal_length_to_fov(resolution: float, focal_length: float) -> float:
    """2 * atan(resolution / (2 * focal_length)), in degrees"""
    return 2.0 * 180.0 / math.pi * math.atan(float(resolution) / (float(focal_length) * 2.0))


def calc_focal_length(resolution, relative_focal_length, fov_axis: int, zoom: float):
    """Testbed::calc_focal_length (src/testbed.cu:4649-4651): pixels, both axes from the resolution along `fov_axis`"""
    r = float(resolution[fov_axis]) * float(zoom)
    return (float(relative_focal_length[0]) * r, float(relative_focal_length[1]) * r)


def render_screen_center(screen_center, zoom: float):
    """Testbed::render_screen_center (src/testbed.cu:4653-4657): the point of the image plane the optical axis goes through"""
    return ((0.5 - float(screen_center[0])) * float(zoom) + 0.5, (0.5 - float(screen_center[1])) * float(zoom) + 0.5)


def default_camera(scale: float = 1.5) -> np.ndarray:
    """Testbed::reset_camera (src/testbed.cu:507-540) for the NeRF mode: at (0.5, 0.5, 0.5) - scale * view_dir, looking down -z
    with y flipped; 3x4 row-major [right | up | forward | origin] in ngp convention"""
    m = np.array([[1.0, 0.0, 0.0, 0.5], [0.0, -1.0, 0.0, 0.5], [0.0, 0.0, -1.0, 0.5]], dtype=np.float32)
    m[:, 3] -= np.float32(scale) * m[:, 2]
    return m


class CameraState:
    """m_camera, m_fov_axis, m_relative_focal_length, m_zoom, m_screen_center of the reference Testbed"""

    def __init__(self):
        self.reset()

    def reset(self) -> None:
        self.fov_axis = 1
        self.zoom = 1.0
        self.screen_center = (0.5, 0.5)
        self.relative_focal_length = (1.0, 1.0)
        self.scale = 1.5
        self.matrix = default_camera(self.scale)
        self.set_fov(50.625)        # reset_camera: set_fov(50.625f)

    def fov(self) -> float:
        return focal_length_to_fov(1.0, self.relative_focal_length[self.fov_axis])

    def set_fov(self, degrees: float) -> None:
        f = fov_to_focal_length(1.0, degrees)
        self.relative_focal_length = (f, f)

    def fov_xy(self):
        return (focal_length_to_fov(1.0, self.relative_focal_length[0]), focal_length_to_fov(1.0, self.relative_focal_length[1]))

    def set_fov_xy(self, degrees_xy) -> None:
        self.relative_focal_length = (fov_to_focal_length(1.0, degrees_xy[0]), fov_to_focal_length(1.0, degrees_xy[1]))

    def to_training_view(self, xform_3x4, focal_length, resolution, principal_point) -> None:
        """set_camera_to_training_view: the view's transform; focal length relative to the resolution along fov_axis; the
        screen centre mirrored so that render_screen_center gives the principal point back"""
        self.matrix = np.ascontiguousarray(np.asarray(xform_3x4, dtype=np.float32)[:3, :4]).copy()
        r = float(resolution[self.fov_axis])
        self.relative_focal_length = (float(focal_length[0]) / r, float(focal_length[1]) / r)
        self.screen_center = (1.0 - float(principal_point[0]), 1.0 - float(principal_point[1]))

    def render_args(self, width: int, height: int):
        """(camera 3x4, (fx, fy) in pixels, (cx, cy) as image fractions) for a frame of the given size"""
        return self.matrix, calc_focal_length((width, height), self.relative_focal_length, self.fov_axis, self.zoom), render_screen_center(self.screen_center, self.zoom)
