"""CPU: per-image exposure in the loss kernel's oracle (testbed_nerf.cu:979-995 target colour, :1142-1155 gradient) — the target is the
view's colour times 2^exposure per channel, and the accumulated gradient is the reference's expression
loss_scale * (-dL/dprediction [/ srgb_to_linear'(target)]) * 2^exposure * ln 2, summed over the rays of the view."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import march_oracle as M

LN2 = 0.6931471805599453


def one_ray(exposure, linear_colors, with_gradient=True, n=10, loss_type=0):
    imgs = np.zeros((2, 8, 8, 4), dtype=np.float32)
    imgs[..., 0], imgs[..., 1], imgs[..., 2], imgs[..., 3] = 0.30, 0.55, 0.20, 0.6   # premultiplied colour, alpha 0.6
    cam = np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.5], [0, 0, 1, -1.0]], dtype=np.float32)
    cfg = util.make_train_cfg(aabb_scale=1, random_bg_color=0, loss_type=loss_type, near_distance=0.0, linear_colors=linear_colors)
    cfg.background_color[0], cfg.background_color[1], cfg.background_color[2] = 0.8, 0.1, 0.4
    views, keep = util.make_views(imgs, [cam, cam], 8.0)
    expo = None if exposure is None else np.ascontiguousarray(np.asarray(exposure, dtype=np.float32).reshape(2, 3))
    grad = np.zeros((2, 3), dtype=np.float32)
    cfg.cam_exposure = None if expo is None else expo.ctypes.data
    cfg.cam_exposure_gradient = grad.ctypes.data if (with_gradient and expo is not None) else None
    rng = M.pcg32_seed(3)
    coords = np.zeros((n, 7), dtype=np.float32)
    coords[:, 0:3] = np.linspace(0.3, 0.7, n)[:, None]
    coords[:, 3] = 0.05
    coords[:, 4:7] = 0.5
    numsteps = np.array([[n, 0]], dtype=np.uint32)
    ray_indices = np.array([1], dtype=np.uint32)     # ray 1 of 2 -> view 1 (image_idx = ray * n_views / n_rays)
    rays = np.array([[0.5, 0.5, -1.0, 0, 0, 1]], dtype=np.float32)
    co = np.zeros((64, 7), dtype=np.float32)
    dl = np.zeros((64, 4), dtype=np.float16)
    loss = np.zeros(1, dtype=np.float32)
    o = np.random.default_rng(11).normal(0, 0.7, size=(n, 4))
    o[:, 3] -= 1.0
    no = np.ascontiguousarray(o.astype(np.float16))
    got = M.lib().orc_compute_loss(1, 2, rng[0], rng[1], C.byref(cfg), C.addressof(views), 2, no.ctypes.data, 64, ray_indices.ctypes.data, rays.ctypes.data,
                                   numsteps.ctypes.data, coords.ctypes.data, co.ctypes.data, dl.ctypes.data, loss.ctypes.data, np.float32(1.0))
    assert got == n
    return dict(dl=dl[:n].copy(), loss=float(loss[0]), grad=grad, net=no.astype(np.float64), cfg=cfg)


def composited(net):
    mn = np.sqrt(3.0) / 1024.0
    dt = 0.05 * (mn * 128.0 - mn) + mn
    rgb = 1.0 / (1.0 + np.exp(-net[:, :3]))
    alpha = 1.0 - np.exp(-np.exp(net[:, 3]) * dt)
    T_before = np.concatenate([[1.0], np.cumprod(1 - alpha)[:-1]])
    return (rgb * (alpha * T_before)[:, None]).sum(0), float(np.prod(1 - alpha))


def srgb_to_linear(s):
    s = np.asarray(s, dtype=np.float64)
    return np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)


def linear_to_srgb(l):
    l = np.asarray(l, dtype=np.float64)
    return np.where(l < 0.0031308, 12.92 * l, 1.055 * l ** 0.41666 - 0.055)


def test_zero_exposure_changes_nothing():
    a = one_ray(None, 1)
    b = one_ray(np.zeros(6), 1)
    assert a["dl"].tobytes() == b["dl"].tobytes() and a["loss"] == b["loss"]
    c = one_ray(np.zeros(6), 0)
    d = one_ray(None, 0)
    assert c["dl"].tobytes() == d["dl"].tobytes() and c["loss"] == d["loss"]


@pytest.mark.parametrize("linear_colors", [1, 0])
def test_exposure_scales_the_target_and_accumulates_the_reference_gradient(linear_colors):
    e = np.array([[9.0, 9.0, 9.0], [0.5, -0.25, 0.125]])       # view 0's is never read: the ray belongs to view 1
    r = one_ray(e, linear_colors)
    scale = 2.0 ** e[1]
    bg_lin = srgb_to_linear([0.8, 0.1, 0.4])
    colour, alpha = np.array([0.30, 0.55, 0.20]), 0.6
    rgb_ray, T_end = composited(r["net"])
    if linear_colors:
        target, bg = scale * colour + (1 - alpha) * bg_lin, bg_lin
    else:   # EColorSpace::Linear images, training in sRGB (the reference's default, testbed_nerf.cu:985-991): composite, then convert
        bg = linear_to_srgb(bg_lin)
        target = linear_to_srgb(scale * colour + (1 - alpha) * bg_lin)
    pred = rgb_ray + T_end * bg
    # L2: loss (pred - target)^2 per channel, shown as the mean over channels / n_rays; gradient 2 (pred - target)
    assert r["loss"] == pytest.approx(float(np.mean((pred - target) ** 2)) / 2.0, rel=2e-3)
    dgt = -2.0 * (pred - target)
    if not linear_colors:
        dgt = dgt / np.where(target <= 0.04045, 1 / 12.92, 2.4 / 1.055 * ((target + 0.055) / 1.055) ** 1.4)
    want = (128.0 / 2.0) * dgt * scale * LN2
    assert np.allclose(r["grad"][1], want, rtol=3e-3, atol=1e-5)
    assert not r["grad"][0].any()
    # a brighter view asks for a brighter prediction: the colour gradients differ from the zero-exposure ones
    assert r["dl"].tobytes() != one_ray(np.zeros(6), linear_colors)["dl"].tobytes()
    # no accumulator, same sample gradients
    assert one_ray(e, linear_colors, with_gradient=False)["dl"].tobytes() == r["dl"].tobytes()
