"""CPU: the three NeRF train modes of the loss kernel's oracle (ETrainMode, common.h:47-51; gradients as formed in
fused_kernels/train_nerf.cuh:391-410) — each mode's dL/d(network output) against a finite-difference derivative of the objective
that mode stands for, on one ray with synthetic network outputs:
  Nerf      L = sum_c loss(target_c, composited_c)                                   (testbed_nerf.cu:1006, 1078-1110)
  Rfl       L = sum_c [ sum_k w_k loss(target_c, rgb_kc) + T_end loss(target_c, bg_c) ]   (train_nerf.cuh:231, 251-254, 393-397)
  RflRelax  has no closed-form objective (its gradient is evaluated at a lerped colour): checked for its defining identities."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import march_oracle as M


def run_oracle(train_mode, net_out, n=12, loss_type=0):
    """one ray with n samples; returns (dloss [n,4] float32 with the loss scale divided out, target colour, bg colour)"""
    imgs = np.zeros((1, 8, 8, 4), dtype=np.float32)
    imgs[..., 0], imgs[..., 1], imgs[..., 2], imgs[..., 3] = 0.30, 0.55, 0.20, 0.6   # premultiplied colour, alpha 0.6
    cam = np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.5], [0, 0, 1, -1.0]], dtype=np.float32)
    cfg = util.make_train_cfg(aabb_scale=1, random_bg_color=0, loss_type=loss_type, near_distance=0.0, linear_colors=1, train_mode=train_mode)
    cfg.background_color[0], cfg.background_color[1], cfg.background_color[2] = 0.8, 0.1, 0.4
    views, keep = util.make_views(imgs, [cam], 8.0)
    rng = M.pcg32_seed(3)
    coords = np.zeros((n, 7), dtype=np.float32)
    coords[:, 0:3] = np.linspace(0.3, 0.7, n)[:, None]
    coords[:, 3] = 0.05      # warped dt
    coords[:, 4:7] = 0.5
    numsteps = np.array([[n, 0]], dtype=np.uint32)
    ray_indices = np.array([0], dtype=np.uint32)
    rays = np.array([[0.5, 0.5, -1.0, 0, 0, 1]], dtype=np.float32)
    co = np.zeros((64, 7), dtype=np.float32)
    dl = np.zeros((64, 4), dtype=np.float16)
    loss = np.zeros(1, dtype=np.float32)
    no = np.ascontiguousarray(net_out.astype(np.float16))
    got = M.lib().orc_compute_loss(1, 1, rng[0], rng[1], C.byref(cfg), C.addressof(views), 1, no.ctypes.data, 64, ray_indices.ctypes.data, rays.ctypes.data,
                                   numsteps.ctypes.data, coords.ctypes.data, co.ctypes.data, dl.ctypes.data, loss.ctypes.data, np.float32(1.0))
    assert got == numsteps[0, 0] == n            # nothing saturates in these inputs: the whole ray is consumed, bg term included
    srgb_to_linear = lambda s: s / 12.92 if s <= 0.04045 else ((s + 0.055) / 1.055) ** 2.4
    bg = np.array([srgb_to_linear(v) for v in (0.8, 0.1, 0.4)])
    target = np.array([0.30, 0.55, 0.20]) + (1 - 0.6) * bg
    return dl[:n].astype(np.float64) / 128.0, target, bg, float(M.lib().orc_from_stepping_space(0, C.byref(cfg.march)) * 0 + 1)


def composite(o, dt):
    rgb = 1.0 / (1.0 + np.exp(-o[:, :3]))            # Logistic
    sigma = np.exp(o[:, 3])                          # Exponential
    alpha = 1.0 - np.exp(-sigma * dt)
    T_before = np.concatenate([[1.0], np.cumprod(1 - alpha)[:-1]])
    w = alpha * T_before
    return rgb, w, float(np.prod(1 - alpha))


def unwarped_dt():
    # unwarp_dt(0.05) (nerf_device.cuh:372-377): dt = 0.05 * (max - min) + min, min = sqrt(3)/1024, max = 128 min
    mn = np.sqrt(3.0) / 1024.0
    mx = mn * 128.0
    return 0.05 * (mx - mn) + mn


def objective(mode, o, target, bg):
    rgb, w, T_end = composite(o, unwarped_dt())
    l2 = lambda t, p: (p - t) ** 2
    if mode == 0:
        c = (w[:, None] * rgb).sum(0) + T_end * bg
        return l2(target, c).sum()
    return (w[:, None] * l2(target[None, :], rgb)).sum() + T_end * l2(target, bg).sum()


@pytest.mark.parametrize("mode", [0, 1])
def test_gradient_is_the_derivative_of_the_modes_objective(mode):
    rng = np.random.default_rng(mode)
    n = 12
    o = np.concatenate([rng.normal(0, 0.7, size=(n, 3)), rng.normal(3.0, 0.4, size=(n, 1))], axis=1)   # densities that make alpha ~0.1-0.3 at this dt
    o = o.astype(np.float16).astype(np.float64)      # what the kernel reads
    dl, target, bg, _ = run_oracle(mode, o, n=n)
    num = np.zeros_like(o)
    eps = 1e-4
    for k in range(n):
        for c in range(4):
            a, b = o.copy(), o.copy()
            a[k, c] += eps
            b[k, c] -= eps
            num[k, c] = (objective(mode, a, target, bg) - objective(mode, b, target, bg)) / (2 * eps)
    scale = np.abs(num).max()
    err = np.abs(dl - num).max()
    print("mode", mode, "max |grad|", scale, "max abs err", err)
    assert err <= 2e-2 * scale        # fp16 storage of the gradients (scaled by 128) and float32 arithmetic


def test_rfl_relax_identities():
    """RflRelax: the colour gradient is weight * dloss/dc evaluated at the lerped colour, and for an opaque last sample (alpha -> 1,
    nothing behind it) the lerped colour is the sample's own colour: the gradient then equals Rfl's colour gradient"""
    n = 6
    o = np.zeros((n, 4))
    o[:, :3] = np.random.default_rng(5).normal(0, 0.7, size=(n, 3))
    o[:, 3] = -6.0                   # transparent ...
    o[-1, 3] = 9.0                   # ... except an opaque last sample
    o = o.astype(np.float16).astype(np.float64)
    d_relax, target, bg, _ = run_oracle(2, o, n=n)
    d_rfl, _, _, _ = run_oracle(1, o, n=n)
    assert np.allclose(d_relax[-1, :3], d_rfl[-1, :3], rtol=2e-2, atol=1e-6)
    assert np.isfinite(d_relax).all() and np.abs(d_relax[:, 3]).max() > 0
