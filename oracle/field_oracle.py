"""CPU oracle (numpy) for the image / SDF primitives — TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never by the
product (instant-ngp_b200/).  Restates, independently of the CUDA code:

  * NetworkWithInputEncoding      tiny-cuda-nn/include/tiny-cuda-nn/network_with_input_encoding.h:38-170 (params: MLP | grid, :121-128)
  * tcnn losses                   losses/l2.h:36-71, l1.h, mape.h:40-77, smape.h, relative_l2.h (pdf = 1)
  * generate_random_uniform       random.h:40-67 (thread i draws 4 values for indices i + n_threads * j)
  * stratify2_kernel              src/testbed_image.cu:66-82
  * eval_image_kernel_and_snap    src/testbed_image.cu:176-229
  * shuffle / permute             tiny-cuda-nn/common_device.h:1097-1111
  * Trainer::initialize_params    trainer.h:69-87, gpu_matrix.h:292-307 (Xavier uniform on the host), grid.h (U(-1e-4, 1e-4))

Pinning: the grid + MLP arithmetic is the code pinned for the NeRF network (tests/test_oracle_vs_reference_tcnn.py) applied
to D = 2 / 3 positions; the NetworkWithInputEncoding wiring, losses and initialisation are checked against vectors produced by
the compiled reference (oracle/ref/ref_tcnn_harness.cu, `field` section) under tests/golden/ when available.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import net_oracle as no

F16, F32 = np.float16, np.float32

LOSS_L2, LOSS_L1, LOSS_MAPE, LOSS_SMAPE, LOSS_RELATIVE_L2 = 0, 1, 2, 3, 6


@dataclass
class FieldLayout:
    grid: no.GridLayout
    n_hidden: int = 2
    n_out: int = 3

    @property
    def shapes(self):
        return [(64, 32)] + [(64, 64)] * (self.n_hidden - 1) + [(16, 64)]

    @property
    def n_mlp_params(self) -> int:
        return sum(a * b for a, b in self.shapes)

    @property
    def n_params(self) -> int:
        return self.n_mlp_params + self.grid.n_params

    def split(self, params):
        o, ws = 0, []
        for (r, c) in self.shapes:
            ws.append(params[o:o + r * c].reshape(r, c))
            o += r * c
        return ws, params[o:o + self.grid.n_params]


def field_forward(L: FieldLayout, params_fp16, pos, keep=False):
    """[n, D] float32 -> [n, 16] float16 (padded output row)."""
    ws, grid = L.split(params_fp16)
    enc = no.grid_encode(L.grid, grid, pos)
    out, acts = no.mlp_forward(ws, enc, keep=True)
    return (out, acts) if keep else out


def loss_evaluate(loss_type: int, pred_fp16, targets, loss_scale: float, n_out: int):
    """pred [n, 16] fp16, targets [n, n_out] f32 -> (values [n, n_out] f32, dL/dout [n, 16] fp16); n_total = n * n_out."""
    n = pred_fp16.shape[0]
    n_total = F32(n * n_out)
    p = pred_fp16[:, :n_out].astype(F32)
    t = np.asarray(targets, dtype=F32).reshape(n, n_out)
    d = (p - t).astype(F32)
    one = F32(1.0)
    if loss_type == LOSS_L2:
        values = (d * d / n_total).astype(F32)
        grad = (F32(2.0) * d).astype(F32)
    elif loss_type == LOSS_L1:
        values = (np.abs(d) / n_total).astype(F32)
        grad = np.copysign(one, d).astype(F32)
    elif loss_type == LOSS_MAPE:
        scale = (one / (np.abs(t) + F32(1e-2))).astype(F32)
        values = (np.abs(d) * scale / n_total).astype(F32)
        grad = np.copysign(scale, d).astype(F32)
    elif loss_type == LOSS_SMAPE:
        scale = (one / (F32(0.5) * (np.abs(t) + np.abs(p)) + F32(1e-2))).astype(F32)
        values = (np.abs(d) * scale / n_total).astype(F32)
        grad = np.copysign(scale, d).astype(F32)
    elif loss_type == LOSS_RELATIVE_L2:
        psq = (p * p + F32(0.01)).astype(F32)
        values = ((d * d / psq).astype(F32) / n_total).astype(F32)
        grad = ((F32(2.0) * d).astype(F32) / psq).astype(F32)
    else:
        raise ValueError(loss_type)
    g16 = np.zeros((n, 16), dtype=F16)
    g16[:, :n_out] = ((F32(loss_scale) * grad).astype(F32) / n_total).astype(F16)
    return values, g16


def field_backward(L: FieldLayout, params_fp16, pos, dL_dout_fp16):
    """dL/dout [n, 16] fp16 -> flat gradient (float64), layout MLP | grid."""
    ws, _ = L.split(params_fp16)
    _, acts = field_forward(L, params_fp16, pos, keep=True)
    wg, d_enc = no.mlp_backward(ws, acts, dL_dout_fp16)
    gg = no.grid_backward(L.grid, pos, d_enc.astype(F16))
    return np.concatenate([g.reshape(-1).astype(np.float64) for g in wg] + [gg])


def field_train_step(L: FieldLayout, params_fp16, pos, targets, loss_type, loss_scale):
    out = field_forward(L, params_fp16, pos)
    values, g16 = loss_evaluate(loss_type, out, targets, loss_scale, L.n_out)
    return out, values, field_backward(L, params_fp16, pos, g16)


def field_init_params(L: FieldLayout, seed: int) -> np.ndarray:
    """Trainer::initialize_params for NetworkWithInputEncoding: std::seed_seq{seed} -> pcg32, Xavier per matrix, then the grid."""
    rng = no.Pcg32(no.seed_seq_first(seed))
    out = np.empty(L.n_params, dtype=F32)
    o = 0
    for (r, c) in L.shapes:
        scale = F32(math.sqrt(F32(6.0) / F32(r + c)))
        for i in range(r * c):
            out[o + i] = F32(rng.next_float()) * F32(2.0) * scale - scale
        o += r * c
    out[o:] = tcnn_random_uniform(rng, L.grid.n_params, F32(-1e-4), F32(1e-4), advance=False)
    return out


def tcnn_random_uniform(rng: no.Pcg32, n_elements: int, lower=F32(0.0), upper=F32(1.0), advance=True) -> np.ndarray:
    """generate_random_uniform (random.h:40-67).  With advance=True the generator is moved on by n_elements like the host code."""
    n_threads_req = (n_elements + 3) // 4
    n_threads = ((n_threads_req + 127) // 128) * 128
    out = np.empty(n_elements, dtype=F32)
    state0 = rng.state
    for i in range(min(n_threads, n_elements)):
        rng.state = state0
        rng.advance(i * 4)
        for j in range(4):
            idx = i + n_threads * j
            if idx >= n_elements:
                break
            out[idx] = F32(rng.next_float()) * F32(upper - lower) + F32(lower)
    rng.state = state0
    if advance:
        rng.advance(n_elements)
    return out


def stratify2(positions: np.ndarray, log2_batch_size: int) -> np.ndarray:
    n = positions.shape[0]
    log2_size = log2_batch_size // 2
    size = F32(1 << log2_size)
    i = np.arange(n, dtype=np.uint32) & np.uint32((1 << log2_batch_size) - 1)
    x = (i & np.uint32((1 << log2_size) - 1)).astype(F32)
    y = (i >> np.uint32(log2_size)).astype(F32)
    out = np.empty_like(positions, dtype=F32)
    out[:, 0] = (positions[:, 0] / size).astype(F32) + (x / size).astype(F32)
    out[:, 1] = (positions[:, 1] / size).astype(F32) + (y / size).astype(F32)
    return out


def _linear_to_srgb(a: np.ndarray) -> np.ndarray:
    import ctypes as C

    from . import march_oracle as mo

    a = np.ascontiguousarray(a, dtype=F32)
    out = np.empty_like(a)
    mo.lib().orc_linear_to_srgb_n(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), a.size)
    return out


def eval_image_and_snap(image: np.ndarray, positions: np.ndarray, snap: bool, linear_colors: bool):
    """image [h, w, 4] float32 (or float16) -> (positions' [n, 2], targets [n, 3])."""
    h, w = image.shape[:2]
    pos = np.asarray(positions, dtype=F32).copy()
    img = image.astype(F32)

    def read(x, y):
        c = img[y, x, :3].astype(F32)
        return c if linear_colors else _linear_to_srgb(c)

    fw, fh = F32(w), F32(h)
    if snap:
        ix = np.floor(pos[:, 0] * fw).astype(np.int32)
        iy = np.floor(pos[:, 1] * fh).astype(np.int32)
        pos[:, 0] = ((ix.astype(F32) + F32(0.5)) / fw).astype(F32)
        pos[:, 1] = ((iy.astype(F32) + F32(0.5)) / fh).astype(F32)
        val = read(np.clip(ix, 0, w - 1), np.clip(iy, 0, h - 1))
    else:
        px = np.clip((pos[:, 0] * fw).astype(F32) - F32(0.5), F32(0.0), F32(fw - F32(F32(1.0) + F32(1e-4)))).astype(F32)
        py = np.clip((pos[:, 1] * fh).astype(F32) - F32(0.5), F32(0.0), F32(fh - F32(F32(1.0) + F32(1e-4)))).astype(F32)
        ix, iy = px.astype(np.int32), py.astype(np.int32)
        wx, wy = (px - ix.astype(F32)).astype(F32), (py - iy.astype(F32)).astype(F32)
        jx, jy = np.clip(ix, 0, w - 2), np.clip(iy, 0, h - 2)
        one = F32(1.0)
        w00 = ((one - wx) * (one - wy)).astype(F32)[:, None]
        w10 = (wx * (one - wy)).astype(F32)[:, None]
        w01 = ((one - wx) * wy).astype(F32)[:, None]
        w11 = (wx * wy).astype(F32)[:, None]
        val = ((((w00 * read(jx, jy)).astype(F32) + (w10 * read(jx + 1, jy)).astype(F32)).astype(F32) + (w01 * read(jx, jy + 1)).astype(F32)).astype(F32)
               + (w11 * read(jx + 1, jy + 1)).astype(F32)).astype(F32)
    return pos, val.astype(F32)


def image_training_data(rng: no.Pcg32, n: int, image, stratify: bool, snap: bool, linear_colors: bool):
    """train_image's generate_training_data (src/testbed_image.cu:241-283); advances rng by 2 n."""
    pos = tcnn_random_uniform(rng, 2 * n).reshape(n, 2)
    if stratify:
        log2_n = int(round(math.log2(n)))
        assert (1 << log2_n) == n and log2_n % 2 == 0
        pos = stratify2(pos, log2_n)
    return eval_image_and_snap(image, pos, snap, linear_colors)


def permute(num, size):
    return ((np.asarray(num, dtype=np.uint64) * np.uint64(1434869437) + np.uint64(2097192037)) % np.uint64(size)).astype(np.uint32)


def shuffle(arr: np.ndarray, stride: int, seed: int) -> np.ndarray:
    a = np.asarray(arr).reshape(-1, stride)
    n = a.shape[0]
    idx = permute((np.arange(n, dtype=np.uint64) + np.uint64(seed)) & np.uint64(0xFFFFFFFF), n)
    return a[idx].reshape(arr.shape)
