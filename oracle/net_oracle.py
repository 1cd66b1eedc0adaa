"""CPU oracle (numpy) for the network part of the hot path — TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never by
the product (instant-ngp_b200/).  Restates, independently of the CUDA code:

  * hash-grid layout            tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:673-737 (constructor), common_device.h:886-895
  * hash-grid forward           grid.h:48-212 (kernel_grid), common_device.h:787-791 (coherent prime hash), :847-884 (grid_index),
                                :1000-1043 (pos_fract)
  * hash-grid backward          grid.h:214-320 (kernel_grid_backward)
  * SH degree 4                 common_device.h:475-503 (sh_enc), encodings/spherical_harmonics.h:44-72
  * fully fused MLP semantics   src/fully_fused_mlp.cu:499-557 (ReLU hidden layers, no output activation, fp16 storage between layers)
  * NerfNetwork wiring          include/neural-graphics-primitives/nerf_network.h:105-268, param order :357-372
  * Adam + EMA                  optimizers/adam.h:48-127, ema.h:63-77, exponential_decay.h:60-72

Pinning: offsets / n_params are checked against the reference's own known answers (tests/test_grid.cu:57-71) in
tests/test_oracle_goldens.py; forward / backward / optimizer outputs are checked against vectors produced by the compiled
reference (oracle/ref/ref_tcnn_harness.cu) committed under tests/golden/ when available.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F16 = np.float16
F32 = np.float32


# --------------------------------------------------------------------------------------------------------------------
# layout
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class GridLayout:
    n_levels: int
    n_features: int
    log2_hashmap_size: int
    base_resolution: int
    per_level_scale: float
    offsets: list = field(default_factory=list)
    resolutions: list = field(default_factory=list)
    scales: list = field(default_factory=list)
    n_pos_dims: int = 3

    @property
    def n_params(self) -> int:
        return self.offsets[-1] * self.n_features


def grid_layout(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, n_pos_dims=3) -> GridLayout:
    """grid.h:699-730 for GridType::Hash."""
    g = GridLayout(n_levels, n_features, log2_hashmap_size, base_resolution, float(per_level_scale), n_pos_dims=n_pos_dims)
    log2_scale = F32(math.log2(float(F32(per_level_scale))))
    offset = 0
    for l in range(n_levels):
        scale = F32(2.0 ** float(F32(l) * log2_scale)) * F32(base_resolution) - F32(1.0)  # exp2f, correctly rounded
        res = int(math.ceil(float(scale))) + 1
        max_params = 0xFFFFFFFF // 2
        params = max_params if float(F32(res) ** n_pos_dims) > float(max_params) else res ** n_pos_dims
        params = ((params + 7) // 8) * 8
        params = min(params, 1 << log2_hashmap_size)
        g.offsets.append(offset)
        g.resolutions.append(res)
        g.scales.append(F32(scale))
        offset += params
    g.offsets.append(offset)
    return g


def per_level_scale_for(aabb_scale: int, base_resolution: int, n_levels: int, desired_resolution: float = 2048.0) -> float:
    """src/testbed.cu:4241-4255."""
    return float(F32(math.exp(float(F32(math.log(float(F32(desired_resolution) * F32(aabb_scale) / F32(base_resolution)))) / F32(n_levels - 1)))))


@dataclass
class NerfLayout:
    grid: GridLayout
    n_hidden_density: int = 1
    n_hidden_rgb: int = 2

    def mlp_shapes(self, n_hidden):
        return [(64, 32)] + [(64, 64)] * (n_hidden - 1) + [(16, 64)]

    @property
    def density_shapes(self):
        return self.mlp_shapes(self.n_hidden_density)

    @property
    def rgb_shapes(self):
        return self.mlp_shapes(self.n_hidden_rgb)

    @property
    def n_mlp_params(self) -> int:
        return sum(a * b for a, b in self.density_shapes + self.rgb_shapes)

    @property
    def n_params(self) -> int:
        return self.n_mlp_params + self.grid.n_params

    def split(self, params):
        """flat buffer -> (density weight list, rgb weight list, grid) following nerf_network.h:357-372"""
        o = 0
        dens, rgb = [], []
        for (r, c) in self.density_shapes:
            dens.append(params[o:o + r * c].reshape(r, c))
            o += r * c
        for (r, c) in self.rgb_shapes:
            rgb.append(params[o:o + r * c].reshape(r, c))
            o += r * c
        return dens, rgb, params[o:o + self.grid.n_params]


# --------------------------------------------------------------------------------------------------------------------
# hash grid
# --------------------------------------------------------------------------------------------------------------------
PRIMES = (np.uint32(1), np.uint32(2654435761), np.uint32(805459861))
MAX_BASES = {2: 0xFFFF, 3: 0x659}   # common_device.h:854-866, indexed by N_DIMS
MAX_BASE_3D = MAX_BASES[3]


def _fma32(a, b, c):
    """fmaf for float32 arrays: exact product and sum in float64 (exact for the magnitudes used here), one rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(F32)


def _grid_index(g: GridLayout, level: int, *corner):
    """grid_index<N_DIMS, CoherentPrime> (common_device.h:847-884) for N_DIMS = len(corner) in {2, 3}."""
    D = len(corner)
    res = np.uint32(g.resolutions[level])
    size = np.uint32(g.offsets[level + 1] - g.offsets[level])
    with np.errstate(over="ignore"):
        if int(res) <= MAX_BASES[D]:
            stride = 1
            index = np.zeros_like(corner[0])
            for d in range(D):
                index = index + corner[d] * np.uint32(stride & 0xFFFFFFFF)
                stride *= int(res)
        else:
            stride = 0xFFFFFFFF
            index = np.zeros_like(corner[0])
        if int(size) < stride:
            index = np.zeros_like(corner[0])
            for d in range(D):
                index = index ^ (corner[d] * PRIMES[d])
    return index % size


def _level_setup(g: GridLayout, level: int, pos):
    scale = F32(g.scales[level])
    D = g.n_pos_dims
    p = [_fma32(np.full_like(pos[:, d], scale), pos[:, d], 0.5) for d in range(D)]
    fl = [np.floor(q) for q in p]
    gi = [f.astype(np.int32).astype(np.uint32) for f in fl]
    w1 = [(q - f).astype(F32) for q, f in zip(p, fl)]
    w0 = [(F32(1.0) - w).astype(F32) for w in w1]
    return gi, w0, w1


def _corner(g, l, c, gi, w0, w1):
    """index and fp32 weight of corner c (bit d of c = offset along axis d); weight product in axis order (grid.h:144-156)."""
    D = g.n_pos_dims
    bits = [(c >> d) & 1 for d in range(D)]
    idx = _grid_index(g, l, *[gi[d] + np.uint32(bits[d]) for d in range(D)])
    w = (w1[0] if bits[0] else w0[0])
    for d in range(1, D):
        w = (w * (w1[d] if bits[d] else w0[d])).astype(F32)
    return idx, w.astype(F32)


def grid_encode(g: GridLayout, grid_fp16: np.ndarray, pos: np.ndarray) -> np.ndarray:
    """kernel_grid: returns [n, L*F] float16, bit-exact with fp16 fused multiply-adds in corner order 0..2^D-1."""
    pos = np.ascontiguousarray(pos, dtype=F32)
    n = pos.shape[0]
    Fe = g.n_features
    table = grid_fp16.reshape(-1, Fe)
    out = np.zeros((n, g.n_levels * Fe), dtype=F16)
    for l in range(g.n_levels):
        gi, w0, w1 = _level_setup(g, l, pos)
        acc = np.zeros((n, Fe), dtype=np.float64)
        for c in range(1 << g.n_pos_dims):
            idx, w = _corner(g, l, c, gi, w0, w1)
            wh = w.astype(F16).astype(np.float64)[:, None]
            val = table[g.offsets[l] + idx.astype(np.int64)].astype(np.float64)
            acc = (wh * val + acc).astype(F16).astype(np.float64)  # one fp16 rounding per fma
        out[:, l * Fe:(l + 1) * Fe] = acc.astype(F16)
    return out


def grid_backward(g: GridLayout, pos: np.ndarray, dL_denc: np.ndarray, accumulate_fp16: bool = False) -> np.ndarray:
    """kernel_grid_backward: scatter (fp16)weight * (fp16)grad to the 2^D corners.  Returns float64 sums of the fp16 products
    (the reference accumulates them with fp16 atomics in arbitrary order; tests allow for that)."""
    pos = np.ascontiguousarray(pos, dtype=F32)
    Fe = g.n_features
    grad = np.zeros((g.offsets[-1], Fe), dtype=np.float64)
    dl = dL_denc.astype(F16)
    for l in range(g.n_levels):
        gi, w0, w1 = _level_setup(g, l, pos)
        gl = dl[:, l * Fe:(l + 1) * Fe].astype(np.float64)
        for c in range(1 << g.n_pos_dims):
            idx, w = _corner(g, l, c, gi, w0, w1)
            contrib = (w.astype(F16).astype(np.float64)[:, None] * gl).astype(F16).astype(np.float64)
            np.add.at(grad, g.offsets[l] + idx.astype(np.int64), contrib)
    return grad.reshape(-1)


# --------------------------------------------------------------------------------------------------------------------
# SH, MLP
# --------------------------------------------------------------------------------------------------------------------
def sh4(dirs01: np.ndarray) -> np.ndarray:
    d = dirs01.astype(F32) * F32(2.0) - F32(1.0)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    f = F32
    o = np.stack([
        np.full_like(x, f(0.28209479177387814)),
        f(-0.48860251190291987) * y,
        f(0.48860251190291987) * z,
        f(-0.48860251190291987) * x,
        f(1.0925484305920792) * xy,
        f(-1.0925484305920792) * yz,
        f(0.94617469575755997) * z2 - f(0.31539156525251999),
        f(-1.0925484305920792) * xz,
        f(0.54627421529603959) * x2 - f(0.54627421529603959) * y2,
        f(0.59004358992664352) * y * (f(-3.0) * x2 + y2),
        f(2.8906114426405538) * xy * z,
        f(0.45704579946446572) * y * (f(1.0) - f(5.0) * z2),
        f(0.3731763325901154) * z * (f(5.0) * z2 - f(3.0)),
        f(0.45704579946446572) * x * (f(1.0) - f(5.0) * z2),
        f(1.4453057213202769) * z * (x2 - y2),
        f(0.59004358992664352) * x * (-x2 + f(3.0) * y2),
    ], axis=1)
    return o.astype(F16)


def mlp_forward(weights_fp16, x_fp16, keep=False):
    """hidden = fp16(relu(W h)) ... ; output layer without activation.  fp32 accumulation."""
    h = x_fp16.astype(F32)
    acts = [x_fp16]
    for i, W in enumerate(weights_fp16):
        y = h @ W.astype(F32).T
        if i + 1 < len(weights_fp16):
            y = np.maximum(y, 0.0)
        y16 = y.astype(F16)
        acts.append(y16)
        h = y16.astype(F32)
    return (acts[-1], acts) if keep else acts[-1]


def mlp_backward(weights_fp16, acts, dL_dout_fp16):
    """returns (list of fp32 weight gradients, dL/dinput fp32).  Inter-layer gradients are rounded to fp16 like the
    reference's backward activations (fully_fused_mlp.cu:150-259)."""
    grads = [None] * len(weights_fp16)
    dy = dL_dout_fp16.astype(F32)
    dx = None
    for i in range(len(weights_fp16) - 1, -1, -1):
        x = acts[i].astype(F32)
        grads[i] = dy.T @ x
        dx = dy @ weights_fp16[i].astype(F32)
        if i > 0:
            dx = dx * (acts[i].astype(F32) > 0)
            dy = dx.astype(F16).astype(F32)
    return grads, dx


# --------------------------------------------------------------------------------------------------------------------
# NerfNetwork
# --------------------------------------------------------------------------------------------------------------------
def nerf_forward(L: NerfLayout, params_fp16: np.ndarray, coords: np.ndarray, keep=False):
    """coords [n,7] (pos, dt, dir) -> [n,4] float16 (rgb raw x3, density raw)."""
    dens_w, rgb_w, grid = L.split(params_fp16)
    enc = grid_encode(L.grid, grid, coords[:, 0:3])
    dens_out, dens_acts = mlp_forward(dens_w, enc, keep=True)
    rgb_in = np.concatenate([dens_out, sh4(coords[:, 4:7])], axis=1)
    rgb_out, rgb_acts = mlp_forward(rgb_w, rgb_in, keep=True)
    out = np.concatenate([rgb_out[:, 0:3], dens_out[:, 0:1]], axis=1).astype(F16)
    if keep:
        return out, (enc, dens_acts, rgb_acts)
    return out


def nerf_density(L: NerfLayout, params_fp16: np.ndarray, pos: np.ndarray) -> np.ndarray:
    dens_w, _, grid = L.split(params_fp16)
    enc = grid_encode(L.grid, grid, pos[:, 0:3])
    return mlp_forward(dens_w, enc)[:, 0]


def nerf_backward(L: NerfLayout, params_fp16: np.ndarray, coords: np.ndarray, dL_dout_fp16: np.ndarray):
    """dL_dout [n,4] fp16 -> flat gradient (float64, same layout as params)."""
    dens_w, rgb_w, _ = L.split(params_fp16)
    _, (enc, dens_acts, rgb_acts) = nerf_forward(L, params_fp16, coords, keep=True)
    n = coords.shape[0]
    d_rgb_out = np.zeros((n, 16), dtype=F16)
    d_rgb_out[:, 0:3] = dL_dout_fp16[:, 0:3]
    rgb_grads, d_rgb_in = mlp_backward(rgb_w, rgb_acts, d_rgb_out)
    d_dens_out = d_rgb_in[:, 0:16].astype(F16)
    # add_density_gradient (nerf_network.h:62-74): fp16 add into row 0
    d_dens_out[:, 0] = (d_dens_out[:, 0].astype(F32) + dL_dout_fp16[:, 3].astype(F32)).astype(F16)
    dens_grads, d_enc = mlp_backward(dens_w, dens_acts, d_dens_out)
    ggrid = grid_backward(L.grid, coords[:, 0:3], d_enc.astype(F16))
    flat = np.concatenate([g.reshape(-1).astype(np.float64) for g in dens_grads + rgb_grads] + [ggrid])
    return flat


# --------------------------------------------------------------------------------------------------------------------
# optimizer
# --------------------------------------------------------------------------------------------------------------------
def adam_ema_step(n_matrix, w32, w16, ema16, grads16, m1, m2, steps, *, lr, beta1, beta2, eps, l2_reg, loss_scale, ema_decay, step,
                  optimize_matrix=True, optimize_non_matrix=True):
    """One Ema{Adam} step in place (adam.h:48-127, ema.h:63-77). `step` is the 1-based optimizer step."""
    n = w32.size
    g = grads16.astype(F32) / F32(loss_scale)
    idx = np.arange(n)
    is_matrix = idx < n_matrix
    active = np.where(is_matrix, optimize_matrix, optimize_non_matrix & (g != 0))
    g = np.where(is_matrix, g + F32(l2_reg) * w32, g).astype(F32)
    a = active
    m1[a] = (F32(beta1) * m1[a] + F32(1 - F32(beta1)) * g[a]).astype(F32)
    m2[a] = (F32(beta2) * m2[a] + F32(1 - F32(beta2)) * g[a] * g[a]).astype(F32)
    steps[a] += 1
    s = steps[a].astype(F32)
    lr_t = (F32(lr) * np.sqrt(F32(1) - np.power(F32(beta2), s)) / (F32(1) - np.power(F32(beta1), s))).astype(F32)
    eff = (lr_t / (np.sqrt(m2[a]) + F32(eps))).astype(F32)
    w32[a] = (w32[a] - eff * m1[a]).astype(F32)
    w16[a] = w32[a].astype(F16)
    debias_old = F32(1 - F32(ema_decay) ** (step - 1))
    debias_new = F32(1.0 / (1 - F32(ema_decay) ** step))
    ema16[:] = ((ema16.astype(F32) * F32(ema_decay) * debias_old + w16.astype(F32) * F32(1 - F32(ema_decay))) * debias_new).astype(F16)
    grads16[:] = 0


def seed_seq_first(seed: int, n: int = 2) -> int:
    """First word of std::seed_seq{(uint32_t)seed}.generate over n words ([rand.util.seedseq]); the reference seeds its
    default_rng_t this way (trainer.h / testbed.cu: `std::seed_seq seq{seed}; seq.generate(...); rng = pcg32(seeds.front())`)."""
    M = 0xFFFFFFFF
    v = [seed & M]
    s = len(v)
    b = [0x8b8b8b8b] * n
    t = 11 if n >= 623 else 7 if n >= 68 else 5 if n >= 39 else 3 if n >= 7 else (n - 1) // 2
    p = (n - t) // 2
    q = p + t
    m = max(s + 1, n)
    T = lambda x: (x ^ (x >> 27)) & M
    for k in range(m):
        r1 = (1664525 * T(b[k % n] ^ b[(k + p) % n] ^ b[(k - 1) % n])) & M
        r2 = (r1 + (s if k == 0 else (k % n + v[k - 1]) if k <= s else k % n)) & M
        b[(k + p) % n] = (b[(k + p) % n] + r1) & M
        b[(k + q) % n] = (b[(k + q) % n] + r2) & M
        b[k % n] = r2
    for k in range(m, m + n):
        r3 = (1566083941 * T((b[k % n] + b[(k + p) % n] + b[(k - 1) % n]) & M)) & M
        r4 = (r3 - k % n) & M
        b[(k + p) % n] ^= r3
        b[(k + q) % n] ^= r4
        b[k % n] = r4
    return b[0]


# --------------------------------------------------------------------------------------------------------------------
# pcg32 (for seeded test inputs that mirror the reference's generators; pcg32.h)
# --------------------------------------------------------------------------------------------------------------------
class Pcg32:
    MULT = 0x5851f42d4c957f2d
    MASK = (1 << 64) - 1

    def __init__(self, initstate=None, initseq=1):
        if initstate is None:
            self.state, self.inc = 0x853c49e6748fea9b, 0xda3e39cb94b95bdb
        else:
            self.state = 0
            self.inc = ((initseq << 1) | 1) & self.MASK
            self.next_uint()
            self.state = (self.state + initstate) & self.MASK
            self.next_uint()

    def next_uint(self):
        old = self.state
        self.state = (old * self.MULT + self.inc) & self.MASK
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def next_float(self):
        u = (self.next_uint() >> 9) | 0x3f800000
        return float(np.uint32(u).view(np.float32) - np.float32(1.0))

    def advance(self, delta=1 << 32):
        cur_mult, cur_plus, acc_mult, acc_plus = self.MULT, self.inc, 1, 0
        delta &= self.MASK
        while delta > 0:
            if delta & 1:
                acc_mult = (acc_mult * cur_mult) & self.MASK
                acc_plus = (acc_plus * cur_mult + cur_plus) & self.MASK
            cur_plus = ((cur_mult + 1) * cur_plus) & self.MASK
            cur_mult = (cur_mult * cur_mult) & self.MASK
            delta >>= 1
        self.state = (acc_mult * self.state + acc_plus) & self.MASK
