"""CPU: include/ngp_detmath.h against libm.  The header is shared by the CUDA product and the C oracle (that is what makes the march bit-exact
between them), so a mistake in it cancels out of every CUDA-vs-oracle test: here its log / exp / pow are compared with numpy's float64 functions
over the ranges the march and the colour transfer curves use.  Bound: 1 ulp for log and exp (the header promises 2) (the reference's fast-math
intrinsics are looser: __logf 2^-21.4 absolute, __expf 2 ulp + argument reduction)."""
import numpy as np

from oracle import march_oracle as M


def detmath(which, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), dtype=np.float32)
    out = np.empty_like(x)
    M.lib().orc_detmath_n(which, x.ctypes.data, y.ctypes.data, out.ctypes.data, x.size)
    return out


def ulps(got, want64):
    want = want64.astype(np.float32)
    spacing = np.spacing(np.abs(want)).astype(np.float64)
    return np.abs(got.astype(np.float64) - want64) / np.maximum(spacing, np.finfo(np.float32).tiny)


def test_log_exp_pow_are_within_two_ulp_of_libm():
    rng = np.random.default_rng(0)
    # log: ray parameters t in [1e-4, 2e4] (to_stepping_space), 1 + cone_angle, step sizes
    t = np.exp(rng.uniform(np.log(1e-4), np.log(2e4), 200000)).astype(np.float32)
    t = np.concatenate([t, np.float32([1.0, 1.0 + 1.0 / 256.0, 1.73205080757 / 1024.0, 1.73205080757, 0.5, 2.0, 1e-30, 3e38])])
    assert ulps(detmath(0, t), np.log(t.astype(np.float64))).max() <= 1.0       # measured 0.74
    # exp: stepping space n * log1p(c) in [-12, 12], -density * dt in [-80, 0], activations in [-15, 15]
    x = np.concatenate([rng.uniform(-80, 15, 200000), np.float64([0.0, -0.0, 1.0, -1.0, 10.0, -10.0, 88.0, -87.0])]).astype(np.float32)
    assert ulps(detmath(1, x), np.exp(x.astype(np.float64))).max() <= 1.0       # measured 0.97
    assert detmath(1, np.float32([100.0]))[0] == np.inf and detmath(1, np.float32([-200.0]))[0] == 0.0 and detmath(0, np.float32([0.0]))[0] == -np.inf
    # pow: the sRGB curves x^2.4 and x^0.41666 on [0.003, 1]; the error of log is amplified by y * log(x) <= 14 here
    c = rng.uniform(0.003, 1.0, 200000).astype(np.float32)
    for e in (2.4, 0.41666):
        y = np.full_like(c, e)
        want = np.power(c.astype(np.float64), np.float64(np.float32(e)))
        assert ulps(detmath(2, c, y), want).max() <= 24.0          # log's error x |y log x| (up to 14 here) + exp's; measured 16.7 / 3.9
        assert np.abs(detmath(2, c, y).astype(np.float64) - want).max() <= 1e-7


def test_log_and_exp_are_monotone_and_inverse_on_the_stepping_range():
    t = np.linspace(0.05, 16.0, 100001).astype(np.float32)
    l = detmath(0, t)
    assert (np.diff(l) >= 0).all()
    back = detmath(1, l)
    assert np.abs(back / t - 1.0).max() <= 4e-7
