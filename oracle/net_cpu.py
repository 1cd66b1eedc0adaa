"""ctypes wrapper of oracle/libngp_net_cpu.so (oracle/ngp_net_cpu.c): the multi-threaded fp32 CPU port of the NeRF network used as bench.py's
CPU baseline — TEST / MEASUREMENT INFRASTRUCTURE ONLY (imported by tests/ and bench.py's cpu_baseline / --impl reference leg)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_lib = None


def build() -> Path:
    src, out = HERE / "ngp_net_cpu.c", HERE / "libngp_net_cpu.so"
    if not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", str(src), "-o", str(out), "-lm"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed building the CPU baseline:\n" + r.stdout + r.stderr)
    return out


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(str(build()))
        vp, u32 = C.c_void_p, C.c_uint32
        l.orc_net_cpu.restype = None
        l.orc_net_cpu.argtypes = [u32, vp, vp, vp, vp, vp, u32, u32, vp, vp, vp, u32, u32]
        l.orc_net_cpu_threads.restype = C.c_int
        l.orc_adam_cpu.restype = None
        l.orc_adam_cpu.argtypes = [C.c_uint64, vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, u32]
        _lib = l
    return _lib


def threads() -> int:
    return int(lib().orc_net_cpu_threads())


class NetCpu:
    """L: oracle.net_oracle.NerfLayout; params: flat parameter buffer (any float dtype; held as fp32)"""

    def __init__(self, L, params):
        self.L = L
        g = L.grid
        self.params = np.ascontiguousarray(params, dtype=np.float32)
        self.offsets = np.ascontiguousarray(g.offsets, dtype=np.uint32)
        self.res = np.ascontiguousarray(g.resolutions, dtype=np.uint32)
        self.scales = np.ascontiguousarray(g.scales, dtype=np.float32)

    def _call(self, coords, dl, grads, out):
        g = self.L.grid
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        lib().orc_net_cpu(coords.shape[0], coords.ctypes.data, self.params.ctypes.data, dl.ctypes.data if dl is not None else None,
                          grads.ctypes.data if grads is not None else None, out.ctypes.data if out is not None else None, g.n_levels, g.n_features,
                          self.offsets.ctypes.data, self.res.ctypes.data, self.scales.ctypes.data, self.L.n_hidden_density, self.L.n_hidden_rgb)

    def forward(self, coords) -> np.ndarray:
        out = np.empty((len(coords), 4), dtype=np.float32)
        self._call(coords, None, None, out)
        return out

    def forward_backward(self, coords, dL_dout):
        out = np.empty((len(coords), 4), dtype=np.float32)
        grads = np.zeros(self.params.shape[0], dtype=np.float32)
        self._call(coords, np.ascontiguousarray(dL_dout, dtype=np.float32), grads, out)
        return out, grads

    def adam_step(self, grads, m1, m2, step: int, lr=1e-2, beta1=0.9, beta2=0.99, eps=1e-15, loss_scale=128.0) -> None:
        lib().orc_adam_cpu(self.params.shape[0], self.params.ctypes.data, grads.ctypes.data, m1.ctypes.data, m2.ctypes.data, lr, beta1, beta2, eps, loss_scale, step)
