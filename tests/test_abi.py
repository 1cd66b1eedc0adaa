"""CPU: the C-ABI library builds, loads and exports every symbol include/ngp_b200.h declares; host-only entry points work;
compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ngp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(built_lib, ngp):
    from importlib import import_module

    binding = import_module("instant-ngp_b200.binding")
    syms = declared_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in ngp_b200.h but not exported"
        assert s in binding.PROTOTYPES, f"{s} has no ctypes prototype"
    assert built_lib.ngp_version() == 1


def test_struct_sizes_match_header(built_lib, ngp):
    # compile a tiny C program against the header and compare sizeof()
    import subprocess, tempfile

    src = r'''
#include "ngp_b200.h"
#include <stdio.h>
int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ngp_grid_desc), sizeof(ngp_nerf_desc), sizeof(ngp_train_view), sizeof(ngp_march_consts),
 sizeof(ngp_nerf_train_cfg), sizeof(ngp_nerf_counters), sizeof(ngp_adam_cfg), sizeof(ngp_render_cfg), sizeof(ngp_field_desc), sizeof(ngp_tonemap_cfg)); return 0;}
'''
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "s.c"
        p.write_text(src)
        exe = Path(td) / "s"
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(p), "-o", str(exe)])
        sizes = list(map(int, subprocess.check_output([str(exe)]).split()))
    got = [C.sizeof(x) for x in (ngp.GridDesc, ngp.NerfDesc, ngp.TrainView, ngp.MarchConsts, ngp.NerfTrainCfg, ngp.NerfCounters, ngp.AdamCfg, ngp.RenderCfg, ngp.FieldDesc, ngp.TonemapCfg)]
    assert got == sizes


def test_grid_descriptor_matches_reference_known_answers(built_lib, ngp):
    # tiny-cuda-nn/tests/test_grid.cu:57-71
    g = ngp.GridDesc()
    assert built_lib.ngp_grid_desc_init(C.byref(g), 20, 2, 16, 32, 1.5, 1) == 0
    assert g.offsets[1] - g.offsets[0] == 32 * 32 * 32
    assert g.offsets[0] == 0
    assert g.offsets[2] - g.offsets[1] == 65536
    assert g.offsets[1] == 32 * 32 * 32
    assert g.offsets[3] - g.offsets[2] == 65536
    assert g.offsets[2] == 32 * 32 * 32 + 65536
    assert g.n_params == 2555904


def test_compute_calls_fail_loudly_without_a_device(built_lib, ngp):
    if built_lib.ngp_device_count() > 0:
        pytest.skip("a device is present")
    g = ngp.GridDesc()
    assert built_lib.ngp_grid_desc_init(C.byref(g), 16, 2, 19, 16, 0.0, 4) == 0
    d = ngp.NerfDesc()
    assert built_lib.ngp_nerf_desc_init(C.byref(d), C.byref(g), 1, 2) == 0
    rc = built_lib.ngp_nerf_inference(C.byref(d), None, 256, None, None, None, 4)
    assert rc != 0
    assert b"no CPU fallback" in built_lib.ngp_last_error()
    assert built_lib.ngp_testbed_create(0, None) is None
    with pytest.raises(ngp.NgpError):
        ngp.Testbed()


def test_unsupported_configs_are_rejected(built_lib, ngp):
    g = ngp.GridDesc()
    assert built_lib.ngp_grid_desc_init(C.byref(g), 16, 2, 19, 16, 0.0, 1) == 0
    d = ngp.NerfDesc()
    assert built_lib.ngp_nerf_desc_init(C.byref(d), C.byref(g), 0, 2) != 0
    g2 = ngp.GridDesc()
    assert built_lib.ngp_grid_desc_init(C.byref(g2), 12, 2, 19, 16, 0.0, 1) == 0
    assert built_lib.ngp_nerf_desc_init(C.byref(d), C.byref(g2), 1, 2) != 0  # 24-wide encoding is not fused here


def test_param_init_layout_and_ranges(built_lib, ngp):
    g = ngp.GridDesc()
    built_lib.ngp_grid_desc_init(C.byref(g), 16, 2, 15, 16, 0.0, 1)
    d = ngp.NerfDesc()
    built_lib.ngp_nerf_desc_init(C.byref(d), C.byref(g), 1, 2)
    p = np.zeros(d.n_params, dtype=np.float32)
    assert built_lib.ngp_nerf_init_params_host(C.byref(d), 1337, p.ctypes.data) == 0
    mlp, grid = p[: d.n_mlp_params], p[d.n_mlp_params:]
    assert d.n_mlp_params == 3072 + 7168
    s0 = np.sqrt(6.0 / (64 + 32))
    assert np.abs(mlp[:2048]).max() <= s0 and np.abs(mlp[:2048]).max() > 0.9 * s0
    assert np.abs(grid).max() <= 1e-4 and np.abs(grid).min() >= 0 and grid.std() > 3e-5


def test_library_descriptor_equals_oracle_layout_bit_for_bit(built_lib, ngp):
    import util

    for kw in (dict(n_levels=16, F=2, aabb_scale=4), dict(n_levels=8, F=4, aabb_scale=4), dict(n_levels=16, F=2, aabb_scale=1), dict(n_levels=16, F=2, aabb_scale=16)):
        d, L = util.make_desc(**kw)
        g = d.grid
        assert list(g.offsets[: g.n_levels + 1]) == L.grid.offsets
        assert list(g.resolutions[: g.n_levels]) == L.grid.resolutions
        assert np.array(list(g.scales[: g.n_levels]), dtype=np.float32).tobytes() == np.array(L.grid.scales, dtype=np.float32).tobytes()
        assert d.n_params == L.n_params and d.n_mlp_params == L.n_mlp_params
