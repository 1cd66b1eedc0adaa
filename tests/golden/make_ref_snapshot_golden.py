"""Regenerates tests/golden/ref_snapshot.ingp (+ ref_snapshot.dump.json, ref_pack.json / ref_pack.msgpack) by running oracle/_ref/ref_snapshot — the REFERENCE's
own snapshot serializer stack (nlohmann::json to_msgpack, zstr gzip stream, tcnn vec_json.h and the application's json_binding.h
converters, compiled from /root/reference by oracle/ref/Makefile) on this CPU.  Needs /root/reference; run in the build container:

    make -C oracle/ref snapshot && python tests/golden/make_ref_snapshot_golden.py

The network is configs/nerf/base.json's shape with a 2^12-entry hash grid (n_params is taken from the library's own descriptor so a
layout change shows up as a test failure, not as a silently different file).
"""
import ctypes as C
import importlib
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
exe = ROOT / "oracle" / "_ref" / "ref_snapshot"
if not exe.exists():
    sys.exit("oracle/_ref/ref_snapshot missing: run `make -C oracle/ref snapshot` (needs /root/reference)")
P = importlib.import_module("instant-ngp_b200")
S = importlib.import_module("instant-ngp_b200.synthetic")
LOG2_T = 12
lib = P.load_library()
g = P.GridDesc()
assert lib.ngp_grid_desc_init(C.byref(g), 16, 2, LOG2_T, 16, 0.0, 1) == 0
d = P.NerfDesc()
assert lib.ngp_nerf_desc_init(C.byref(d), C.byref(g), 1, 2) == 0
cfg_path = Path("/tmp/ref_snapshot_config.json")
cfg_path.write_text(json.dumps(S.base_config(16, 2, LOG2_T)))
here = Path(__file__).resolve().parent
out = here / "ref_snapshot.ingp"
subprocess.check_call([str(exe), "write", str(out), str(cfg_path), str(d.n_params), "1", "1"])
dump = subprocess.check_output([str(exe), "dump", str(out)])
(here / "ref_snapshot.dump.json").write_bytes(dump)
print("wrote", out, out.stat().st_size, "bytes; n_params", d.n_params)

# byte-level known answers of nlohmann::json::to_msgpack for the number / string / container encodings a snapshot can hold
PACK = {
    "ints": [0, 1, 127, 128, 255, 256, 65535, 65536, 4294967295, 4294967296, -1, -32, -33, -128, -129, -32768, -32769, -2147483648, -2147483649],
    "floats": [0.0, 0.5, -1.25, 0.1, 1e-15, 1e-06, 0.33, 0.95, 3.4028234663852886e38, 1e39, 0.00123, 1.0000000000000002],
    "strings": ["", "a", "a" * 31, "a" * 32, "a" * 255, "a" * 256, "Ema", "\u00e9\u00e8"],
    "nested": {"z": [], "a": {}, "m": [[1.5, 2, -3], [True, False, None]], "list16": list(range(16)), "map16": {f"k{i:02d}": i for i in range(16)}},
    "config": S.base_config(16, 2, 19),
}
(here / "ref_pack.json").write_text(json.dumps(PACK, indent=1))
subprocess.check_call([str(exe), "pack", str(here / "ref_pack.json"), str(here / "ref_pack.msgpack")])
print("wrote ref_pack.json / ref_pack.msgpack", (here / "ref_pack.msgpack").stat().st_size, "bytes")
