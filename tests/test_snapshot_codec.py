"""CPU: the msgpack / gzip codec of the reference-compatible snapshot container (.ingp / .msgpack; Testbed::save_snapshot,
src/testbed.cu:5288-5355 writes nlohmann::json::to_msgpack through zstr's gzip stream) against the independent `msgpack` and
`gzip` Python modules, plus known-answer bytes of the MessagePack specification."""
import ctypes as C
import gzip
import json

import msgpack
import pytest

import util


@pytest.fixture(scope="module")
def lib():
    return util.pkg().load_library()


def to_msgpack(lib, obj, gz=False):
    text = json.dumps(obj).encode()
    buf = (C.c_uint8 * (len(text) * 2 + 4096))()
    n = C.c_size_t(0)
    assert lib.ngp_json_to_msgpack(text, int(gz), buf, len(buf), C.byref(n)) == 0, lib.ngp_last_error()
    return bytes(buf[: n.value])


def from_msgpack(lib, data, gz=False):
    out = C.create_string_buffer(len(data) * 8 + 4096)
    n = C.c_size_t(0)
    src = (C.c_uint8 * len(data)).from_buffer_copy(data)
    assert lib.ngp_msgpack_to_json(src, len(data), int(gz), out, len(out), C.byref(n)) == 0, lib.ngp_last_error()
    return json.loads(out.value.decode())


def test_known_answer_bytes(lib):
    # examples of the MessagePack specification / msgpack.org front page
    assert to_msgpack(lib, {"compact": True, "schema": 0}) == bytes.fromhex("82a7636f6d70616374c3a6736368656d6100")
    assert to_msgpack(lib, [1, -1, 127, 128, 255, 256, 65535, 65536, -32, -33, -128, -129]) == msgpack.packb([1, -1, 127, 128, 255, 256, 65535, 65536, -32, -33, -128, -129])
    assert to_msgpack(lib, None) == b"\xc0" and to_msgpack(lib, False) == b"\xc2"
    assert to_msgpack(lib, 0.5) == b"\xca\x3f\x00\x00\x00"                      # exactly a float32 -> float 32 (nlohmann's rule)
    assert to_msgpack(lib, 0.1) == b"\xcb" + bytes.fromhex("3fb999999999999a")  # otherwise float 64
    assert to_msgpack(lib, "a" * 40)[:2] == b"\xd9\x28"


CONFIG = {
    "loss": {"otype": "Huber"},
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20000, "decay_interval": 10000, "decay_base": 0.33,
                  "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
    "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16},
    "list": [[1.5, 2, -3], [], {}, "text", True, None, 4294967296, -2147483649],
}


def test_round_trip_against_python_msgpack(lib):
    packed = to_msgpack(lib, CONFIG)
    assert msgpack.unpackb(packed, raw=False) == CONFIG                     # our writer -> independent reader
    assert from_msgpack(lib, msgpack.packb(CONFIG, use_single_float=False)) == CONFIG   # independent writer -> our reader


def test_gzip_container_is_what_zstr_reads_and_writes(lib):
    z = to_msgpack(lib, CONFIG, gz=True)
    assert z[:2] == b"\x1f\x8b"                                              # gzip header (deflateInit2 window bits 15 + 16)
    assert msgpack.unpackb(gzip.decompress(z), raw=False) == CONFIG
    assert from_msgpack(lib, gzip.compress(msgpack.packb(CONFIG)), gz=True) == CONFIG
    import zlib
    assert from_msgpack(lib, zlib.compress(msgpack.packb(CONFIG)), gz=True) == CONFIG   # zstr::istream auto-detects zlib streams too


def test_binary_values_and_errors(lib):
    blob = msgpack.packb({"params_binary": b"\x00\x01" * 300, "n": 300})
    assert from_msgpack(lib, blob) == {"params_binary": {"bytes": 600}, "n": 300}
    out = C.create_string_buffer(64)
    n = C.c_size_t(0)
    bad = (C.c_uint8 * 3)(0x82, 0xA1, 0x61)
    assert lib.ngp_msgpack_to_json(bad, 3, 0, out, 64, C.byref(n)) != 0 and b"truncated" in lib.ngp_last_error()
    # a 5-byte header announcing 2^32 - 1 elements must not reserve memory for them, and nesting is bounded
    huge = (C.c_uint8 * 5)(0xDD, 0xFF, 0xFF, 0xFF, 0xFF)
    assert lib.ngp_msgpack_to_json(huge, 5, 0, out, 64, C.byref(n)) != 0 and b"truncated" in lib.ngp_last_error()
    deep = (C.c_uint8 * 200)(*([0x91] * 199 + [0xC0]))
    assert lib.ngp_msgpack_to_json(deep, 200, 0, out, 64, C.byref(n)) != 0 and b"nesting" in lib.ngp_last_error()


# ---- pinned against the reference's own serializer stack (oracle/ref/ref_snapshot_harness.cu: nlohmann::json, zstr,
# tcnn vec_json.h, json_binding.h compiled from /root/reference); goldens by tests/golden/make_ref_snapshot_golden.py
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
REF_EXE = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "ref_snapshot"


def fnv1a64(b: bytes) -> str:
    # 64-bit FNV-1a over the bytes, folded through numpy-free big-int arithmetic in 64 KiB pieces
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def strip(j):
    """{"bytes": n, "fnv1a64": h, "wsum64": w} -> {"bytes": n}: the product's JSON rendering of a binary value carries the size only"""
    if isinstance(j, dict):
        if set(j) == {"bytes", "fnv1a64", "wsum64"}:
            return {"bytes": j["bytes"]}
        return {k: strip(v) for k, v in j.items()}
    if isinstance(j, list):
        return [strip(v) for v in j]
    return j


def test_writer_is_byte_identical_to_nlohmann_to_msgpack(lib):
    doc = json.loads((GOLDEN / "ref_pack.json").read_text())
    want = (GOLDEN / "ref_pack.msgpack").read_bytes()
    got = to_msgpack(lib, doc)
    assert got == want
    assert from_msgpack(lib, want) == doc


def test_reads_a_snapshot_written_by_the_reference_serializer(lib):
    raw = (GOLDEN / "ref_snapshot.ingp").read_bytes()
    want = json.loads((GOLDEN / "ref_snapshot.dump.json").read_text())
    got = from_msgpack(lib, raw, gz=True)
    assert got == strip(want)
    # and the payload is what the harness says it wrote: closed-form fp16 patterns
    d = msgpack.unpackb(gzip.decompress(raw), raw=False)
    s = d["snapshot"]
    n = s["n_params"]
    i = np.arange(n, dtype=np.uint64)
    pattern = (((i * 37) % 1001).astype(np.float32) - 500.0) / 4000.0
    assert np.array_equal(np.frombuffer(s["params_binary"], dtype=np.float16), pattern.astype(np.float16))
    assert want["snapshot"]["params_binary"] == {"bytes": 2 * n, "fnv1a64": fnv1a64(s["params_binary"]), "wsum64": util.wsum64(s["params_binary"])}
    assert len(s["density_grid_binary"]) == 2 * 128 ** 3
    ds = s["nerf"]["dataset"]
    assert ds["n_images"] == 3 and ds["metadata"][1]["lens"] == {"is_fisheye": False, "k1": 0.0625, "k2": -0.03125, "p1": 0.001953125, "p2": -0.0009765625}
    assert ds["metadata"][0]["lens"] is None or ds["metadata"][0]["lens"] == {}   # perspective: to_json(Lens) assigns nothing


@pytest.mark.skipif(not REF_EXE.exists(), reason="oracle/_ref/ref_snapshot not built (make -C oracle/ref snapshot; needs /root/reference)")
def test_live_reference_serializer_agrees_on_fresh_documents(lib, tmp_path):
    import subprocess

    rng = np.random.default_rng(5)
    doc = {"f32": [float(np.float32(x)) for x in rng.normal(size=64)], "f64": [float(x) for x in rng.normal(size=64)],
           "i": [int(x) for x in rng.integers(-2 ** 40, 2 ** 40, size=64)], "u": [int(x) for x in rng.integers(0, 2 ** 53, size=16)],   # the product's Json holds numbers as double: integers exact to 2^53
           "s": ["x" * int(k) for k in rng.integers(0, 400, size=16)] + ["tab\there", "caf\u00e9 \U0001F600", "q\"b\\"], "cfg": CONFIG}
    (tmp_path / "d.json").write_text(json.dumps(doc))
    subprocess.check_call([str(REF_EXE), "pack", str(tmp_path / "d.json"), str(tmp_path / "d.msgpack")])
    assert to_msgpack(lib, doc) == (tmp_path / "d.msgpack").read_bytes()
