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
