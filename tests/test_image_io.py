"""CPU: the image-mode file readers (instant-ngp_b200/image_io.py ≙ Testbed::load_image, src/testbed_image.cu:393-457) against the
reference's own EXR codec: tests/golden/exr/*.exr were written by tinyexr and *.bin hold tinyexr's decoding of them
(oracle/ref/ref_exr_harness.cpp, tests/golden/make_ref_exr_goldens.py)."""
import importlib
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

IO = importlib.import_module("instant-ngp_b200.image_io")
GOLD = Path(__file__).resolve().parent / "golden" / "exr"
REF_EXE = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "ref_exr"
ALBERT = Path("/root/reference/data/image/albert.exr")
CASES = sorted(p.stem for p in GOLD.glob("*.exr"))


def read_ref_bin(path):
    b = Path(path).read_bytes()
    w, h = struct.unpack_from("<ii", b, 0)
    return np.frombuffer(b, dtype="<f4", offset=8).reshape(h, w, 4)


def pattern(w, h, c):
    x, y = np.meshgrid(np.arange(w), np.arange(h), indexing="xy")
    return (((x * 7 + y * 13 + c * 29) % 97).astype(np.float32) / np.float32(32.0) - np.float32(0.75) + np.float32(1.0 if c == 3 else 0.0)).astype(np.float32)


def test_goldens_are_present():
    assert len(CASES) == 7


@pytest.mark.parametrize("name", CASES)
def test_exr_reader_matches_tinyexr_bit_for_bit(name):
    want = read_ref_bin(GOLD / f"{name}.bin")
    got = IO.read_exr(GOLD / f"{name}.exr")
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # and both are the closed-form pattern the harness encoded (rounded to half where the file stores halves)
    h, w = got.shape[:2]
    n_ch = 1 if name.startswith("gray") else (3 if name.startswith("rgb_") else 4)
    for k in range(4):
        src = pattern(w, h, 0 if n_ch == 1 else k)
        if "half" in name:
            src = src.astype(np.float16).astype(np.float32)
        if k == 3 and n_ch == 3:
            src = np.ones_like(src)
        assert np.array_equal(got[..., k], src), (name, k)


@pytest.mark.skipif(not (ALBERT.exists() and REF_EXE.exists()), reason="reference data / oracle/_ref/ref_exr not present (GPU box)")
def test_albert_exr_live(tmp_path):
    subprocess.check_call([str(REF_EXE), "decode", str(ALBERT), str(tmp_path / "a.bin")])
    want = read_ref_bin(tmp_path / "a.bin")
    got = IO.load_image(ALBERT)
    assert got.shape == (1024, 1024, 4) and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_bin_and_ldr_and_errors(tmp_path):
    from PIL import Image

    rng = np.random.default_rng(1)
    half = rng.uniform(0, 2, size=(5, 7, 4)).astype(np.float16)
    (tmp_path / "img.bin").write_bytes(struct.pack("<ii", 5, 7) + half.tobytes())          # height first (testbed_image.cu:447-448)
    assert np.array_equal(IO.load_image(tmp_path / "img.bin"), half.astype(np.float32))
    px = rng.integers(0, 256, size=(6, 4, 4), dtype=np.uint8)
    Image.fromarray(px, "RGBA").save(tmp_path / "img.png")
    got = IO.load_image(tmp_path / "img.png")
    s = px[..., :3].astype(np.float64) / 255.0
    lin = np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)
    a = px[..., 3:4].astype(np.float64) / 255.0
    assert got.shape == (6, 4, 4) and np.allclose(got[..., :3], lin * a, atol=2e-6) and np.allclose(got[..., 3:], a, atol=1e-7)
    with pytest.raises(FileNotFoundError, match="does not exist"):
        IO.load_image(tmp_path / "nothing.exr")
    (tmp_path / "bad.exr").write_bytes(b"not an exr file at all")
    with pytest.raises(IO.ExrError, match="not an OpenEXR"):
        IO.load_image(tmp_path / "bad.exr")
    good = (GOLD / "rgba_half_zip.exr").read_bytes()
    (tmp_path / "cut.exr").write_bytes(good[: len(good) - 40])
    with pytest.raises((IO.ExrError, Exception)):
        IO.read_exr(tmp_path / "cut.exr")
    # an unsupported codec is refused by name, not mis-decoded: flip the compression byte to PIZ
    i = good.index(b"compression\0compression\0") + len(b"compression\0compression\0") + 4
    (tmp_path / "piz.exr").write_bytes(good[:i] + b"\x04" + good[i + 1:])
    with pytest.raises(IO.ExrError, match="compression type 4"):
        IO.read_exr(tmp_path / "piz.exr")
