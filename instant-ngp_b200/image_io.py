"""Image files for the image mode's training data (Testbed::load_image, src/testbed_image.cu:393-457): OpenEXR scanline files,
the reference's `.bin` dump, and 8-bit formats.

The reference decodes EXR with tinyexr (`load_exr` -> `LoadEXRFromMemory`, src/tinyexr_wrapper.cu:119-143) and everything else with
stb_image; this image has neither, so the scanline EXR container is read here (numpy + zlib: NONE / RLE / ZIPS / ZIP compression, HALF
and FLOAT channels) and 8-bit files go through PIL followed by the arithmetic of `from_rgba32<float>` (common_device.cuh:698-735).
Pinned against tinyexr itself: oracle/ref/ref_exr_harness.cpp writes the files under tests/golden/exr/ and its own decoding of them
(tests/test_image_io.py).  Not read: tiled, multipart and deep files, PIZ / PXR24 / B44 / DWA compression, subsampled channels."""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

_LINES_PER_BLOCK = {0: 1, 1: 1, 2: 1, 3: 16}          # NONE, RLE, ZIPS, ZIP
_PIXEL_DTYPE = {1: np.dtype("<f2"), 2: np.dtype("<f4")}


class ExrError(ValueError):
    pass


def _cstr(b: bytes, p: int):
    e = b.index(b"\0", p)
    return b[p:e], e + 1


def _parse_header(b: bytes):
    if len(b) < 8 or b[:4] != b"\x76\x2f\x31\x01":
        raise ExrError("not an OpenEXR file")
    version, = struct.unpack_from("<I", b, 4)
    if (version & 0xFF) != 2:
        raise ExrError("unsupported OpenEXR version")
    if version & 0x200:
        raise ExrError("tiled EXR files are not supported")
    if version & 0x1800:
        raise ExrError("EXR file must be singlepart")                  # the reference's own message (tinyexr_wrapper.cu:162-164)
    attrs = {}
    p = 8
    while b[p] != 0:
        name, p = _cstr(b, p)
        typ, p = _cstr(b, p)
        size, = struct.unpack_from("<I", b, p)
        p += 4
        attrs[name.decode()] = (typ.decode(), b[p:p + size])
        p += size
    return attrs, p + 1


def _channels(v: bytes):
    out = []
    p = 0
    while v[p] != 0:
        name, p = _cstr(v, p)
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", v, p)
        p += 16
        out.append((name.decode(), ptype, xs, ys))
    return out


def _unrle(src: bytes, n_out: int) -> bytes:
    out = bytearray()
    p = 0
    while p < len(src):
        c = struct.unpack_from("b", src, p)[0]
        p += 1
        if c < 0:
            out += src[p:p - c]
            p += -c
        else:
            out += src[p:p + 1] * (c + 1)
            p += 1
    if len(out) != n_out:
        raise ExrError("corrupt RLE block")
    return bytes(out)


def _unpredict(t: bytes) -> np.ndarray:
    """undo the byte predictor (d[i] += d[i-1] - 128) and the even / odd byte split of the ZIP and RLE codecs"""
    a = np.frombuffer(t, dtype=np.uint8).astype(np.int64)
    a[1:] -= 128
    a = (np.cumsum(a) & 0xFF).astype(np.uint8)
    half = (a.size + 1) // 2
    out = np.empty(a.size, dtype=np.uint8)
    out[0::2] = a[:half]
    out[1::2] = a[half:]
    return out


def read_exr(path) -> np.ndarray:
    """[H, W, 4] float32 RGBA as tinyexr's LoadEXRFromMemory assembles it: R, G, B required (A = 1 when absent); a single-channel file
    is replicated into all four"""
    b = Path(path).read_bytes()
    attrs, p = _parse_header(b)
    for need in ("channels", "compression", "dataWindow"):
        if need not in attrs:
            raise ExrError(f"EXR header lacks '{need}'")
    comp = attrs["compression"][1][0]
    if comp not in _LINES_PER_BLOCK:
        raise ExrError(f"EXR compression type {comp} is not supported (NONE, RLE, ZIPS, ZIP are)")
    xmin, ymin, xmax, ymax = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = xmax - xmin + 1, ymax - ymin + 1
    if w <= 0 or h <= 0:
        raise ExrError("empty data window")
    chans = _channels(attrs["channels"][1])
    for name, ptype, xs, ys in chans:
        if ptype not in _PIXEL_DTYPE:
            raise ExrError(f"channel '{name}': pixel type {ptype} is not supported (HALF and FLOAT are)")
        if xs != 1 or ys != 1:
            raise ExrError("subsampled channels are not supported")
    line_bytes = sum(w * _PIXEL_DTYPE[pt].itemsize for _, pt, _, _ in chans)
    lpb = _LINES_PER_BLOCK[comp]
    n_blocks = (h + lpb - 1) // lpb
    offsets = struct.unpack_from(f"<{n_blocks}Q", b, p)
    planes = {name: np.empty((h, w), dtype=np.float32) for name, _, _, _ in chans}
    for off in offsets:
        y, size = struct.unpack_from("<ii", b, off)
        y -= ymin
        if not (0 <= y < h):
            raise ExrError("scanline block outside the data window")
        lines = min(lpb, h - y)
        want = lines * line_bytes
        data = b[off + 8: off + 8 + size]
        if len(data) != size:
            raise ExrError("truncated EXR file")
        if comp == 0 or size == want:               # a block that does not shrink is stored raw
            raw = np.frombuffer(data, dtype=np.uint8)
        elif comp == 1:
            raw = _unpredict(_unrle(data, want))
        else:
            raw = _unpredict(zlib.decompress(data))
        if raw.size != want:
            raise ExrError("corrupt scanline block")
        rows = raw.reshape(lines, line_bytes)
        q = 0
        for name, ptype, _, _ in chans:             # within a scanline the channels follow each other in header (alphabetical) order
            dt = _PIXEL_DTYPE[ptype]
            nb = w * dt.itemsize
            planes[name][y:y + lines] = np.ascontiguousarray(rows[:, q:q + nb]).view(dt).reshape(lines, w).astype(np.float32)
            q += nb
    out = np.empty((h, w, 4), dtype=np.float32)
    if len(chans) == 1:
        out[...] = next(iter(planes.values()))[..., None]
        return out
    for k, name in enumerate("RGB"):
        if name not in planes:
            raise ExrError(f"{name} channel not found")
        out[..., k] = planes[name]
    out[..., 3] = planes["A"] if "A" in planes else 1.0
    return out


def read_bin(path) -> np.ndarray:
    """load_binary_image (src/testbed_image.cu:439-457): int32 height, int32 width, then half RGBA"""
    b = Path(path).read_bytes()
    h, w = struct.unpack_from("<ii", b, 0)
    if h <= 0 or w <= 0 or len(b) < 8 + h * w * 8:
        raise ValueError("binary image: bad header or truncated file")
    return np.frombuffer(b, dtype="<f2", count=h * w * 4, offset=8).reshape(h, w, 4).astype(np.float32)


def rgba32_to_linear_premultiplied(img_u8: np.ndarray) -> np.ndarray:
    """from_rgba32<float> (common_device.cuh:698-735): sRGB bytes -> linear colour multiplied by alpha, float32 arithmetic"""
    a = np.asarray(img_u8, dtype=np.uint8)
    s = a[..., :3].astype(np.float32) * np.float32(1.0 / 255.0)
    lin = np.where(s <= np.float32(0.04045), s / np.float32(12.92), np.power((s + np.float32(0.055)) / np.float32(1.055), np.float32(2.4))).astype(np.float32)
    alpha = a[..., 3:4].astype(np.float32) * np.float32(1.0 / 255.0)
    return np.concatenate([lin * alpha, alpha], axis=-1).astype(np.float32)


def read_ldr(path) -> np.ndarray:
    """load_stbi_gpu for 8-bit files (src/common_host.cu:258-290): decoded to RGBA bytes, then from_rgba32"""
    from PIL import Image

    with Image.open(path) as im:
        if im.size[0] == 0 or im.size[1] == 0:
            raise ValueError("Image has zero pixels.")
        return rgba32_to_linear_premultiplied(np.asarray(im.convert("RGBA"), dtype=np.uint8))


def load_image(path) -> np.ndarray:
    """Testbed::load_image: by extension — .exr, .bin, anything else as an 8-bit image; [H, W, 4] float32"""
    p = Path(path)
    if not p.exists():
        raise FileNotFoundError(f"Image file '{p}' does not exist.")
    ext = p.suffix.lower()
    if ext == ".exr":
        return read_exr(p)
    if ext == ".bin":
        return read_bin(p)
    return read_ldr(p)
