"""Regenerates tests/golden/exr/*: OpenEXR files written by the REFERENCE's codec (tinyexr, compiled from /root/reference into
oracle/_ref/ref_exr by `make -C oracle/ref exr`) and tinyexr's own decoding of each (`<name>.bin`: int32 w, int32 h, float RGBA).
Needs /root/reference; run in the build container:

    make -C oracle/ref exr && python tests/golden/make_ref_exr_goldens.py
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
exe = ROOT / "oracle" / "_ref" / "ref_exr"
if not exe.exists():
    sys.exit("oracle/_ref/ref_exr missing: run `make -C oracle/ref exr` (needs /root/reference)")
out = Path(__file__).resolve().parent / "exr"
out.mkdir(exist_ok=True)
# (name, width, height, channels, half, compression): sizes that leave partial 16-line blocks and odd byte counts
CASES = [("rgba_half_zip", 37, 23, 4, 1, 3), ("rgba_float_zip", 37, 23, 4, 0, 3), ("rgb_half_zips", 21, 9, 3, 1, 2), ("rgba_float_none", 16, 5, 4, 0, 0),
         ("rgb_float_rle", 19, 7, 3, 0, 1), ("gray_half_zip", 33, 17, 1, 1, 3), ("rgba_half_zip_64", 64, 64, 4, 1, 3)]
for name, w, h, c, half, comp in CASES:
    subprocess.check_call([str(exe), "encode", str(out / f"{name}.exr"), str(w), str(h), str(c), str(half), str(comp)])
    subprocess.check_call([str(exe), "decode", str(out / f"{name}.exr"), str(out / f"{name}.bin")])
    print(name, (out / f"{name}.exr").stat().st_size, "bytes")
