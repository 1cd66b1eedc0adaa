"""Regenerates tests/golden/ref_host.bin.gz by running oracle/_ref/ref_host — the REFERENCE's own NGP_HOST_DEVICE helpers
(compiled from /root/reference by oracle/ref/Makefile) on this CPU.  Needs /root/reference; run in the build container:

    make -C oracle/ref host && python tests/golden/make_ref_host_goldens.py
"""
import gzip
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
exe = ROOT / "oracle" / "_ref" / "ref_host"
if not exe.exists():
    sys.exit("oracle/_ref/ref_host missing: run `make -C oracle/ref host` (needs /root/reference)")
tmp = Path("/tmp/ref_host.bin")
subprocess.check_call([str(exe), str(tmp)])
out = Path(__file__).resolve().parent / "ref_host.bin.gz"
with gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as f:
    f.write(tmp.read_bytes())
print("wrote", out, out.stat().st_size, "bytes")
