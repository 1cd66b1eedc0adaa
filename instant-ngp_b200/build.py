"""In-tree build of libngp_b200.so (sm_100a) — explicit nvcc commands, no JIT cache.

The shared library is the product: hand-written CUDA kernels + the C++ host Testbed behind the C-ABI of
include/ngp_b200.h.  Objects and the .so stay in instant-ngp_b200/_build/ and instant-ngp_b200/ (git-ignored, but they
travel to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libngp_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-I", str(ROOT / "include")]

# translation unit -> extra flags.  march/render are compiled without FMA contraction so that their arithmetic is
# bit-identical to the CPU oracle (see include/ngp_detmath.h).
UNITS = {
    "nerf_net.cu": [],
    "optimizer.cu": ["--use_fast_math"],
    "testbed.cu": ["-fmad=false"],
    "march.cu": ["-fmad=false"],
    "march_ref.cu": ["--use_fast_math"],   # the reference build's flags (CMakeLists.txt:88): see march_ref.cu
    "render.cu": ["-fmad=false"],
    "field.cu": ["-fmad=false"],
}


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: libngp_b200 cannot be built")
    return nv


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((ROOT / "include").glob("*.h"))
    nvcc = _nvcc()
    jobs = []
    objs = []
    for unit, extra in UNITS.items():
        src = CSRC / unit
        obj = BUILD / (unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append([nvcc, *ARCH, *COMMON, *extra, "-c", str(src), "-o", str(obj)])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _newer(LIB, objs):
        run([nvcc, *ARCH, "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", str(LIB), *map(str, objs), "-lz"])
    return LIB


def build_oracle(verbose: bool = False) -> Path:
    """Compile the CPU oracle (test infrastructure, never on the product path)."""
    odir = ROOT / "oracle"
    out = odir / "libngp_oracle.so"
    src = odir / "ngp_oracle.c"
    deps = [src, ROOT / "include" / "ngp_detmath.h", ROOT / "include" / "ngp_b200.h"]
    if not src.exists():
        raise RuntimeError(f"{src} is missing")
    if _newer(out, deps):
        cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-I", str(ROOT / "include"), str(src), "-o", str(out), "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed building the oracle:\n" + r.stdout + r.stderr)
    return out


def build_reference(verbose: bool = False) -> bool:
    """oracle/_ref: the reference's own sources compiled where they lie (only when /root/reference exists)."""
    if not Path("/root/reference").exists():
        return False
    targets = ["tcnn"]
    if (ROOT / "oracle" / "ref" / "ref_tcnn_harness.cu").exists():
        targets.append("harness")
    if (ROOT / "oracle" / "ref" / "ref_host_harness.cu").exists():
        targets.append("host")
    if (ROOT / "oracle" / "ref" / "ref_snapshot_harness.cu").exists():
        targets.append("snapshot")
    if (ROOT / "oracle" / "ref" / "ref_loader_harness.cu").exists():
        targets.append("loader")
    if (ROOT / "oracle" / "ref" / "ref_exr_harness.cpp").exists():
        targets.append("exr")
    if (ROOT / "oracle" / "ref" / "ref_nerf_harness.cu").exists():
        targets.append("nerf")
    r = subprocess.run(["make", "-C", str(ROOT / "oracle" / "ref"), "-j8", *targets], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/_ref failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return True


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built", p)
