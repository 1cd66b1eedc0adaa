#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libngp_b200.so (cuobjdump -sass), the evidence for which hardware paths the product's kernels use:
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG / UTMASTG = TMA bulk tensor copies, REDG / ATOMG = global reductions /
atomics, MUFU = SFU.   usage: python tools/sass_histogram.py [library] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS", "LDG", "STG", "LDS", "STS", "MUFU", "HFMA2", "FFMA", "SHFL", "VOTE",
       "BAR", "WARPSYNC"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "instant-ngp_b200" / "libngp_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# SASS opcode histogram of {Path(lib).name} (sm_100a), {len(kernels)} kernels; columns: total instructions, then the opcodes that matter")
    print("# " + " ".join(KEY))
    for (name, c), dn in zip(kernels.items(), demangled):
        total.update(c)
        short = re.sub(r"\(.*", "", dn).replace("void ", "")
        cells = " ".join(f"{k}={c[k]}" for k in KEY if c[k])
        print(f"{short}: n={sum(c.values())} {cells}")
    print("# library totals: " + " ".join(f"{k}={total[k]}" for k in KEY if total[k]))


if __name__ == "__main__":
    main()
