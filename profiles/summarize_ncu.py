#!/usr/bin/env python
"""Turn an Nsight Compute report (gpurun_out/*.ncu-rep, captured with `ncu --set full --clock-control none
--import-source on`) into the markdown table committed under profiles/.   usage: summarize_ncu.py report.ncu-rep > out.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time us", 1e-3),  # ns -> us (ncu prints us already for --page raw; scale fixed below)
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__shared_mem_per_block_dynamic", "dyn smem B", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 thr %", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %", 1),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
    ("dram__bytes_read.sum", "DRAM rd MB", 1),
    ("dram__bytes_write.sum", "DRAM wr MB", 1),
    ("smsp__inst_executed.sum", "warp insts", 1),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"ncu report `{rep}` ({len(rows) - 2} kernel launches)\n")
    print("| kernel | " + " | ".join(m[1] for m in METRICS) + " |")
    print("|---|" + "---|" * len(METRICS))
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "")
        cells = []
        for key, label, _ in METRICS:
            if key in hdr:
                i = hdr.index(key)
                v, u = r[i], units[i]
                try:
                    f = float(v.replace(",", ""))
                    if label == "time us":
                        f = f if u in ("us", "usecond") else (f / 1e3 if u.startswith("n") else f * 1e3 if u.startswith("m") else f)
                    if label.endswith("MB"):
                        f = {"Mbyte": f, "Kbyte": f / 1e3, "Gbyte": f * 1e3, "byte": f / 1e6}.get(u, f)
                    cells.append(f"{f:.4g}")
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
