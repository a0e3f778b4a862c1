#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into a small markdown table of the metrics the roofline needs.
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "dur_us"), ("launch__grid_size", "grid"),
        ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_rd"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__cycles_active.avg", "sm_cycles"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"ncu --set full --clock-control none, source: {rep}\n")
    print("| # | kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|" + "---|" * len(WANT))
    for k, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
        cells = []
        for m, _ in WANT:
            i = col.get(m)
            cells.append("-" if i is None else f"{r[i]} {units[i]}".strip())
        print(f"| {k} | `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
