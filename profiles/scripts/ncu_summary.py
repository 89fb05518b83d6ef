#!/usr/bin/env python3
"""Summarise an Nsight Compute report: one line per profiled launch with duration, DRAM bytes / throughput, tensor-pipe,
ALU-pipe and issue utilisation.  Usage:  python profiles/scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
    ("launch__registers_per_thread", "regs"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = []
    for name, short in WANT:
        hit = [i for i, h in enumerate(hdr) if h == name or h.endswith("." + name)]
        if hit:
            cols.append((hit[0], short, units[hit[0]]))
    k = hdr.index("Kernel Name")
    g, b = hdr.index("Grid Size"), hdr.index("Block Size")
    print("# %s" % path)
    print("# columns: kernel | grid | block | " + " | ".join("%s[%s]" % (s, u) for _, s, u in cols))
    for r in rows[2:]:
        name = r[k].split("(")[0].replace("qgtc::<unnamed>::", "").replace("qgtc::", "")[:60]
        print(" | ".join([name, r[g], r[b]] + ["%s=%s" % (s, r[i]) for i, s, _ in cols]))


if __name__ == "__main__":
    main(sys.argv[1])
