#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under
profiles/ (the .ncu-rep files themselves stay in gpurun_out/, which is scratch).

    python profiles/summarize_ncu.py launches gpurun_out/r01_launches_dna_m32.csv
    python profiles/summarize_ncu.py full gpurun_out/r01_scan_dna.ncu-rep
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg.per_second",
]
STALLS = "smsp__average_warps_issue_stalled_"


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    h = rows[0]
    ik, im, iv, ii, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
    per = OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[ii], {"kernel": r[ik]})[r[im]] = (float(r[iv].replace(",", "")), r[iu])
    print(f"# per-launch list from {path} (ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none;")
    print("# serialised, cold-cache launches: compare SHARES, not absolutes)")
    print(f"{'id':>4} {'kernel':46s} {'time_us':>10} {'dram_read_MB':>13} {'dram_write_MB':>14}")
    agg = OrderedDict()
    for i, d in per.items():
        t = d.get("gpu__time_duration.sum", (0, ""))
        t_us = t[0] / 1e3 if t[1] in ("ns", "nsecond") else t[0]
        rd = d.get("dram__bytes_read.sum", (0, ""))[0] / 1e6
        wr = d.get("dram__bytes_write.sum", (0, ""))[0] / 1e6
        name = d["kernel"][:46]
        print(f"{i:>4} {name:46s} {t_us:10.1f} {rd:13.2f} {wr:14.2f}")
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t_us; a[2] += rd; a[3] += wr
    tot = sum(a[1] for a in agg.values())
    print("\n# totals")
    for name, a in agg.items():
        print(f"{name:46s} launches={a[0]:<3d} time_us={a[1]:10.1f} share={a[1] / tot:6.1%} avg_us={a[1] / a[0]:9.1f} "
              f"avg_dram_read_MB={a[2] / a[0]:10.2f} avg_dram_write_MB={a[3] / a[0]:10.2f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none, summary of {path}")
    for r in rows[2:]:
        print(f"\n== {r[h.index('Kernel Name')]}  (id {r[h.index('ID')]})")
        for name, u, v in zip(h, units, r):
            if name in KEEP:
                print(f"{name:75s} {v:>16s} {u}")
        print("-- warp stall reasons (warps per issue-active cycle)")
        st = [(float(v), name[len(STALLS):].replace("_per_issue_active.ratio", "")) for name, v in zip(h, r)
              if name.startswith(STALLS) and name.endswith("_per_issue_active.ratio") and "not_issued" not in name]
        for v, name in sorted(st, reverse=True)[:8]:
            print(f"   {name:28s} {v:8.3f}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
