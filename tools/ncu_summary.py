#!/usr/bin/env python
"""Summarise ncu outputs for profiles/: a launch list (--metrics gpu__time_duration.sum --csv) and/or a
full report (.ncu-rep) -> compact text.  Usage:
    python tools/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep > profiles/rNN_x.txt
"""
import argparse
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v  # -> us
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# launch list ({path}): per-kernel device time, cold-cache / serialised -- compare SHARES")
    print(f"{'kernel':70s} {'n':>4s} {'avg_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:70]:70s} {len(v):4d} {sum(v)/len(v):10.1f} {100*sum(v)/tot:6.1f}%")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"\n# ncu --set full ({path})")
    for r in rows[2:]:
        print(f"\n## {r[ki][:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]:>18s} {units[i]}")
        try:
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            t = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            print(f"  {'traffic = dram read + write (bytes per launch)':72s} {t:18.0f}")
        except Exception:
            pass


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches)
    if a.rep:
        report(a.rep)
