#!/usr/bin/env python
"""Per-SASS-instruction hot spots from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K`.
Prints the instructions with the most executions and the most stall samples."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot_inst = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(float(r[col["# Samples"]] or 0) for r in data)
print(f"{rows[0][1][:100]}\ninstructions executed (warp-level) {tot_inst:.0f}, samples {tot_samp:.0f}, SASS lines {len(data)}")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(float(r[col[s]] or 0) for r in data) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]}={100*v/max(tot_samp,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("\n-- by executions")
for r in sorted(data, key=lambda r: -float(r[col["Instructions Executed"]] or 0))[:top]:
    print(f"{float(r[col['Instructions Executed']])/tot_inst*100:5.2f}% exec {float(r[col['# Samples']] or 0)/max(tot_samp,1)*100:5.2f}% samp  {r[col['Source']][:90]}")
