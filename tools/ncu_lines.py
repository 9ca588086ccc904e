#!/usr/bin/env python3
"""Aggregates `ncu --page source --print-source cuda,sass --csv` output per CUDA source line:
share of executed warp instructions, average active threads, stall samples.
usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = {}
fname = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, te, ns = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= te or r[2] != "-":
        continue  # keep only the per-source-line summary rows (Address == "-")
    try:
        a, b, s = float(r[ie] or 0), float(r[te] or 0), float(r[ns] or 0)
    except ValueError:
        continue
    k = (fname, r[0])
    v = agg.setdefault(k, [0, 0, 0, r[1]])
    v[0] += a; v[1] += b; v[2] += s
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
print(f"total warp instructions {tot:.3e}, samples {tots:.0f}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if v[0] == 0:
        continue
    print(f"{v[0]/tot*100:5.2f}% inst {v[2]/tots*100:5.2f}% smp  thr {v[1]/v[0]:4.1f}  {f}:{l:>5}  {v[3][:100]}")
