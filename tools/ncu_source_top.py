#!/usr/bin/env python
"""Rank the source lines of ONE kernel of an ncu report (captured with --import-source on, built with -lineinfo) by shared-memory
bank conflicts, excessive shared wavefronts and warp-stall samples; run on the GPU box, where the .ncu-rep lives:

    python tools/ncu_source_top.py /tmp/kernel.ncu-rep > gpurun_out/kernel_source_top.txt
"""
import csv, sys, io, subprocess
rep = sys.argv[1]
r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True)
txt = r.stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
for i, row in enumerate(rows):
    if any("Source" == c or c.startswith("Source") for c in row) and len(row) > 5:
        hdr = i; break
if hdr is None:
    print("no header", txt[:2000]); sys.exit(0)
h = rows[hdr]
print("COLUMNS:", h)
def col(pat):
    for i, c in enumerate(h):
        if pat.lower() in c.lower(): return i
    return None
ci = {k: col(k) for k in ["Source", "L1 Conflicts Shared N-Way", "Shared Bank Conflicts", "L1 Wavefronts Shared", "Warp Stall Sampling (All", "# Samples", "Instructions Executed"]}
print(ci)
key = ci.get("L1 Conflicts Shared N-Way") or ci.get("Shared Bank Conflicts")
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
data = rows[hdr + 1:]
for name, k in (("conflicts", key), ("wavefronts", ci.get("L1 Wavefronts Shared")), ("stall samples", ci.get("Warp Stall Sampling (All") or ci.get("# Samples"))):
    if k is None: continue
    top = sorted(data, key=lambda r: -num(r[k]) if len(r) > k else 0)[:25]
    print(f"\n== top by {name} ({h[k]})")
    for r_ in top:
        print(f"{num(r_[k]):14.0f}  " + " | ".join(r_[j][:90] for j in range(min(3, len(r_)))))
