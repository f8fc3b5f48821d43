#!/usr/bin/env python
"""Print the hand-off timeline recorded by a -DORCAI_FUSED_TRACE build (net_fused.cuh: FB_TRACE).

    ORCAI_B200_NVCC_EXTRA=-DORCAI_FUSED_TRACE python -m orcai_b200.build --force
    ORCAI_B200_TRACE=/tmp/trace.txt python tools/bringup/trace_run.py ; python tools/bringup/trace_timeline.py /tmp/trace.txt
"""
import sys
from collections import defaultdict

NAMES = {1: "issuer: X landed, start conv1", 2: "issuer: conv1 issued", 3: "issuer: start conv2 (waits s1_full)", 4: "issuer: conv2 issued",
         5: "producer: X free, TMA issued", 20: "workers: epi1 done", 21: "workers: pool+store done", 22: "workers: carry S2 done, wait conv2",
         19: "workers: epi1 stores done (warp 0), sync", 23: "workers: depthwise -> D2 done (warp 0)", 39: "workers: epi2 drained (warp 0)", 40: "workers: all drained (sync)", 41: "workers: S1 carry done = step end"}
for t in range(8):
    NAMES[10 + t] = f"workers: conv1 tile {t} complete"
    NAMES[30 + t] = f"workers: conv2 tile {t} complete"

UF = {1: "issuer: halo free, TMA issued", 2: "issuer: A planes full", 4: "issuer: accumulator drained, MMAs start", 3: "issuer: MMAs committed",
      10: "workers: halo landed, depthwise starts", 11: "workers: depthwise done (A planes written)", 12: "workers: accumulator complete, epilogue starts",
      13: "workers: epilogue done"}

runs = [[]]
for line in open(sys.argv[1]):
    a, b = map(int, line.split())
    if a < 0:
        runs.append([])
    else:
        runs[-1].append((a, b))
run = [r for r in runs if r][-1]          # the last forward
by_kernel = defaultdict(list)
for tag, clk in run:
    by_kernel[tag >> 32].append(((tag >> 8) & 0xFFFFFF, tag & 0xFF, clk))
for k, ev in sorted(by_kernel.items()):
    ev.sort(key=lambda e: e[2])
    t0 = ev[0][2]
    print(f"===== block with CIN = {k}: {len(ev)} events, span {ev[-1][2] - t0} cycles =====")
    for g, tag, clk in ev:
        print(f"  {clk - t0:8d}  step {g:3d}  {(UF if k == 70 else NAMES).get(tag, tag)}")
    ends = [clk for g, tag, clk in ev if tag == (13 if k == 70 else 41)]
    if len(ends) > 1:
        print("  step period (cycles):", [b - a for a, b in zip(ends, ends[1:])])
