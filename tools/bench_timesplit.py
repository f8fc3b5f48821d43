#!/usr/bin/env python
"""ONE long recording annotated by several GPUs of one box (time chunks with halos, orcai_b200/timesplit.py) against the
same recording on one GPU.  One process, one context per device.

    python tools/bench_timesplit.py --hours 4 --devices 0,1 [--steps 3]

Prints one JSON line: seconds and hours of audio per second for both, speed-up, and whether the results are identical.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402  (make_recording: the bench's synthetic audio)
from orcai_b200._lib import Context  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.timesplit import plan_chunks, predict_pcm_timesplit  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=4.0)
    ap.add_argument("--devices", default="0,1")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tile-hours", type=float, default=0.0, help="generate this many hours and repeat them (a 24-h recording takes minutes to synthesise)")
    a = ap.parse_args()
    devs = [int(x) for x in a.devices.split(",")]
    P, S = runtime.bundled_parameters()
    W = synthetic_weights(P, S, seed=1234)
    ctxs = []
    for d in devs:
        c = runtime.get_context(P, S, d) if d == devs[0] else Context(P, S, device=d)
        c.load_weights(W)
        c.set_option("net_path", 3)
        c.calibrate()
        ctxs.append(c)
    if a.tile_hours > 0:
        unit = bench.make_recording(a.tile_hours, 20251018)
        raw = np.tile(unit, int(np.ceil(a.hours / a.tile_hours)))[: int(a.hours * 3600 * 48000)]
    else:
        raw = bench.make_recording(a.hours, 20251018)
    pcm = torch.from_numpy(raw).pin_memory().numpy()
    del raw

    def timed(fn):
        fn()
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            out = fn()
        for d in devs:
            torch.cuda.synchronize(d)
        return (time.perf_counter() - t0) / a.steps, out

    t_one, one = timed(lambda: ctxs[0].predict_pcm(pcm))
    t_split, split = timed(lambda: predict_pcm_timesplit(ctxs, pcm))
    same = all(np.array_equal(x, y) for x, y in zip(one[1:], split[1:])) and (one[0].lo, one[0].hi, one[0].db_ref) == (split[0].lo, split[0].hi, split[0].db_ref)
    chunks = plan_chunks(pcm.size, len(devs))
    print(json.dumps({
        "workload": f"orcai predict on ONE synthetic {a.hours:g}-hour recording" + (f" ({a.tile_hours:g} h of audio repeated)" if a.tile_hours > 0 else "") + ", host buffers -> segments (upload inside the timed region)",
        "devices": devs, "chunks": [{"snippets": c.n_snippets, "samples": c.sample1 - c.sample0} for c in chunks],
        "one_gpu": {"seconds": t_one, "h_audio_per_s": a.hours / t_one},
        "time_split": {"seconds": t_split, "h_audio_per_s": a.hours / t_split},
        "speedup": t_one / t_split, "results_identical": bool(same), "segments": int(len(one[3])),
    }))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
