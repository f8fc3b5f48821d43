#!/usr/bin/env python
"""Bring-up check of the tensor-core network path (tcgen05) against the oracle, stage by stage.

    timeout 300 python tools/gpu_check_tc.py [--fmt 1|2] [--snippets 2]
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import network_oracle  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402

STAGES = [(0, "conv0"), (1, "block1"), (2, "block2"), (3, "block3"), (4, "block4"), (5, "final")]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--snippets", type=int, default=2)
    ap.add_argument("--fmts", default="1,2")
    args = ap.parse_args()
    P, S = runtime.bundled_parameters()
    ctx = runtime.get_context(P, S, 0)
    W = synthetic_weights(P, S, seed=1234)
    ctx.load_weights(W)
    x = np.random.default_rng(5).random((args.snippets, 736, 171), dtype=np.float32)
    ref, inter = network_oracle.forward(x, W, return_intermediates=True)
    ok = True
    for fmt in [int(v) for v in args.fmts.split(",")]:
        name = {0: "fp32", 1: "fp16-tc", 2: "bf16-tc"}[fmt]
        ctx.set_option("net_path", fmt)
        for stage, key in STAGES:
            t0 = time.time()
            got = ctx.debug_stage(x, stage)
            want = np.transpose(inter[key], (0, 2, 3, 1))  # NCHW -> NHWC
            err = np.abs(got - want)
            scale = np.abs(want).max()
            print(f"[{name}] stage {key:7s} shape {got.shape} max|ref| {scale:8.4f} max err {err.max():.3e} mean err {err.mean():.3e} ({time.time() - t0:.2f}s)", flush=True)
            if not np.isfinite(got).all() or err.max() > 0.05 * scale + 1e-2:
                ok = False
                bad = np.unravel_index(np.argmax(err), err.shape)
                print(f"      worst at {bad}: got {got[bad]} want {want[bad]}")
        out = ctx.forward_host(x)
        e = np.abs(out - ref)
        print(f"[{name}] probabilities max err {e.max():.3e} mean {e.mean():.3e}", flush=True)
        ok &= bool(e.max() < 2e-2)
    ctx.set_option("net_path", 0)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
