#!/usr/bin/env python
"""Bring-up check of block 1 with the entry convolution fused in (conv0_path 2) against conv0_path 1 and the oracle.

    timeout 300 python tools/gpu_check_conv0_fused.py [--snippets 5]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import network_oracle  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--snippets", type=int, default=5)
    ap.add_argument("--time-seconds", type=float, default=600.0)
    args = ap.parse_args()
    P, S = runtime.bundled_parameters()
    ctx = runtime.get_context(P, S, 0)
    W = synthetic_weights(P, S, seed=1234)
    ctx.load_weights(W)
    x = np.random.default_rng(5).random((args.snippets, 736, 171), dtype=np.float32)
    ref, inter = network_oracle.forward(x, W, return_intermediates=True)
    want = np.maximum(np.transpose(inter["block1"], (0, 2, 3, 1)), 0)
    want_sub = np.transpose(inter["block1"], (0, 2, 3, 1))[:, ::2, ::2]
    ctx.set_option("net_path", 3)
    ok = True
    outs = {}
    for path in (1, 2):
        ctx.set_option("conv0_path", path)
        for stage, w, name in ((1, want, "block1 relu"), (21, want_sub, "block1 sub")):
            got = ctx.debug_stage(x, stage)
            err = np.abs(got - w)
            scale = np.abs(w).max()
            print(f"[conv0_path {path}] {name:12s} max|ref| {scale:.3f} max err {np.nanmax(err):.3e} mean {np.nanmean(err):.3e} finite {np.isfinite(got).all()}", flush=True)
            if not np.isfinite(got).all() or err.max() > 0.02 * scale + 1e-2:
                ok = False
                e_hw = np.where(np.isfinite(err), err, 1e9).max(axis=(0, 3))
                hh, ww = np.where(e_hw > 0.02 * scale + 1e-2)
                print(f"      bad rows {np.unique(hh)[:24]} ... cols {np.unique(ww)[:24]} ({len(hh)} bad pixels of {e_hw.size})")
        out = ctx.forward_host(x)
        outs[path] = out
        e = np.abs(out - ref)
        print(f"[conv0_path {path}] probabilities max err {e.max():.3e} mean {e.mean():.3e}", flush=True)
        ok &= bool(e.max() < 5e-3)
        ctx.set_option("chunk", 2)
        np.testing.assert_array_equal(ctx.forward_host(x), out)
        ctx.set_option("chunk", 2048)
    # timing on a resident recording
    pcm = synth_pcm16(args.time_seconds, seed=20251018)
    ctx.calibrate()
    for path in (1, 2, 1, 2):
        ctx.set_option("conv0_path", path)
        ctx.predict_pcm(pcm)
        r = ctx.predict_pcm(pcm)
        t = ctx.timings(); tm = [round(v, 3) for v in t["net_stage_ms"][:6]] + [round(t.get("network_ms", 0), 3)]
        print(f"[conv0_path {path}] stage ms: {tm}", flush=True)
    ctx.set_option("conv0_path", 1)
    ctx.set_option("net_path", 0)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
