#!/usr/bin/env python
"""Bring-up check of the fused residual-block path (net_path 3) against the oracle, stage by stage.

    timeout 300 python tools/gpu_check_fused.py [--snippets 3]
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


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--snippets", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    P, S = runtime.bundled_parameters()
    ctx = runtime.get_context(P, S, 0)
    W = synthetic_weights(P, S, seed=args.seed)
    ctx.load_weights(W)
    x = np.random.default_rng(5).random((args.snippets, 736, 171), dtype=np.float32)
    ref, inter = network_oracle.forward(x, W, return_intermediates=True)
    nhwc = {k: np.transpose(v, (0, 2, 3, 1)) for k, v in inter.items() if v.ndim == 4}
    ok = True
    ctx.set_option("net_path", 3)
    stages = [(0, "conv0", nhwc["conv0"])]
    for b in (1, 2, 3):
        y = nhwc[f"block{b}"]
        stages.append((b, f"block{b} relu", np.maximum(y, 0)))
        stages.append((20 + b, f"block{b} sub", y[:, ::2, ::2]))
    stages.append((4, "block4", nhwc["block4"]))
    stages.append((5, "final", nhwc["final"]))
    for stage, key, want in stages:
        t0 = time.time()
        got = ctx.debug_stage(x, stage)
        if got.shape != want.shape:
            print(f"[fused] stage {key}: shape {got.shape} != {want.shape}")
            ok = False
            continue
        err = np.abs(got - want)
        scale = np.abs(want).max()
        fin = np.isfinite(got).all()
        print(f"[fused] stage {key:12s} shape {got.shape} max|ref| {scale:8.4f} max err {np.nanmax(err):.3e} mean err {np.nanmean(err):.3e} finite {fin} ({time.time() - t0:.2f}s)", flush=True)
        if not fin or err.max() > 0.05 * scale + 1e-2:
            ok = False
            bad = np.unravel_index(np.nanargmax(np.where(np.isfinite(err), err, np.inf)), err.shape)
            print(f"      worst at {bad}: got {got[bad]} want {want[bad]}")
            # error map by (h, w) position to spot tile-edge bugs
            e_hw = np.where(np.isfinite(err), err, 1e9).max(axis=(0, 3))
            hh, ww = np.where(e_hw > 0.05 * scale + 1e-2)
            print(f"      bad rows {np.unique(hh)[:24]} ... cols {np.unique(ww)[:24]} ({len(hh)} bad pixels of {e_hw.size})")
    for tail in (0, 1):
        ctx.set_option("tail_path", tail)
        out = ctx.forward_host(x)
        e = np.abs(out - ref)
        print(f"[fused] tail_path {tail}: probabilities max err {e.max():.3e} mean {e.mean():.3e}", flush=True)
        ok &= bool(e.max() < 5e-3)
    # weight-rounding bias calibration (built-in synthetic recording) and a realistic input: snippets of another synthetic recording
    from oracle import postprocess_oracle as po, spectrogram_oracle as so
    from orcai_b200.synth import pcm16_to_float, synth_pcm16
    pcm = synth_pcm16(40.0, seed=4242, calls_per_minute=20.0)
    spec, _, _ = so.make_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    xa = po.cut_snippets(spec, 736)[:8]
    refa = network_oracle.forward(xa, W)
    ctx.set_option("tail_path", 1)
    ea = np.abs(ctx.forward_host(xa) - refa)
    print(f"[fused] audio snippets, uncalibrated: max err {ea.max():.3e} mean {ea.mean():.3e}", flush=True)
    ctx.calibrate()
    ea = np.abs(ctx.forward_host(xa) - refa)
    e = np.abs(ctx.forward_host(x) - ref)
    print(f"[fused] audio snippets, calibrated:   max err {ea.max():.3e} mean {ea.mean():.3e}   (uniform-random input: {e.max():.3e})", flush=True)
    ctx.set_option("net_path", 1)
    out1 = ctx.forward_host(x)
    print(f"[fused] vs layer-wise fp16 path: max diff {np.abs(out - out1).max():.3e}")
    ctx.set_option("net_path", 0)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
