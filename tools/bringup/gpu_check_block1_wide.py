#!/usr/bin/env python
"""Bring-up check of the N-widened block 1 (block1_path 1, net_fused_w.cuh) against block1_path 0 and the oracle.

    timeout 300 python tools/bringup/gpu_check_block1_wide.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import network_oracle  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402

P, S = runtime.bundled_parameters()
ctx = runtime.get_context(P, S, 0)
W = synthetic_weights(P, S, seed=1234)
ctx.load_weights(W)
x = np.random.default_rng(5).random((5, 736, 171), dtype=np.float32)
ref, inter = network_oracle.forward(x, W, return_intermediates=True)
want = np.maximum(np.transpose(inter["block1"], (0, 2, 3, 1)), 0)
want_sub = np.transpose(inter["block1"], (0, 2, 3, 1))[:, ::2, ::2]
ctx.set_option("net_path", 3)
ok = True
got = {}
for path in (0, 1):
    ctx.set_option("block1_path", path)
    for stage, w, name in ((1, want, "block1 relu"), (21, want_sub, "block1 sub")):
        g = ctx.debug_stage(x, stage)
        got[(path, stage)] = g
        err = np.abs(g - w)
        scale = np.abs(w).max()
        print(f"[block1_path {path}] {name:12s} max|ref| {scale:.3f} max err {np.nanmax(err):.3e} mean {np.nanmean(err):.3e} finite {np.isfinite(g).all()}", flush=True)
        if not np.isfinite(g).all() or err.max() > 0.02 * scale + 1e-2:
            ok = False
            e_hw = np.where(np.isfinite(err), err, 1e9).max(axis=(0, 3))
            hh, ww = np.where(e_hw > 0.02 * scale + 1e-2)
            print(f"      bad rows {np.unique(hh)[:24]} ... cols {np.unique(ww)[:40]} ({len(hh)} bad pixels of {e_hw.size})")
    out = ctx.forward_host(x)
    e = np.abs(out - ref)
    print(f"[block1_path {path}] probabilities max err {e.max():.3e} mean {e.mean():.3e}", flush=True)
    ok &= bool(e.max() < 5e-3)
d = np.abs(got[(0, 1)] - got[(1, 1)])
print(f"block1_path 1 vs 0: max diff {d.max():.3e} (fp32 accumulation order differs; fp16 storage rounding may flip)")
pcm = synth_pcm16(600.0, seed=20251018)
ctx.calibrate()
for path in (0, 1, 0, 1):
    ctx.set_option("block1_path", path)
    ctx.predict_pcm(pcm)
    ctx.predict_pcm(pcm)
    t = ctx.timings()
    print(f"[block1_path {path}] stage ms: {[round(v, 3) for v in t['net_stage_ms'][:6]]} network {t['network_ms']:.3f}", flush=True)
ctx.set_option("block1_path", 0)
ctx.set_option("net_path", 0)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
