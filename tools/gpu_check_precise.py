#!/usr/bin/env python
"""Bring-up check of the fp32-grade tensor-core path (net_path 4) against the oracle, stage by stage, and against the fp32
CUDA-core path (net_path 0) over a whole recording, with stage timings.

    timeout 600 python tools/gpu_check_precise.py [--snippets 3] [--minutes 10] [--bn-matched]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import network_oracle  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402

STAGE_NAMES = ["conv0", "block1", "block2", "block3", "block4", "final_sep", "lstm1_proj", "lstm1_rec", "lstm2_proj", "lstm2_rec", "dense"]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--snippets", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--bn-matched", action="store_true", help="also check weights whose BatchNorm statistics match the activations")
    ap.add_argument("--path", type=int, default=4)
    ap.add_argument("--json", default="")
    ap.add_argument("--only-timing", action="store_true", help="skip the stage checks (profiling runs)")
    args = ap.parse_args()
    P, S = runtime.bundled_parameters()
    ctx = runtime.get_context(P, S, 0)
    W = synthetic_weights(P, S, seed=args.seed)
    ctx.load_weights(W)
    x = np.random.default_rng(5).random((args.snippets, 736, 171), dtype=np.float32)
    ref, inter = network_oracle.forward(x, W, return_intermediates=True)
    nhwc = {k: np.transpose(v, (0, 2, 3, 1)) for k, v in inter.items() if v.ndim == 4}
    ok = True
    report = {}
    ctx.set_option("net_path", args.path)
    stages = [(0, "conv0 (hi, ch 0-7)", nhwc["conv0"][..., :8]), (1, "block1", nhwc["block1"]),
              (2, "block2", nhwc["block2"]), (3, "block3", nhwc["block3"]), (4, "block4", nhwc["block4"]), (5, "final", nhwc["final"])]
    for stage, key, want in ([] if args.only_timing else stages):
        t0 = time.time()
        got = ctx.debug_stage(x, stage)
        if got.shape != want.shape:
            print(f"[precise] stage {key}: shape {got.shape} != {want.shape}")
            ok = False
            continue
        err = np.abs(got - want)
        scale = np.abs(want).max()
        fin = np.isfinite(got).all()
        tol = (2e-3 if stage == 0 else 2e-5) * scale + 1e-5
        print(f"[precise] stage {key:16s} shape {got.shape} max|ref| {scale:9.4f} max err {np.nanmax(err):.3e} mean err {np.nanmean(err):.3e} finite {fin} ({time.time() - t0:.2f}s)", flush=True)
        if not fin or err.max() > tol:
            ok = False
            bad = np.unravel_index(np.nanargmax(np.where(np.isfinite(err), err, np.inf)), err.shape)
            print(f"      worst at {bad}: got {got[bad]} want {want[bad]}")
            e_hw = np.where(np.isfinite(err), err, 1e9).max(axis=(0, 3))
            hh, ww = np.where(e_hw > tol)
            print(f"      bad rows {np.unique(hh)[:24]} ... cols {np.unique(ww)[:24]} ({len(hh)} bad pixels of {e_hw.size})")
    out = ctx.forward_host(x)
    e = np.abs(out - ref)
    print(f"[precise] probabilities vs oracle: max err {e.max():.3e} mean {e.mean():.3e}", flush=True)
    report["random_input_vs_oracle"] = {"max": float(e.max()), "mean": float(e.mean())}
    ok &= bool(e.max() < 2e-4)
    # ragged chunking / batch independence
    x7 = np.random.default_rng(9).random((7, 736, 171), dtype=np.float32)
    full = ctx.forward_host(x7)
    ctx.set_option("chunk", 3)
    chunked = ctx.forward_host(x7)
    ctx.set_option("chunk", 1024)
    same = np.array_equal(full, chunked) and np.array_equal(ctx.forward_host(x7[4:5]), full[4:5])
    print(f"[precise] chunking / batch independence bit-identical: {same}")
    ok &= same
    if args.path == 4:   # depthwise fused into the split GEMM (default) against the two-kernel form
        ctx.set_option("precise_sep_path", 0)
        two = ctx.forward_host(x7)
        ctx.set_option("precise_sep_path", 1)
        same = np.array_equal(two, full)
        print(f"[precise] fused depthwise+GEMM kernel bit-identical to depthwise kernel -> GEMM: {same} (max diff {np.abs(two - full).max():.3e})")
        ok &= same

    # whole recording: per-snippet probabilities against the fp32 CUDA-core path, stage timings
    def whole(tag, weights):
        nonlocal ok
        ctx.load_weights(weights)
        pcm = synth_pcm16(args.minutes * 60.0, seed=20251018)
        ctx.upload_pcm(pcm)
        st = ctx.spectrogram_resident(normalise=False)
        n = int((st.n_frames - 736) // 368 + 1)
        ctx.set_option("net_path", 0)
        ref32 = ctx.forward_resident(0, n)
        res = {}
        for path in (args.path, 3):
            ctx.set_option("net_path", path)
            if path == 3:
                ctx.calibrate()          # runs on the built-in calibration recording: bring ours back
                ctx.upload_pcm(pcm)
                ctx.spectrogram_resident(normalise=False)
            if path == 4:   # shared interior (tall image + border rows) against the snippet-by-snippet evaluation: bit-identical
                ctx.set_option("precise_tall", 0)
                per_snippet = ctx.forward_resident(0, n)
                ns0 = ctx.timings()["net_stage_ms"]
                print("          snippet-by-snippet stage ms: " + " ".join(f"{k}={v:.3f}" for k, v in zip(STAGE_NAMES, ns0)))
                ctx.set_option("precise_tall", 1)
                ctx.set_option("chunk", 100)     # several chunks, ragged last one
                chunked = ctx.forward_resident(0, n)
                ctx.set_option("chunk", 1024)
            ctx.forward_resident(0, min(n, 64))
            t0 = time.time()
            got = ctx.forward_resident(0, n)
            dt = time.time() - t0
            tm = ctx.timings()
            if path == 4:
                same = np.array_equal(got, per_snippet) and np.array_equal(got, chunked)
                nd = int((got != per_snippet).sum())
                print(f"[precise] {tag}: tall-image evaluation bit-identical to snippet-by-snippet: {same} ({nd} of {got.size} values differ, max {np.abs(got - per_snippet).max():.3e}; chunked equal: {np.array_equal(got, chunked)})")
                if nd:
                    bad = np.argwhere(got != per_snippet)
                    print("          first differing (snippet, row, label):", bad[:8].tolist(), " rows:", np.unique(bad[:, 1])[:46].tolist())
                ok &= same
            d = np.abs(got - ref32)
            ns = tm["net_stage_ms"]
            res[path] = {"max": float(d.max()), "mean": float(d.mean()), "p999": float(np.quantile(d, 0.999)), "snippets": n, "wall_s": dt,
                         "stage_ms_first_chunk": {k: round(float(v), 4) for k, v in zip(STAGE_NAMES, ns)}, "marked_snippets": int(ns[15])}
            print(f"[precise] {tag}: {n} snippets, net_path {path} vs fp32 path: max {d.max():.3e} mean {d.mean():.3e} p99.9 {np.quantile(d, 0.999):.3e}  wall {dt * 1e3:.1f} ms")
            print(f"          stage ms (first chunk of {int(ns[15])}): " + " ".join(f"{k}={v:.3f}" for k, v in zip(STAGE_NAMES, ns)), flush=True)
        ok &= res[args.path]["max"] <= 5e-4
        report[tag] = res

    whole("seeded weights", W)
    if args.bn_matched:
        sys.path.insert(0, str(Path(__file__).resolve().parent))
        import precision_study as ps
        from oracle import postprocess_oracle as po, spectrogram_oracle as so
        from orcai_b200.synth import pcm16_to_float

        spec, _, _ = so.make_spectrogram(pcm16_to_float(synth_pcm16(20.0, seed=20251018)), P["spectrogram"])
        whole("BatchNorm-matched weights", ps.calibrate_bn(W, po.cut_snippets(spec, 736)[:2]))
    ctx.load_weights(W)
    ctx.set_option("net_path", 0)
    if args.json:
        Path(args.json).parent.mkdir(parents=True, exist_ok=True)
        Path(args.json).write_text(json.dumps(report, indent=1))
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
