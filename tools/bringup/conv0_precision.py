#!/usr/bin/env python
"""Fast-path deviation from the fp32 path with (a) the plain seeded weights and (b) weights whose BatchNorm statistics match
the activations (what training leaves behind; tools/precision_study.py: calibrate_bn), for both entry-convolution variants.

    timeout 600 python tools/bringup/conv0_precision.py [--seconds 600]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
from oracle import postprocess_oracle as po, spectrogram_oracle as so  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import pcm16_to_float, synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402
import precision_study as ps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=600.0)
a = ap.parse_args()
P, S = runtime.bundled_parameters()
ctx = runtime.get_context(P, S, 0)
W0 = synthetic_weights(P, S, seed=1234)
cal_pcm = synth_pcm16(30.0, seed=77, calls_per_minute=30.0)
spec, _, _ = so.make_spectrogram(pcm16_to_float(cal_pcm), P["spectrogram"])
Wc = ps.calibrate_bn(W0, po.cut_snippets(spec, 736)[:4])
pcm = synth_pcm16(a.seconds, seed=20251018)
for wname, W in (("seed1234", W0), ("seed1234 + BatchNorm statistics matched", Wc)):
    ctx.load_weights(W)
    ctx.set_option("net_path", 0)
    ref = ctx.predict_pcm(pcm)
    ctx.calibrate()
    ctx.set_option("net_path", 3)
    for c0 in (1, 0):
        ctx.set_option("conv0_path", c0)
        ctx.predict_pcm(pcm)
        out = ctx.predict_pcm(pcm)
        t = ctx.timings()
        d = np.abs(out[1] - ref[1])
        seg_f = set(zip(out[3].tolist(), out[4].tolist(), out[5].tolist()))
        seg_r = set(zip(ref[3].tolist(), ref[4].tolist(), ref[5].tolist()))
        print(f"{wname:42s} conv0_path {c0}: aggregated probability deviation max {d.max():.2e} mean {d.mean():.2e} | "
              f"segments {len(seg_f)} (fp32 path {len(seg_r)}, identical {len(seg_f & seg_r)}) | network {t['network_ms']:.2f} ms "
              f"(entry conv {t['net_stage_ms'][0]:.3f})", flush=True)
    ctx.set_option("conv0_path", 1)
    ctx.set_option("net_path", 0)
