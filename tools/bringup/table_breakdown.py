import sys, time, json, os
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from orcai_b200 import runtime, predict as op
from orcai_b200.spectrogram import load_recording
from orcai_b200.synth import synth_pcm16
from orcai_b200.wavio import write_wav_pcm16
from orcai_b200.weights import synthetic_weights
P, S = runtime.bundled_parameters()
root = Path("/tmp/orcai_table/wav"); root.mkdir(parents=True, exist_ok=True)
p = root / "rec0.wav"
if not p.exists(): write_wav_pcm16(p, synth_pcm16(3600.0, seed=20251018), 48000)
ctx = runtime.get_context(P, S, 0); ctx.load_weights(synthetic_weights(P, S, seed=1234)); ctx.set_option("net_path", 3); ctx.calibrate()
def t(f, n=3):
    f(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0) / n * 1e3, r
ms_read, samples = t(lambda: load_recording(p, 1, P["spectrogram"]))
ms_up, _ = t(lambda: ctx.upload_pcm(samples))
ms_pred, out = t(lambda: ctx.predict_pcm(samples, resident=True))
stats, agg, cnt, lab, sta, sto = out
calls = P["calls"]
ms_lab, labels = t(lambda: op.compute_labels([int(v) for v in sta], [int(v) for v in sto], [calls[int(i)] for i in lab], 16, "*"))
ms_tsv, txt = t(lambda: op.labels_to_tsv(labels, 256 / 48000))
print(json.dumps({"read_wav_ms": ms_read, "upload_pageable_ms": ms_up, "predict_resident_ms": ms_pred, "compute_labels_ms": ms_lab, "labels_to_tsv_ms": ms_tsv, "segments": len(lab)}))
