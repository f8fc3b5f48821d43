#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference ships no golden vectors and its numeric dependencies
(librosa, keras/tensorflow) cannot be imported in this image, so these fixtures pin the
ORACLE (regression) after it has been triangulated against independent implementations
(torch.stft float64, torch.nn.LSTM, closed-form signals - see tests/test_oracle_*.py).

    python tools/make_golden.py
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import network_oracle, postprocess_oracle as po, spectrogram_oracle as so  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import pcm16_to_float, synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402

OUT = ROOT / "tests" / "golden"


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    P, S = runtime.bundled_parameters()
    sp = P["spectrogram"]

    # 1. spectrogram stage: 2 s of synthetic audio (T = 376 frames)
    pcm = synth_pcm16(2.0, seed=20251018, calls_per_minute=90.0)
    y = pcm16_to_float(pcm)
    db, freqs, times = so.calculate_spectrogram(y, sp)
    spec, lo, hi = so.preprocess_spectrogram(db, freqs, sp)
    np.savez_compressed(
        OUT / "spectrogram_2s.npz",
        pcm=pcm,
        db_band=db[:171].T.astype(np.float32),
        spec=spec.astype(np.float32),
        lo=np.float32(lo),
        hi=np.float32(hi),
        ref_power=np.float32(np.max(np.abs(so.stft_complex64(y))) ** 2),
    )

    # 2. network: seeded weights (regenerated in the tests), seeded inputs, expected probabilities
    W = synthetic_weights(P, S, seed=1234)
    rng = np.random.default_rng(5)
    x = rng.random((2, 736, 171), dtype=np.float32)
    out, inter = network_oracle.forward(x, W, return_intermediates=True)
    np.savez_compressed(
        OUT / "network_seed1234.npz",
        probs=out.astype(np.float32),
        conv0_mean=np.float32(inter["conv0"].mean()),
        block1_mean=np.float32(inter["block1"].mean()),
        block4_mean=np.float32(inter["block4"].mean()),
        final_mean=np.float32(inter["final"].mean()),
        lstm2_mean=np.float32(inter["lstm2"].mean()),
        weight_checksum=np.float64(sum(float(np.abs(v).sum()) for v in W.values())),
        n_params=np.int64(sum(v.size for v in W.values())),
    )

    # 3. post-processing: seeded predictions for a T that leaves uncovered trailing steps
    T = 736 + 368 * 9 + 200
    N = (T - 736) // 368 + 1
    rng = np.random.default_rng(11)
    base = rng.random((N, 1, 7), dtype=np.float32)
    preds = (0.55 * (0.5 * base + 0.5 * rng.random((N, 46, 7), dtype=np.float32))).astype(np.float32)
    agg, cnt = po.aggregate_predictions(preds, T, 736, 4, 7)
    s, e, n = po.binary_predictions(agg, cnt, P["calls"])
    rows = po.label_rows(s, e, n, 16, "*")
    dt = float(times[1] - times[0])
    np.savez_compressed(OUT / "postprocess_seed11.npz", preds=preds, agg=agg, cnt=cnt, starts=np.asarray(s, np.int64), stops=np.asarray(e, np.int64))
    fixtures = {
        "T": T,
        "labels": n,
        "tsv": po.labels_tsv(rows, dt),
        "delta_t": dt,
        "writer_cases": [
            {"rows": [], "tsv": po.labels_tsv([], dt)},
            {"rows": [[0, 6000, "BR*"]], "tsv": po.labels_tsv([(0, 6000, "BR*")], dt)},
            {"rows": [[0, 240, "SS*"]], "tsv": po.labels_tsv([(0, 240, "SS*")], dt)},
            {"rows": [[16, 32, "BR*"], [16, 32, "BUZZ*"], [6000, 12000, "WHISTLE*"]],
             "tsv": po.labels_tsv([(16, 32, "BR*"), (16, 32, "BUZZ*"), (6000, 12000, "WHISTLE*")], dt)},
        ],
    }
    (OUT / "postprocess_seed11.json").write_text(json.dumps(fixtures, indent=1))
    print("golden fixtures written to", OUT)
    for p in sorted(OUT.iterdir()):
        print(f"  {p.name:32s} {p.stat().st_size:>9d} B")


if __name__ == "__main__":
    main()
