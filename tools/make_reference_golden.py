#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN CODE (ethz-tb/orcAI v1.0.3 mounted at /root/reference).

The reference's numerics for the STFT and the network live in librosa / Keras, which cannot be installed here.  Everything
else on the hot path is plain numpy / pandas inside the reference's modules, and those modules import fine once the missing
third-party packages are replaced by inert stubs.  This script runs, unmodified from /root/reference/src:

    orcAI.spectrogram.preprocess_spectrogram          (spectrogram.py:58-87)   crop, nearest-rank percentiles, clip, normalise
    orcAI.predict.compute_aggregated_predictions      (predict.py:235-295)     snippet batcher + overlap-average (fake model)
    orcAI.predict.compute_binary_predictions          (predict.py:298-317)     threshold / max(count), per-label runs
    orcAI.auxiliary.find_consecutive_ones             (auxiliary.py:420-440)
    orcAI.predict.compute_labels                      (predict.py:320-340)     frame units, suffix, sort
    orcAI.predict.filter_predictions                  (predict.py:69-159)      duration limits
    orcAI.predict.save_prediction_probabilities       (predict.py:502-531)     gzip CSV text

on seeded inputs and freezes their outputs under tests/golden/reference_*.  tests/test_reference_golden.py holds the oracle
(and, on a GPU, the CUDA path) to them bit for bit.  The inputs are regenerated from seeds by `inputs()` below, which the tests
import, so only outputs are stored.

    python tools/make_reference_golden.py        (needs /root/reference; not needed at test time)
"""
from __future__ import annotations

import gzip
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden"
REFERENCE_SRC = Path("/root/reference/src")


def inputs() -> dict:
    """Seeded inputs shared by this generator and the tests (numpy Generator streams are stable across versions)."""
    rng = np.random.default_rng(20251018)
    # a dB-like (257, T) float32 array with the -80 floor hit in places, like amplitude_to_db(..., ref=np.max) output
    T1 = 800
    base = -45.0 + 18.0 * rng.standard_normal((257, T1))
    base += 25.0 * np.exp(-0.5 * ((np.arange(257)[:, None] - 60.0) / 9.0) ** 2) * (rng.random((1, T1)) > 0.7)
    db = np.clip(base, -80.0, 0.0).astype(np.float32)
    db[17, 123] = 0.0  # the loudest cell
    freqs = np.fft.rfftfreq(512, d=1.0 / 48000)
    # a normalised-spectrogram-like (T, 171) array for the batcher; T leaves 100 uncovered trailing frames
    T2 = 736 + 368 * 3 + 100
    spec = rng.random((T2, 171), dtype=np.float32)
    # per-snippet predictions the fake model returns: a function of the snippet so that wrong windows cannot pass
    w = rng.random((171, 7), dtype=np.float32)
    return {"db": db, "freqs": freqs, "spec": spec, "w": w, "T2": T2}


def fake_predictions(snippets: np.ndarray, w: np.ndarray) -> np.ndarray:
    """(N, 736, 171[,1]) -> (N, 46, 7) float32 in (0, 0.6): mean over each 16-frame step, projected and squashed."""
    x = np.asarray(snippets, dtype=np.float32)
    if x.ndim == 4:
        x = x[..., 0]
    steps = x.reshape(x.shape[0], 46, 16, 171).mean(axis=2, dtype=np.float32)
    z = (steps @ w) / np.float32(171.0)
    return (0.6 * z / z.max()).astype(np.float32)


def import_reference():
    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    class _Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()

    for name in ("keras", "tensorflow", "zarr", "librosa", "humanize"):
        sys.modules.setdefault(name, _Stub(name))
    sys.path.insert(0, str(REFERENCE_SRC))
    import orcAI.auxiliary as ra
    import orcAI.predict as rp
    import orcAI.spectrogram as rs

    return rs, rp, ra


def main() -> int:
    if not REFERENCE_SRC.exists():
        print("reference not mounted at /root/reference - nothing to do", file=sys.stderr)
        return 1
    rs, rp, ra = import_reference()
    P = json.loads((REFERENCE_SRC / "orcAI" / "models" / "orcai-V1" / "orcai_parameter.json").read_text())
    S = json.loads((REFERENCE_SRC / "orcAI" / "models" / "orcai-V1" / "model_shape.json").read_text())
    inp = inputs()
    OUT.mkdir(parents=True, exist_ok=True)

    # ---- preprocess_spectrogram --------------------------------------------------------------
    spec_ref = rs.preprocess_spectrogram(inp["db"].copy(), inp["freqs"], P["spectrogram"])
    assert spec_ref.shape == (800, 171) and spec_ref.dtype == np.float32
    np.savez_compressed(OUT / "reference_preprocess.npz", spec=np.ascontiguousarray(spec_ref))

    # ---- batcher + aggregation with a fake model ---------------------------------------------
    seen = {}

    class FakeModel:
        def predict(self, snippets, verbose=0, **kw):
            seen["snippets"] = np.array(snippets)
            return fake_predictions(snippets, inp["w"])

    agg, cnt = rp.compute_aggregated_predictions(Path("x.wav"), inp["spec"], FakeModel(), P, S, ra.Messenger(verbosity=0), None)
    sn = seen["snippets"]
    row_starts, row_stops, label_names = rp.compute_binary_predictions(agg, cnt, P["calls"], threshold=0.5)
    labels = rp.compute_labels(row_starts, row_stops, label_names, time_steps_per_output_step=16, label_suffix="*")
    dt = np.float64(256) / np.float64(48000)   # times[1] - times[0] of librosa.frames_to_time: a numpy float64 in the reference
    limits = {"default": [0.05, None], "SS": [0.0, 0.5], "BR": [0.2, 3.0]}
    with tempfile.TemporaryDirectory() as td:
        rp.save_prediction_probabilities(agg, P, dt, Path(td) / "rec_predicted.txt", ra.Messenger(verbosity=0))
        prob_csv = gzip.decompress((Path(td) / "rec_predicted_probabilities.csv.gz").read_bytes()).decode()
    filtered = rp.filter_predictions(labels.copy(), delta_t=float(dt), call_duration_limits=limits, label_suffix="*", msgr=ra.Messenger(verbosity=0))
    ones_cases = [[0, 1, 1, 0, 1], [1, 1, 1], [0, 0, 0], [1], [0], [1, 0, 1, 0, 1, 1]]
    np.savez_compressed(
        OUT / "reference_postprocess.npz",
        snippet_shape=np.asarray(sn.shape, np.int64),
        snippet_first_rows=np.ascontiguousarray(sn[..., 0][:, 0, :] if sn.ndim == 4 else sn[:, 0, :]),   # row 0 of every snippet identifies its window
        snippet_last_rows=np.ascontiguousarray(sn[..., 0][:, -1, :] if sn.ndim == 4 else sn[:, -1, :]),
        agg=agg, cnt=cnt,
        row_starts=np.asarray(row_starts, np.int64), row_stops=np.asarray(row_stops, np.int64),
    )
    (OUT / "reference_postprocess.json").write_text(json.dumps({
        "source": "ethz-tb/orcAI v1.0.3, src/orcAI/{spectrogram,predict,auxiliary}.py executed by tools/make_reference_golden.py",
        "label_names": list(label_names),
        "labels": [[int(a), int(b), str(c)] for a, b, c in zip(labels["start"], labels["stop"], labels["label"])],
        "labels_dtypes": [str(labels["start"].dtype), str(labels["stop"].dtype)],
        "filter_limits": limits,
        "filtered": [[int(a), int(b), str(c)] for a, b, c in zip(filtered["start"], filtered["stop"], filtered["label"])],
        "probabilities_csv_head": prob_csv[:4000],
        "probabilities_csv_len": len(prob_csv),
        "probabilities_csv_sha256": __import__("hashlib").sha256(prob_csv.encode()).hexdigest(),
        "find_consecutive_ones": [{"x": x, "starts": [int(v) for v in ra.find_consecutive_ones(np.array(x))[0]],
                                   "stops": [int(v) for v in ra.find_consecutive_ones(np.array(x))[1]]} for x in ones_cases],
    }, indent=1))
    for p in sorted(OUT.glob("reference_*")):
        print(f"  {p.name:34s} {p.stat().st_size:>9d} B")
    print(f"snippets seen by the model: {sn.shape}; {len(labels)} label rows, {len(filtered)} after the duration filter")
    return 0


if __name__ == "__main__":
    sys.exit(main())
