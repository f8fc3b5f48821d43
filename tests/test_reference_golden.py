"""Oracle, host layer and CUDA post-processing against golden vectors produced by the REFERENCE'S OWN CODE.

tests/golden/reference_* were written by tools/make_reference_golden.py, which imports ethz-tb/orcAI v1.0.3 from
/root/reference (third-party imports stubbed) and runs its unmodified numpy / pandas functions:
preprocess_spectrogram (spectrogram.py:58-87), compute_aggregated_predictions (predict.py:235-295, with a fake model),
compute_binary_predictions (:298-317), find_consecutive_ones (auxiliary.py:420-440), compute_labels (predict.py:320-340),
filter_predictions (:69-159), save_prediction_probabilities (:502-531).  Everything here is bit-exact.
"""

import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

from make_reference_golden import fake_predictions, inputs  # noqa: E402
from oracle import postprocess_oracle as po, spectrogram_oracle as so  # noqa: E402
from orcai_b200 import auxiliary, predict, runtime  # noqa: E402


@pytest.fixture(scope="module")
def ref(golden_dir):
    pre = np.load(golden_dir / "reference_preprocess.npz")
    post = np.load(golden_dir / "reference_postprocess.npz")
    meta = json.loads((golden_dir / "reference_postprocess.json").read_text())
    return {"spec": pre["spec"], **{k: post[k] for k in post.files}, **meta, "inp": inputs()}


def test_oracle_preprocess_equals_reference(ref, params):
    """crop [0:171], nearest-rank 1 % / 99.9 % percentiles, clip, min-max normalise, transpose."""
    P, _ = params
    spec, lo, hi = so.preprocess_spectrogram(ref["inp"]["db"], ref["inp"]["freqs"], P["spectrogram"])
    assert spec.dtype == np.float32 and spec.shape == ref["spec"].shape
    np.testing.assert_array_equal(spec, ref["spec"])
    assert spec.min() == 0.0 and spec.max() == 1.0 and lo < hi


def test_oracle_batcher_and_aggregation_equal_reference(ref):
    inp = ref["inp"]
    sn = po.cut_snippets(inp["spec"], 736)
    assert list(ref["snippet_shape"]) == [sn.shape[0], 736, 171, 1]       # the model sees (N, 736, 171, 1)
    np.testing.assert_array_equal(sn[:, 0, :], ref["snippet_first_rows"])   # window i starts at row 368 * i
    np.testing.assert_array_equal(sn[:, -1, :], ref["snippet_last_rows"])
    preds = fake_predictions(sn, inp["w"])
    agg, cnt = po.aggregate_predictions(preds, inp["T2"], 736, 4, 7)
    assert agg.dtype == np.float64
    np.testing.assert_array_equal(agg, ref["agg"])
    np.testing.assert_array_equal(cnt, ref["cnt"])
    assert (cnt[-6:] == 0).all() and (agg[-6:] == 0).all()                  # trailing frames no window covers stay 0


def test_oracle_and_host_segments_equal_reference(ref, params):
    P, _ = params
    s, e, n = po.binary_predictions(ref["agg"], ref["cnt"], P["calls"])
    assert [int(v) for v in s] == ref["row_starts"].tolist() and [int(v) for v in e] == ref["row_stops"].tolist()
    assert list(n) == ref["label_names"]
    rows = po.label_rows(s, e, n, 16, "*")
    assert [[int(a), int(b), c] for a, b, c in rows] == ref["labels"]
    # the product's host functions (same signatures as the reference's)
    df = predict.compute_labels([int(v) for v in s], [int(v) for v in e], list(n), 16, "*")
    assert [[int(a), int(b), str(c)] for a, b, c in zip(df["start"], df["stop"], df["label"])] == ref["labels"]
    assert [str(df["start"].dtype), str(df["stop"].dtype)] == ref["labels_dtypes"]
    for case in ref["find_consecutive_ones"]:
        a, b = po.find_consecutive_ones(np.array(case["x"]))
        assert [int(v) for v in a] == case["starts"] and [int(v) for v in b] == case["stops"]


def test_duration_filter_and_probabilities_csv_equal_reference(ref, params):
    P, _ = params
    df = pd.DataFrame(ref["labels"], columns=["start", "stop", "label"])
    dt = 256 / 48000
    out = predict.filter_predictions(df, delta_t=dt, call_duration_limits=ref["filter_limits"], label_suffix="*", verbosity=0)
    assert [[int(a), int(b), str(c)] for a, b, c in zip(out["start"], out["stop"], out["label"])] == ref["filtered"]
    assert [[int(a), int(b), c] for a, b, c in po.filter_rows([tuple(r) for r in ref["labels"]], dt, ref["filter_limits"])] == ref["filtered"]
    for text in (po.probabilities_csv(ref["agg"], P["calls"], dt), predict.probabilities_to_csv(ref["agg"], P["calls"], dt)):
        assert len(text) == ref["probabilities_csv_len"] and text[:4000] == ref["probabilities_csv_head"]
        assert hashlib.sha256(text.encode()).hexdigest() == ref["probabilities_csv_sha256"]


@pytest.mark.gpu
def test_cuda_postprocess_equals_reference(ctx, ref, params):
    """K7 (overlap-average + threshold + run-length scan) through the C ABI against the reference's outputs."""
    P, _ = params
    inp = ref["inp"]
    preds = fake_predictions(po.cut_snippets(inp["spec"], 736), inp["w"])
    agg, cnt, lab, sta, sto = ctx.postprocess(preds, inp["T2"], threshold=0.5, want_agg=True)
    np.testing.assert_array_equal(agg, ref["agg"])
    np.testing.assert_array_equal(cnt, ref["cnt"])
    assert sta.tolist() == ref["row_starts"].tolist() and sto.tolist() == ref["row_stops"].tolist()
    assert [P["calls"][int(i)] for i in lab] == ref["label_names"]
    lab2, sta2, sto2 = ctx.threshold_segments(ref["agg"], ref["cnt"], 0.5)
    assert sta2.tolist() == sta.tolist() and sto2.tolist() == sto.tolist() and lab2.tolist() == lab.tolist()
    for case in ref["find_consecutive_ones"]:                               # the product's function runs the scan kernel
        a, b = auxiliary.find_consecutive_ones(np.array(case["x"]))
        assert [int(v) for v in a] == case["starts"] and [int(v) for v in b] == case["stops"]


@pytest.mark.gpu
def test_cuda_normalisation_matches_reference_rule(ctx, params):
    """K1 -> select -> K2 on the device array == the reference's preprocess rule applied (by the pinned oracle) to that array."""
    from orcai_b200.synth import synth_pcm16

    P, _ = params
    pcm = synth_pcm16(6.0, seed=99, calls_per_minute=60.0)
    spec, st = ctx.spectrogram(pcm)
    db_dev = ctx.read_db(0, spec.shape[0])                                  # shifted + floored dB of the kept band, (T, 171)
    full = np.full((257, db_dev.shape[0]), -80.0, np.float32)
    full[:171] = db_dev.T
    want, lo, hi = so.preprocess_spectrogram(full, np.fft.rfftfreq(512, d=1.0 / 48000), P["spectrogram"])
    assert np.float32(st.lo) == np.float32(lo) and np.float32(st.hi) == np.float32(hi)
    np.testing.assert_array_equal(spec, want)
