"""Pin the post-processing oracle: properties (hypothesis), golden vectors, writer fixtures."""

import json

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import postprocess_oracle as po

CALLS = ["BR", "BUZZ", "HERDING", "PHS", "SS", "TAILSLAP", "WHISTLE"]


def test_geometry():
    assert po.snippet_geometry(112501, 736, 4) == (368, 16, 46, 304)  # SURVEY section 6, 10-min recording
    assert po.snippet_geometry(736, 736, 4)[3] == 1 and po.snippet_geometry(735, 736, 4)[3] == 0


@given(st.lists(st.integers(0, 1), min_size=0, max_size=200))
def test_run_lengths(bits):
    v = np.asarray(bits, dtype=int)
    s, e = po.find_consecutive_ones(v)
    assert len(s) == len(e)
    rebuilt = np.zeros_like(v)
    for a, b in zip(s, e):
        assert a <= b and v[a : b + 1].all()
        assert (a == 0 or v[a - 1] == 0) and (b == len(v) - 1 or v[b + 1] == 0)
        rebuilt[a : b + 1] = 1
    assert np.array_equal(rebuilt, v)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 12), st.integers(0, 367), st.integers(0, 2**31 - 1))
def test_aggregate_properties(n_snip, extra, seed):
    T = 736 + 368 * (n_snip - 1) + extra
    rng = np.random.default_rng(seed)
    preds = rng.random((n_snip, 46, 7), dtype=np.float32)
    agg, cnt = po.aggregate_predictions(preds, T, 736, 4, 7)
    assert agg.shape == (T // 16, 7) and agg.dtype == np.float64
    covered = 23 * (n_snip - 1) + 46
    assert np.all(cnt[:covered] >= 1) and np.all(cnt[covered:] == 0) and np.all(agg[covered:] == 0)
    assert cnt.max() == (2 if n_snip > 1 else 1)
    # first half-snippet is covered once and equals the raw prediction
    np.testing.assert_array_equal(agg[:23], preds[0, :23].astype(np.float64))
    if n_snip > 1:
        np.testing.assert_array_equal(agg[23:46], (preds[0, 23:].astype(np.float64) + preds[1, :23].astype(np.float64)) / 2.0)
    s, e, n = po.binary_predictions(agg, cnt, CALLS)
    thr = 0.5 / cnt.max()
    mask = agg > thr
    total = sum(int(((np.diff(np.concatenate([[0], mask[:, l].astype(int), [0]]))) == 1).sum()) for l in range(7))
    assert len(s) == len(e) == len(n) == total


def test_golden_postprocess(golden_dir):
    g = np.load(golden_dir / "postprocess_seed11.npz")
    fx = json.loads((golden_dir / "postprocess_seed11.json").read_text())
    agg, cnt = po.aggregate_predictions(g["preds"], fx["T"], 736, 4, 7)
    np.testing.assert_array_equal(agg, g["agg"])
    np.testing.assert_array_equal(cnt, g["cnt"])
    s, e, n = po.binary_predictions(agg, cnt, CALLS)
    assert [int(v) for v in s] == list(g["starts"]) and [int(v) for v in e] == list(g["stops"]) and n == fx["labels"]
    assert po.labels_tsv(po.label_rows(s, e, n, 16, "*"), fx["delta_t"]) == fx["tsv"]


def test_writer_fixtures(golden_dir):
    fx = json.loads((golden_dir / "postprocess_seed11.json").read_text())
    dt = 256 / 48000
    assert fx["delta_t"] == dt
    cases = {json.dumps(c["rows"]): c["tsv"] for c in fx["writer_cases"]}
    assert cases["[]"] == "start\tstop\tlabel\n"
    # every product integer-valued -> the int64 column survives pandas 2.2.3's in-place set
    assert cases['[[0, 6000, "BR*"]]'] == "start\tstop\tlabel\n0\t32\tBR*\n"
    assert cases['[[0, 240, "SS*"]]'] == "start\tstop\tlabel\n0\t1.28\tSS*\n"
    assert po.labels_tsv([(16, 32, "BR*"), (6000, 12000, "W*")], dt) == "start\tstop\tlabel\n0.0853\t0.1707\tBR*\n32.0\t64.0\tW*\n"


def test_probabilities_csv_and_filter():
    agg = np.array([[0.5, 0.25], [1 / 3, 0.0]])
    txt = po.probabilities_csv(agg, ["A", "B"], 256 / 48000)
    assert txt == "time,A,B\n0.0,0.5,0.25\n0.005333333333333333,0.3333333333333333,0.0\n"
    rows = [(0, 160, "BR*"), (0, 16, "SS*"), (32, 6000, "X*")]
    lim = {"default": [0.5, None], "SS": [0, 0.05], "BR": [None, 1.0]}
    assert po.filter_rows(rows, 256 / 48000, lim) == [(0, 160, "BR*"), (32, 6000, "X*")]
