"""GPU parity: overlap-average + threshold + run-length scan kernel vs the literal numpy restatement (bit-exact)."""

import json

import numpy as np
import pytest

from oracle import postprocess_oracle as po

pytestmark = pytest.mark.gpu

CALLS = ["BR", "BUZZ", "HERDING", "PHS", "SS", "TAILSLAP", "WHISTLE"]


def compare(ctx, preds, T, threshold=0.5):
    agg_r, cnt_r = po.aggregate_predictions(preds, T, 736, 4, 7)
    s_r, e_r, n_r = po.binary_predictions(agg_r, cnt_r, CALLS, threshold)
    agg, cnt, lab, sta, sto = ctx.postprocess(preds, T, threshold)
    np.testing.assert_array_equal(agg, agg_r)
    np.testing.assert_array_equal(cnt, cnt_r)
    assert [int(v) for v in sta] == [int(v) for v in s_r]
    assert [int(v) for v in sto] == [int(v) for v in e_r]
    assert [CALLS[i] for i in lab] == n_r
    lab2, sta2, sto2 = ctx.threshold_segments(agg_r, cnt_r, threshold)
    assert list(sta2) == list(sta) and list(sto2) == list(sto) and list(lab2) == list(lab)
    return len(sta)


def test_golden(ctx, golden_dir):
    g = np.load(golden_dir / "postprocess_seed11.npz")
    fx = json.loads((golden_dir / "postprocess_seed11.json").read_text())
    agg, cnt, lab, sta, sto = ctx.postprocess(g["preds"], fx["T"])
    np.testing.assert_array_equal(agg, g["agg"])
    np.testing.assert_array_equal(cnt, g["cnt"])
    assert list(sta) == list(g["starts"]) and list(sto) == list(g["stops"]) and [CALLS[i] for i in lab] == fx["labels"]


@pytest.mark.parametrize("n_snip,extra", [(1, 0), (1, 367), (2, 0), (2, 15), (3, 16), (7, 200), (40, 367), (304, 133)])
def test_random_predictions(ctx, n_snip, extra):
    T = 736 + 368 * (n_snip - 1) + extra
    rng = np.random.default_rng(1000 * n_snip + extra)
    base = rng.random((n_snip, 1, 7), dtype=np.float32)
    preds = (0.6 * (0.5 * base + 0.5 * rng.random((n_snip, 46, 7), dtype=np.float32))).astype(np.float32)
    assert compare(ctx, preds, T) > 0


def test_extremes(ctx):
    T = 736 + 368 * 5 + 100
    n = 6
    assert compare(ctx, np.zeros((n, 46, 7), np.float32), T) == 0            # nothing above threshold
    assert compare(ctx, np.ones((n, 46, 7), np.float32), T) == 7             # one run per label, ends at the covered range
    alt = np.zeros((n, 46, 7), np.float32)
    alt[:, ::2, :] = 1.0                                                     # alternating steps -> maximal number of runs
    compare(ctx, alt, T)
    edge = np.full((n, 46, 7), 0.25, np.float32)                             # exactly on the adjusted threshold: strict ">" keeps nothing
    assert compare(ctx, edge, T) == 0
    compare(ctx, np.nextafter(edge, np.float32(1)), T)


def test_large_plane_many_tiles(ctx):
    """24-h sized plane (S = 1 012 500 steps x 7 labels = 6 921 tiles) exercises the decoupled look-back chain."""
    n_snip = 44020
    T = 16200001
    rng = np.random.default_rng(5)
    slow = np.repeat(rng.random((n_snip // 10 + 1, 1, 7), dtype=np.float32), 10, axis=0)[:n_snip]
    preds = (0.5 * slow + 0.1 * rng.random((n_snip, 46, 7), dtype=np.float32)).astype(np.float32)
    n = compare(ctx, preds, T)
    assert n > 1000


def test_find_consecutive_ones_helper(ctx):
    from orcai_b200.auxiliary import find_consecutive_ones

    rng = np.random.default_rng(2)
    for n in (1, 2, 17, 1000, 5000):
        v = (rng.random(n) < 0.4).astype(int)
        s, e = find_consecutive_ones(v)
        s_r, e_r = po.find_consecutive_ones(v)
        assert np.array_equal(s, s_r) and np.array_equal(e, e_r)
