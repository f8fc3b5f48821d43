"""numpy restatement of the reference's prediction post-processing (oracle; see oracle/__init__.py).

Follows, stage by stage:

* snippet indexing + overlap-average   ``src/orcAI/predict.py:244-293``
* threshold + per-label run lengths     ``src/orcAI/predict.py:298-317``
* run-length start / inclusive stop     ``src/orcAI/auxiliary.py:420-440``
* label table (sorted)                  ``src/orcAI/predict.py:320-340``
* seconds conversion + TSV text         ``src/orcAI/predict.py:343-364, 474-499``
* probabilities CSV text                ``src/orcAI/predict.py:502-531``
* duration filter                       ``src/orcAI/predict.py:14-159``

The text writers do not use pandas (the reference's own writer raises TypeError on
pandas >= 3); they restate what pandas 2.2.3 ``to_csv`` emits.
"""

from __future__ import annotations

import numpy as np


def snippet_geometry(T: int, snippet_length: int, n_filters: int):
    shift = snippet_length // 2
    ds = 2**n_filters
    pred_len = snippet_length // ds
    num_snippets = (T - snippet_length) // shift + 1
    return shift, ds, pred_len, num_snippets


def cut_snippets(spectrogram: np.ndarray, snippet_length: int) -> np.ndarray:
    shift = snippet_length // 2
    n = (spectrogram.shape[0] - snippet_length) // shift + 1
    return np.array([spectrogram[i * shift : i * shift + snippet_length] for i in range(n)])


def aggregate_predictions(predictions: np.ndarray, T: int, snippet_length: int, n_filters: int, num_labels: int):
    """float64 overlap-average of per-snippet predictions; returns (agg (T//ds, L), count (T//ds,))."""
    shift, ds, pred_len, _ = snippet_geometry(T, snippet_length, n_filters)
    total = T // ds
    agg = np.zeros((total, num_labels))
    cnt = np.zeros(total)
    for i, p in enumerate(predictions):
        s = i * (shift // ds)
        agg[s : s + pred_len] += p
        cnt[s : s + pred_len] += 1
    valid = cnt > 0
    agg[valid] /= cnt[valid, np.newaxis]
    return agg, cnt


def find_consecutive_ones(v: np.ndarray):
    d = np.diff(v, prepend=0, append=0)
    return np.where(d == 1)[0], np.where(d == -1)[0] - 1


def binary_predictions(agg: np.ndarray, cnt: np.ndarray, calls: list[str], threshold: float = 0.5):
    thr = threshold / np.max(cnt)
    b = (agg > thr).astype(int)
    starts: list[int] = []
    stops: list[int] = []
    names: list[str] = []
    for i, name in enumerate(calls):
        if sum(b[:, i]) > 0:
            s, e = find_consecutive_ones(b[:, i])
            starts += list(s)
            stops += list(e)
            names += [name] * len(s)
    return starts, stops, names


def label_rows(starts, stops, names, ds: int, suffix: str | None):
    """Rows (start_frame, stop_frame, label) sorted by (start, stop, label)."""
    if suffix is not None and suffix != "":
        names = [n + suffix for n in names]
    rows = [(int(s) * ds, int(e) * ds, n) for s, e, n in zip(starts, stops, names)]
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    return rows


def _column_text(frames: list[int], delta_t: float) -> list[str]:
    """pandas 2.2.3 semantics of ``df.loc[:, c] = df.loc[:, c] * delta_t`` then ``round(4).to_csv``.

    The int64 column keeps its dtype when every product is integer-valued (lossless
    in-place set), otherwise it is replaced by a float64 column; float64 cells are
    rounded half-to-even to 4 decimals and printed with the shortest round-trip repr.
    """
    prod = np.asarray(frames, dtype=np.int64) * np.float64(delta_t)
    if prod.size and np.all(prod == prod.astype(np.int64)):
        return [str(int(v)) for v in prod.astype(np.int64)]
    return [repr(float(v)) for v in np.round(prod, 4)]


def labels_tsv(rows, delta_t: float) -> str:
    start = _column_text([r[0] for r in rows], delta_t)
    stop = _column_text([r[1] for r in rows], delta_t)
    lines = ["start\tstop\tlabel"]
    lines += [f"{a}\t{b}\t{r[2]}" for a, b, r in zip(start, stop, rows)]
    return "\n".join(lines) + "\n"


def probabilities_csv(agg: np.ndarray, calls: list[str], delta_t: float) -> str:
    """Uncompressed text of ``<stem>_probabilities.csv.gz`` (index uses the FRAME delta_t: reference quirk)."""
    idx = np.float64(delta_t) * np.arange(len(agg))
    lines = ["time," + ",".join(calls)]
    for t, row in zip(idx, agg):
        lines.append(repr(float(t)) + "," + ",".join(repr(float(v)) for v in row))
    return "\n".join(lines) + "\n"


def filter_rows(rows, delta_t: float, limits: dict, suffix: str = "*"):
    """Keep rows whose (stop-start)*delta_t lies within the per-label [min, max] limits."""
    keep = []
    for start, stop, label in rows:
        base = label.replace(suffix, "") if suffix else label
        if base in limits:
            mn, mx = limits[base]
        elif "default" in limits:
            mn, mx = limits["default"]
        else:
            mn, mx = 0, np.inf
        mn = 0 if mn is None else mn
        mx = np.inf if mx is None else mx
        d = (stop - start) * delta_t
        if not (d < mn or d > mx):
            keep.append((start, stop, label))
    return keep
