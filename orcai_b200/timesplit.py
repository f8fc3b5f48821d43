"""One recording split by TIME across several contexts / GPUs ("splitting ONE recording by time").

The reference annotates a recording as one unit: one reference level (``ref=np.max`` over all bins and frames,
``spectrogram.py:51-53``), one pair of nearest-rank percentiles (``spectrogram.py:70-75``), snippets every 368 frames
(``predict.py:252-261``) and one overlap-average / run-length pass (``predict.py:276-317``).  Splitting it by time therefore
needs exactly one exchange of partial statistics; everything else is independent per chunk:

1. plan: the N snippets are cut into contiguous ranges, one per context.  A chunk's samples cover its snippets' frames plus
   halos (one snippet shift = 368 frames before, one frame after) so that every frame a chunk uses or counts is computed
   from real samples; the frames are partitioned between the chunks for the statistics ("owned" rows).
2. every context: STFT -> dB of its chunk, max |S|^2 over its owned rows            -> host: max of the maxima
3. three radix-select passes: every context histograms its owned rows               -> host: sum, locate both ranks
4. every context: lo / hi set, snippet forward over its own snippets                -> host: concatenate the predictions
5. one context: overlap-average + threshold + run-length scan over all snippets.

No collective is needed beyond this host-side gather (``north_star``); results are bit-identical to the one-context path
(``tests/test_gpu_timesplit.py`` runs the chunks on several contexts of one GPU, ``tests/test_dist_cpu.py`` checks the plan
and the rank search on the CPU).
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Chunk:
    snippet0: int       # first global snippet of the chunk
    n_snippets: int
    sample0: int        # [sample0, sample1) of the recording are uploaded
    sample1: int
    frame0: int         # global frame of the chunk's local row 0
    own_row0: int       # local rows [own_row0, own_row1) count towards the recording's statistics
    own_row1: int
    local_snippet0: int  # index of snippet0 among the chunk's local snippets


def plan_chunks(n_samples: int, n_chunks: int, hop: int = 256, snippet_len: int = 736) -> list[Chunk]:
    """Cut a recording of `n_samples` into at most `n_chunks` time chunks (fewer if it has fewer snippets)."""
    shift = snippet_len // 2
    T = 1 + n_samples // hop
    N = (T - snippet_len) // shift + 1 if T >= snippet_len else 0
    if N <= 0:
        raise ValueError(f"recording has {T} frames, shorter than one snippet of {snippet_len}")
    G = max(1, min(n_chunks, N))
    bounds = [N * g // G for g in range(G + 1)]
    chunks = []
    for g in range(G):
        i0, i1 = bounds[g], bounds[g + 1]
        own0 = 0 if g == 0 else shift * i0
        own1 = T if g == G - 1 else shift * i1
        frame0 = 0 if g == 0 else shift * i0 - shift          # one snippet shift of halo: local snippets stay aligned
        end = max(shift * (i1 - 1) + snippet_len, own1) + 1     # one more frame so that the last used frame is exact
        if end >= T:
            sample1, end = n_samples, T
        else:
            sample1 = hop * end
        chunks.append(Chunk(i0, i1 - i0, hop * frame0, sample1, frame0, own0 - frame0, own1 - frame0, (shift * i0 - frame0) // shift))
    return chunks


def key_to_float(key: int) -> np.float32:
    """Inverse of the order-preserving uint32 image of a float32 used by the radix select (select.cu: key2f)."""
    u = (key & 0x7FFFFFFF) if (key & 0x80000000) else (~key & 0xFFFFFFFF)
    return np.array([u], dtype=np.uint32).view(np.float32)[0]


def locate_ranks(hists: list[np.ndarray], ranks: list[int], prefixes: list[int], pass_: int) -> tuple[list[int], list[int]]:
    """One pass of the two-rank radix select on the host: sum the chunks' histograms, find the digit holding each rank.

    Mirrors select_scan_kernel (select.cu): pass 0 shares one histogram between both ranks; digits are 11, 11 and 10 bits.
    """
    total = np.sum(np.stack(hists).astype(np.uint64), axis=0)
    shift = (21, 10, 0)[pass_]
    nbins = 1024 if pass_ == 2 else 2048
    new_ranks, new_prefixes = [], []
    for r in range(2):
        h = total[0 if pass_ == 0 else r][:nbins]
        cum = np.cumsum(h, dtype=np.uint64)
        d = int(np.searchsorted(cum, np.uint64(ranks[r]), side="right"))
        if d >= nbins:
            raise RuntimeError("radix select: rank outside the histogram (inconsistent chunk statistics)")
        below = int(cum[d - 1]) if d else 0
        new_ranks.append(ranks[r] - below)
        new_prefixes.append(prefixes[r] | (d << shift))
    return new_ranks, new_prefixes


def nearest_ranks(n: int, q_lo: float, q_hi: float) -> list[int]:
    """np.percentile(method='nearest'): index around((n - 1) q), half to even (spectrogram.py:70-75)."""
    return [int(np.rint((n - 1) * q)) for q in (q_lo, q_hi)]


def predict_pcm_timesplit(contexts, pcm: np.ndarray, threshold: float = 0.5, want_agg: bool = True, parallel: bool = True):
    """Annotate ONE recording on several contexts (one per GPU): -> (stats, agg, cnt, label_idx, start_step, stop_step),
    exactly what ``Context.predict_pcm`` returns on one context."""
    ctx0 = contexts[0]
    P = ctx0.params
    pcm = np.ascontiguousarray(pcm)
    chunks = plan_chunks(pcm.size, len(contexts), P.hop, P.snippet_len)
    ctxs = list(contexts[: len(chunks)])
    T = 1 + pcm.size // P.hop
    N = sum(c.n_snippets for c in chunks)
    pool = ThreadPoolExecutor(max_workers=len(ctxs)) if parallel and len(ctxs) > 1 else None

    def each(fn):
        if pool is None:
            return [fn(c, k) for c, k in zip(ctxs, chunks)]
        return list(pool.map(lambda ck: fn(*ck), zip(ctxs, chunks)))

    try:
        def stage_spectrogram(c, k):
            c.upload_pcm(pcm[k.sample0 : k.sample1])
            return c.chunk_spectrogram(k.own_row0, k.own_row1)

        pmax = max(each(stage_spectrogram))                      # float32 values: the maximum is exact
        each(lambda c, k: c.chunk_select_begin(pmax))
        n_freq = P.band_hi - P.band_lo
        ranks = nearest_ranks(T * n_freq, P.q_lo, P.q_hi)
        prefixes = [0, 0]
        for pass_ in range(3):
            hists = each(lambda c, k: c.chunk_histogram(pass_, k.own_row0, k.own_row1, prefixes))
            ranks, prefixes = locate_ranks(hists, ranks, prefixes, pass_)
        stats = each(lambda c, k: c.chunk_select_end(prefixes))
        preds = each(lambda c, k: c.forward_resident(k.local_snippet0, k.n_snippets))
    finally:
        if pool is not None:
            pool.shutdown()
    preds = np.concatenate(preds, axis=0)
    assert preds.shape[0] == N
    agg, cnt, lab, sta, sto = ctx0.postprocess(preds, T, threshold=threshold, want_agg=want_agg)
    st = stats[0]
    st.n_frames = T
    st.rank_lo, st.rank_hi = nearest_ranks(T * n_freq, P.q_lo, P.q_hi)
    return st, agg, cnt, lab, sta, sto
