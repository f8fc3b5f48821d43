"""One recording split by time (orcai_b200/timesplit.py): plan geometry and the host side of the radix select on the CPU;
bit-identity with the one-context path on the GPU."""

import numpy as np
import pytest

from orcai_b200.timesplit import key_to_float, locate_ranks, nearest_ranks, plan_chunks


def _float_key(v: np.ndarray) -> np.ndarray:
    u = v.astype(np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


@pytest.mark.parametrize("seconds,n_chunks", [(4.0, 3), (10.0, 2), (10.0, 4), (61.3, 8), (7.9, 1), (5.0, 16)])
def test_plan_covers_every_frame_and_snippet_once(seconds, n_chunks):
    n = int(seconds * 48000) + 123
    T = 1 + n // 256
    N = (T - 736) // 368 + 1
    chunks = plan_chunks(n, n_chunks)
    assert 1 <= len(chunks) <= min(n_chunks, N)
    assert [c.snippet0 for c in chunks] == list(np.cumsum([0] + [c.n_snippets for c in chunks[:-1]]))
    assert sum(c.n_snippets for c in chunks) == N
    owned = np.zeros(T, int)
    for g, c in enumerate(chunks):
        assert c.sample0 == 256 * c.frame0 and 0 <= c.sample0 < c.sample1 <= n
        t_local = 1 + (c.sample1 - c.sample0) // 256
        owned[c.frame0 + c.own_row0 : c.frame0 + c.own_row1] += 1
        # rows computed from real samples only: all but local row 0 (unless it is the recording's frame 0) and the last local
        # row (unless the chunk reaches the end of the recording)
        first_exact = 0 if c.frame0 == 0 else 1
        last_exact = t_local - 1 if c.sample1 == n else t_local - 2
        assert first_exact <= c.own_row0 and c.own_row1 - 1 <= last_exact
        r0 = 368 * c.local_snippet0
        r1 = 368 * (c.local_snippet0 + c.n_snippets - 1) + 736
        assert first_exact <= r0 and r1 - 1 <= last_exact
        assert c.frame0 + r0 == 368 * c.snippet0          # local snippets are the global ones
        assert (t_local - 736) // 368 + 1 >= c.local_snippet0 + c.n_snippets
    assert (owned == 1).all()
    with pytest.raises(ValueError):
        plan_chunks(100_000, 2)


def test_host_radix_select_equals_numpy_percentiles():
    rng = np.random.default_rng(11)
    v = np.maximum(rng.normal(-50, 20, size=171 * 4001).astype(np.float32), -80).astype(np.float32)
    v[rng.integers(0, v.size, 500)] = -80.0
    keys = _float_key(v)
    parts = np.array_split(keys, 5)                     # five "chunks"
    q_lo, q_hi = 0.01, 0.9990000000000001
    ranks = nearest_ranks(v.size, q_lo, q_hi)
    prefixes = [0, 0]
    for pass_ in range(3):
        hists = []
        for part in parts:
            h = np.zeros((2, 2048), np.uint64)
            for r in range(2):
                if pass_ == 0:
                    sel, dig = part, part >> 21
                elif pass_ == 1:
                    sel = part[(part >> 21) == (prefixes[r] >> 21)]
                    dig = (sel >> 10) & 2047
                else:
                    sel = part[(part >> 10) == (prefixes[r] >> 10)]
                    dig = sel & 1023
                if pass_ == 0 and r == 1:
                    continue
                np.add.at(h[r], dig.astype(np.int64), 1)
            hists.append(h)
        ranks, prefixes = locate_ranks(hists, ranks, prefixes, pass_)
    lo, hi = key_to_float(prefixes[0]), key_to_float(prefixes[1])
    assert lo == np.percentile(v, 100 * q_lo, method="nearest") and hi == np.percentile(v, 100 * q_hi, method="nearest")
    assert ranks == [r - int((np.sort(v) < x).sum()) for r, x in zip(nearest_ranks(v.size, q_lo, q_hi), (lo, hi))]


@pytest.mark.gpu
@pytest.mark.parametrize("n_chunks,net_path", [(3, 0), (2, 3), (5, 3)])
def test_time_split_is_bit_identical_to_one_context(ctx, params, n_chunks, net_path):
    """Chunks on several contexts of ONE GPU (what runs on several GPUs in production) against Context.predict_pcm."""
    from orcai_b200._lib import Context
    from orcai_b200.synth import synth_pcm16
    from orcai_b200.timesplit import predict_pcm_timesplit
    from orcai_b200.weights import synthetic_weights

    P, S = params
    pcm = synth_pcm16(31.7, seed=515, calls_per_minute=40.0)
    W = synthetic_weights(P, S, seed=1234)
    extra = [Context(P, S, device=0) for _ in range(n_chunks - 1)]
    try:
        for c in extra:
            c.load_weights(W)
        if net_path == 3:
            for c in [ctx] + extra:
                c.calibrate()
        for c in [ctx] + extra:
            c.set_option("net_path", net_path)
        one = ctx.predict_pcm(pcm)
        for parallel in (False, True):
            split = predict_pcm_timesplit([ctx] + extra, pcm, parallel=parallel)
            assert (split[0].lo, split[0].hi, split[0].db_ref, split[0].ref_power) == (one[0].lo, one[0].hi, one[0].db_ref, one[0].ref_power)
            assert (split[0].n_frames, split[0].rank_lo, split[0].rank_hi) == (one[0].n_frames, one[0].rank_lo, one[0].rank_hi)
            for a, b in zip(split[1:], one[1:]):
                np.testing.assert_array_equal(a, b)
        assert len(one[3]) > 0
    finally:
        for c in [ctx] + extra:
            c.set_option("net_path", 0)
        for c in extra:
            c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seconds", [3.93, 7.9, 9.87, 23.456, 44.4])
def test_time_split_ragged_lengths_and_more_chunks_than_snippets(ctx, params, seconds):
    """Recordings of 1 .. 22 snippets, lengths that are not multiples of the hop, float32 samples, up to 4 chunks."""
    from orcai_b200._lib import Context, OrcaiError
    from orcai_b200.synth import pcm16_to_float, synth_pcm16
    from orcai_b200.timesplit import predict_pcm_timesplit
    from orcai_b200.weights import synthetic_weights

    P, S = params
    pcm = synth_pcm16(seconds, seed=int(seconds * 1000), calls_per_minute=60.0)[: int(seconds * 48000) - 77]
    W = synthetic_weights(P, S, seed=1234)
    extra = [Context(P, S, device=0) for _ in range(3)]
    try:
        for c in extra:
            c.load_weights(W)
        T = 1 + pcm.size // 256
        if T < 736:
            with pytest.raises(ValueError):
                predict_pcm_timesplit([ctx] + extra, pcm)
            return
        for samples in (pcm, pcm16_to_float(pcm)):
            one = ctx.predict_pcm(samples)
            split = predict_pcm_timesplit([ctx] + extra, samples)
            assert (split[0].lo, split[0].hi, split[0].db_ref, split[0].n_frames) == (one[0].lo, one[0].hi, one[0].db_ref, one[0].n_frames)
            for a, b in zip(split[1:], one[1:]):
                np.testing.assert_array_equal(a, b)
    finally:
        for c in extra:
            c.close()


@pytest.mark.gpu
def test_predict_single_wav_on_two_devices_writes_the_same_file(tmp_path, params, monkeypatch):
    """`orcai predict one.wav` with ORCAI_B200_DEVICES=0,1: time chunks on two GPUs, same label file as on one."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json

    from orcai_b200 import predict as opredict
    from orcai_b200.synth import synth_pcm16
    from orcai_b200.wavio import write_wav_pcm16 as write_wav

    P, S = params
    d = tmp_path / "orcai-V1"
    d.mkdir()
    (d / "orcai_parameter.json").write_text(json.dumps(P))
    (d / "model_shape.json").write_text(json.dumps(S))
    monkeypatch.setenv("ORCAI_B200_SYNTHETIC_WEIGHTS", "1234")
    wav = tmp_path / "rec.wav"
    write_wav(wav, synth_pcm16(45.0, seed=99, calls_per_minute=40.0), 48000)
    monkeypatch.delenv("ORCAI_B200_DEVICES", raising=False)
    opredict.predict(wav, model_dir=d, output_path=str(tmp_path / "one.txt"), save_probabilities=True, verbosity=0)
    monkeypatch.setenv("ORCAI_B200_DEVICES", "0,1")
    opredict.predict(wav, model_dir=d, output_path=str(tmp_path / "two.txt"), save_probabilities=True, verbosity=0)
    assert (tmp_path / "one.txt").read_bytes() == (tmp_path / "two.txt").read_bytes()
    assert len((tmp_path / "one.txt").read_text().splitlines()) > 3
    import gzip

    assert gzip.open(tmp_path / "one_probabilities.csv.gz").read() == gzip.open(tmp_path / "two_probabilities.csv.gz").read()


class _OracleChunkContext:
    """numpy stand-in for a device context (test infrastructure, built on oracle/): the chunk_* / forward_resident /
    postprocess calls of _lib.Context with the oracle's arithmetic, so that the HOST orchestration of the time split
    (plan, halo handling, maxima, host-driven radix select, gather) is checked on the CPU against the reference's own
    whole-recording order of operations."""

    def __init__(self, P, calls_log):
        from types import SimpleNamespace

        sp = P["spectrogram"]
        self.sp = sp
        self.params = SimpleNamespace(hop=sp["n_overlap"], snippet_len=736, band_lo=0, band_hi=171, q_lo=sp["quantiles"][0], q_hi=sp["quantiles"][1],
                                      n_blocks=4, n_labels=7)
        self.log = calls_log

    def upload_pcm(self, pcm):
        from orcai_b200.synth import pcm16_to_float

        self.y = pcm16_to_float(np.asarray(pcm))

    def chunk_spectrogram(self, r0, r1):
        from oracle import spectrogram_oracle as so

        S = so.stft_complex64(self.y, self.sp["nfft"], self.sp["n_overlap"])
        power = np.square(np.abs(S))                               # float32 |S|^2, (257, T_local)
        self.raw = (np.float32(10.0) * np.log10(np.maximum(np.float32(1e-10), power)))[:171].T.copy()   # unshifted dB, (T_local, 171)
        self.log.append(("spectrogram", r0, r1, self.raw.shape[0]))
        return float(power[:, r0:r1].max()) if r1 > r0 else 0.0

    def chunk_select_begin(self, pmax):
        self.db_ref = np.float32(10.0) * np.log10(np.maximum(np.float32(1e-10), np.float32(pmax)))

    def _keys(self, r0, r1):
        v = np.maximum(self.raw[r0:r1] - self.db_ref, np.float32(-80.0)).astype(np.float32).ravel()
        return _float_key(v)

    def chunk_histogram(self, pass_, r0, r1, prefix):
        k = self._keys(r0, r1)
        h = np.zeros((2, 2048), np.uint64)
        for r in range(2):
            if pass_ == 0:
                if r == 0:
                    np.add.at(h[0], (k >> 21).astype(np.int64), 1)
                continue
            sel = k[(k >> 21) == (int(prefix[r]) >> 21)] if pass_ == 1 else k[(k >> 10) == (int(prefix[r]) >> 10)]
            np.add.at(h[r], (((sel >> 10) & 2047) if pass_ == 1 else (sel & 1023)).astype(np.int64), 1)
        return h

    def chunk_select_end(self, keys):
        from types import SimpleNamespace

        self.lo, self.hi = key_to_float(int(keys[0])), key_to_float(int(keys[1]))
        return SimpleNamespace(lo=float(self.lo), hi=float(self.hi), db_ref=float(self.db_ref), n_frames=0, rank_lo=0, rank_hi=0)

    def forward_resident(self, first, n):
        # a stand-in "network": per snippet, per output step, 7 statistics of the normalised spectrogram (sensitive to every frame)
        v = np.maximum(self.raw - self.db_ref, np.float32(-80.0))
        spec = (np.clip(v, self.lo, self.hi) - self.lo) / (self.hi - self.lo)
        out = np.empty((n, 46, 7), np.float32)
        for s in range(n):
            snip = spec[368 * (first + s) : 368 * (first + s) + 736].reshape(46, 16, 171)
            out[s] = np.stack([snip[:, :, 24 * j : 24 * j + 24].mean(axis=(1, 2)) for j in range(7)], axis=1)
        return out

    def postprocess(self, preds, T, threshold=0.5, want_agg=True):
        from oracle import postprocess_oracle as po

        agg, cnt = po.aggregate_predictions(preds, T, 736, 4, 7)
        return agg, cnt, np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)


@pytest.mark.parametrize("seconds,n_chunks", [(12.0, 2), (31.7, 3), (31.7, 5), (9.0, 4)])
def test_time_split_orchestration_on_cpu(params, seconds, n_chunks):
    """predict_pcm_timesplit driven with numpy stand-in contexts == the oracle's whole-recording pipeline: identical reference
    level, percentiles (the host-driven radix select over chunk histograms), predictions and aggregates, for every chunk count."""
    from oracle import postprocess_oracle as po, spectrogram_oracle as so
    from orcai_b200.synth import pcm16_to_float, synth_pcm16
    from orcai_b200.timesplit import predict_pcm_timesplit

    P, S = params
    pcm = synth_pcm16(seconds, seed=4242, calls_per_minute=40.0)[: int(seconds * 48000) - 311]
    db, freqs, _ = so.calculate_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    spec_ref, lo, hi = so.preprocess_spectrogram(db, freqs, P["spectrogram"])
    log = []
    ctxs = [_OracleChunkContext(P, log) for _ in range(n_chunks)]
    st, agg, cnt, *_ = predict_pcm_timesplit(ctxs, pcm, parallel=False)
    assert np.float32(st.lo) == lo and np.float32(st.hi) == hi
    T = spec_ref.shape[0]
    assert st.n_frames == T
    # the stand-in network on the oracle's whole-recording spectrogram
    N = (T - 736) // 368 + 1
    want = np.empty((N, 46, 7), np.float32)
    for s in range(N):
        snip = spec_ref[368 * s : 368 * s + 736].reshape(46, 16, 171)
        want[s] = np.stack([snip[:, :, 24 * j : 24 * j + 24].mean(axis=(1, 2)) for j in range(7)], axis=1)
    agg_ref, cnt_ref = po.aggregate_predictions(want, T, 736, 4, 7)
    np.testing.assert_array_equal(cnt, cnt_ref)
    np.testing.assert_allclose(agg, agg_ref, rtol=0, atol=2e-6)    # dB through log10 of float32 powers: last-ulp differences only
    assert len([e for e in log if e[0] == "spectrogram"]) == min(n_chunks, N)
