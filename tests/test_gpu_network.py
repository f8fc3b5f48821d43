"""GPU parity: snippet batcher + orcai-V1 forward through the C ABI vs the torch-CPU oracle."""

import numpy as np
import pytest

from oracle import network_oracle, postprocess_oracle as po, spectrogram_oracle as so
from orcai_b200.synth import pcm16_to_float, synth_pcm16
from orcai_b200.weights import synthetic_weights

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-3  # north_star: per-frame probabilities within 1e-3 absolute (fp32 path: measured 1e-6)
# The DEFAULT network path (net_path 4, "precise"): every GEMM on the fp16 tensor cores as the three-term split
# A_hi*W_hi + A_lo*W_hi + A_hi*W_lo with fp32 accumulation.  Held ten times tighter than the gate; measured 4e-6 on the cases
# below and 2e-5 over a whole recording (tools/gpu_check_precise.py, profiles/r02*).
PRECISE_TOL = 1e-4
# OPT-IN fast path (net_path 3, ORCAI_B200_PRECISION=fast): single fp16 operands, biases calibrated against the weight rounding
# (orcai_calibrate on the built-in synthetic recording).  It does NOT meet the 1e-3 gate: max 0.9e-3 on the cases below but
# 2.7e-3 over a whole 1-h recording, mean 1.0e-4 (tools/precision_plan.py attributes it: every fp16 weight tensor and stored
# activation costs 2e-4 .. 2e-3).  These bounds document that path; nothing shipped by default depends on them.
FAST_TOL = 2.5e-3
FAST_MEAN_TOL = 3e-4


def test_golden_probabilities(ctx, golden_dir):
    g = np.load(golden_dir / "network_seed1234.npz")
    x = np.random.default_rng(5).random((2, 736, 171), dtype=np.float32)
    out = ctx.forward_host(x)
    assert out.shape == (2, 46, 7) and out.dtype == np.float32
    assert np.abs(out - g["probs"]).max() <= PROB_TOL


def test_model_predict_boundary(ctx, params):
    """model.predict((N,736,171,1)) -> (N,46,7), Keras-style, incl. batches that do not divide the chunk."""
    from orcai_b200.model import OrcaiModel

    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    model = OrcaiModel(P, S, W, device=0, precision="reference")
    x = np.random.default_rng(6).random((5, 736, 171, 1), dtype=np.float32)
    ref = network_oracle.forward(x, W)
    model.ctx.set_option("chunk", 2)  # 5 snippets in chunks of 2 -> ragged last chunk
    out = model.predict(x, verbose=0)
    model.ctx.set_option("chunk", 128)
    assert out.shape == (5, 46, 7) and np.abs(out - ref).max() <= PROB_TOL
    one = model.predict(x[:1])
    np.testing.assert_array_equal(one, out[:1])  # a snippet's result does not depend on its batch
    with pytest.raises(ValueError):
        model.predict(np.zeros((0, 736, 171, 1), np.float32))
    with pytest.raises(ValueError):
        model.ctx.forward_host(np.zeros((1, 700, 171), np.float32))


def test_strided_snippet_batcher_matches_copies(ctx, params):
    """Snippets cut as strided windows of the resident recording == the reference's materialised copies."""
    P, S = params
    pcm = synth_pcm16(14.0, seed=31)
    spec, st = ctx.spectrogram(pcm)
    n = int((st.n_frames - 736) // 368 + 1)
    assert n == 6
    resident = ctx.forward_resident(0, n)           # reads the raw dB buffer, normalises on load
    tail = ctx.forward_resident(2, 3)
    copies = ctx.forward_host(po.cut_snippets(spec, 736))  # (N,736,171) materialised from the device's own spectrogram
    np.testing.assert_array_equal(resident, copies)
    np.testing.assert_array_equal(tail, resident[2:5])
    # against the oracle end to end (oracle spectrogram -> oracle network)
    db, f, _ = so.calculate_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    spec_ref, _, _ = so.preprocess_spectrogram(db, f, P["spectrogram"])
    ref = network_oracle.forward(po.cut_snippets(spec_ref, 736), synthetic_weights(P, S, seed=1234))
    assert np.abs(resident - ref).max() <= PROB_TOL
    with pytest.raises(Exception):
        ctx.forward_resident(4, 3)  # past the last snippet


def test_other_weights_seed(ctx, params):
    P, S = params
    W = synthetic_weights(P, S, seed=99)
    ctx.load_weights(W)
    try:
        x = np.random.default_rng(8).random((3, 736, 171), dtype=np.float32)
        assert np.abs(ctx.forward_host(x) - network_oracle.forward(x, W)).max() <= PROB_TOL
        bad = dict(W)
        bad.pop("dense2/bias")
        with pytest.raises(Exception, match="dense2/bias"):
            ctx.load_weights(bad)
    finally:
        ctx.load_weights(synthetic_weights(P, S, seed=1234))


@pytest.mark.parametrize("tail_path", [0, 1])
def test_fast_path_fused_tensor_core_blocks(ctx, params, golden_dir, tail_path):
    """net_path 3: fused tcgen05 residual blocks (+ tensor-core or fp32 LSTM tail) against the oracle and the fp32 path."""
    P, S = params
    g = np.load(golden_dir / "network_seed1234.npz")
    x = np.random.default_rng(5).random((2, 736, 171), dtype=np.float32)
    ref32 = ctx.forward_host(x)
    ctx.calibrate()
    ctx.set_option("net_path", 3)
    ctx.set_option("tail_path", tail_path)
    try:
        out = ctx.forward_host(x)
        assert out.shape == (2, 46, 7) and np.isfinite(out).all()
        assert np.abs(out - g["probs"]).max() <= FAST_TOL and np.abs(out - g["probs"]).mean() <= FAST_MEAN_TOL
        assert np.abs(out - ref32).max() <= FAST_TOL
        # ragged chunking and batch independence: 7 snippets in chunks of 3; a snippet's result does not depend on its batch
        x7 = np.random.default_rng(9).random((7, 736, 171), dtype=np.float32)
        full = ctx.forward_host(x7)
        ctx.set_option("chunk", 3)
        chunked = ctx.forward_host(x7)
        ctx.set_option("chunk", 2048)
        np.testing.assert_array_equal(full, chunked)
        np.testing.assert_array_equal(ctx.forward_host(x7[4:5]), full[4:5])
        W = synthetic_weights(P, S, seed=1234)
        ref7 = network_oracle.forward(x7, W)
        assert np.abs(full - ref7).max() <= FAST_TOL
        # the fp32 CUDA-core entry convolution (conv0_path 0) instead of the tensor-core pixel-group one: each variant is held
        # to the oracle; two variants may differ from each other by up to twice that
        ctx.set_option("conv0_path", 0)
        alt = ctx.forward_host(x7)
        ctx.set_option("conv0_path", 1)
        assert np.abs(alt - ref7).max() <= FAST_TOL
        assert np.abs(alt - full).max() <= 2 * FAST_TOL
        # the entry convolution fused into block 1 (conv0_path 2: CUDA-core producer warps fed by a strip-cut spectrogram)
        ctx.set_option("conv0_path", 2)
        fused0 = ctx.forward_host(x7)
        ctx.set_option("chunk", 3)
        np.testing.assert_array_equal(ctx.forward_host(x7), fused0)
        ctx.set_option("chunk", 2048)
        ctx.set_option("conv0_path", 1)
        assert np.abs(fused0 - ref7).max() <= FAST_TOL
        assert np.abs(fused0 - full).max() <= 2 * FAST_TOL
    finally:
        ctx.set_option("conv0_path", 1)
        ctx.set_option("net_path", 0)
        ctx.set_option("tail_path", 1)
        ctx.set_option("chunk", 128)


def test_fast_path_resident_recording(ctx, params):
    """Fast path on a resident recording (normalise-on-load from the raw dB buffer) == on materialised snippets; oracle within FAST_TOL."""
    P, S = params
    pcm = synth_pcm16(14.0, seed=31)
    spec, st = ctx.spectrogram(pcm)
    n = int((st.n_frames - 736) // 368 + 1)
    ctx.calibrate()                      # replaces the resident recording ...
    spec, st = ctx.spectrogram(pcm)      # ... so make the test recording resident again
    ctx.set_option("net_path", 3)
    try:
        resident = ctx.forward_resident(0, n)
        copies = ctx.forward_host(po.cut_snippets(spec, 736))
        np.testing.assert_array_equal(resident, copies)
        db, f, _ = so.calculate_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
        spec_ref, _, _ = so.preprocess_spectrogram(db, f, P["spectrogram"])
        ref = network_oracle.forward(po.cut_snippets(spec_ref, 736), synthetic_weights(P, S, seed=1234))
        assert np.abs(resident - ref).max() <= FAST_TOL and np.abs(resident - ref).mean() <= FAST_MEAN_TOL
    finally:
        ctx.set_option("net_path", 0)


def test_prefetch_swap_pipeline(ctx):
    """orcai_prefetch_pcm / orcai_swap_pcm: streaming a table of recordings gives exactly the one-call results."""
    recs = [synth_pcm16(secs, seed=70 + k, calls_per_minute=40.0) for k, secs in enumerate((9.0, 12.5, 8.0))]
    one = [ctx.predict_pcm(r) for r in recs]
    streamed = list(ctx.predict_stream(recs))
    assert len(streamed) == 3
    for a, b in zip(one, streamed):
        assert a[0].n_frames == b[0].n_frames and a[0].lo == b[0].lo and a[0].hi == b[0].hi
        np.testing.assert_array_equal(a[1], b[1])
        for i in (2, 3, 4, 5):
            np.testing.assert_array_equal(a[i], b[i])
    with pytest.raises(Exception, match="prefetch"):
        ctx.swap_pcm()


def test_two_predict_calls_in_flight(ctx):
    """orcai_predict_resident_begin / _end: recording k+1 is enqueued before recording k is collected; every result - statistics,
    aggregates, segments - equals the one-call result bit for bit, for recordings of different lengths (buffers are re-allocated
    under a call in flight), with and without aggregates, on the fp32 and the default network path; misuse fails loudly."""
    recs = [synth_pcm16(secs, seed=170 + k, calls_per_minute=40.0) for k, secs in enumerate((9.0, 21.0, 8.0, 30.0, 12.0))]
    for path in (0, 4):
        ctx.set_option("net_path", path)
        try:
            one = [ctx.predict_pcm(r) for r in recs]
            got = []
            tokens = []
            ctx.prefetch_pcm(recs[0])
            for k, r in enumerate(recs):
                ctx.swap_pcm()
                tokens.append(ctx.predict_begin(r.size, want_agg=(k % 2 == 0)))
                assert ctx.predict_in_flight() == len(tokens)
                if len(tokens) == 2:
                    got.append(ctx.predict_end(tokens.pop(0)))
                if k + 1 < len(recs):
                    ctx.prefetch_pcm(recs[k + 1])      # into the buffer of the recording collected above
            while tokens:
                got.append(ctx.predict_end(tokens.pop(0)))
            assert ctx.predict_in_flight() == 0
            for k, (a, b) in enumerate(zip(one, got)):
                assert (a[0].n_frames, a[0].lo, a[0].hi, a[0].db_ref, a[0].rank_lo, a[0].rank_hi) == (b[0].n_frames, b[0].lo, b[0].hi, b[0].db_ref, b[0].rank_lo, b[0].rank_hi)
                if k % 2 == 0:
                    np.testing.assert_array_equal(a[1], b[1])
                    np.testing.assert_array_equal(a[2], b[2])
                else:
                    assert b[1] is None and b[2] is None
                for i in (3, 4, 5):
                    np.testing.assert_array_equal(a[i], b[i])
                assert len(a[3]) > 0
        finally:
            ctx.set_option("net_path", 0)
    with pytest.raises(Exception, match="no predict call in flight"):
        ctx.predict_end((1, 7, 1024, False))
    ctx.upload_pcm(recs[0])
    t1 = ctx.predict_begin(recs[0].size)
    t2 = ctx.predict_begin(recs[0].size)
    with pytest.raises(Exception, match="already in flight"):
        ctx.predict_begin(recs[0].size)
    a, b = ctx.predict_end(t1), ctx.predict_end(t2)
    for i in (1, 2, 3, 4, 5):
        np.testing.assert_array_equal(a[i], b[i])
    ctx.upload_pcm(synth_pcm16(1.0, seed=3))
    with pytest.raises(Exception, match="shorter than one snippet"):
        ctx.predict_begin(48000)
    assert ctx.predict_in_flight() == 0


def test_calibration_reduces_weight_rounding_error(ctx, params):
    """orcai_calibrate: deterministic, cleared by load_weights, and it shrinks the fast path's deviation on unseen audio."""
    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    pcm = synth_pcm16(40.0, seed=4242, calls_per_minute=20.0)
    spec, _ = ctx.spectrogram(pcm)
    xa = po.cut_snippets(spec, 736)[:6]
    ref = network_oracle.forward(xa, W)
    ctx.load_weights(W)                  # clears any calibration
    ctx.set_option("net_path", 3)
    try:
        raw = ctx.forward_host(xa)
        ctx.calibrate()
        cal1 = ctx.forward_host(xa)
        ctx.calibrate()
        cal2 = ctx.forward_host(xa)
        np.testing.assert_array_equal(cal1, cal2)
        e_raw, e_cal = np.abs(raw - ref).max(), np.abs(cal1 - ref).max()
        assert e_cal <= FAST_TOL and e_cal < e_raw
        ctx.load_weights(W)
        np.testing.assert_array_equal(ctx.forward_host(xa), raw)
        with pytest.raises(Exception, match="shorter than one snippet"):
            ctx.calibrate(synth_pcm16(2.0, seed=1))
    finally:
        ctx.set_option("net_path", 0)


def test_fast_path_edge_lengths(ctx, params):
    """One snippet exactly (T = 736), one snippet plus an uncovered tail, and an odd multi-snippet length on the fast path."""
    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    ctx.calibrate()
    for n_samples in (735 * 256, 735 * 256 + 367 * 256 + 17, (736 + 368 * 4 + 5) * 256 + 3):
        pcm = synth_pcm16(n_samples / 48000.0 + 0.01, seed=123, calls_per_minute=60.0)[:n_samples]
        T = 1 + n_samples // 256
        n = (T - 736) // 368 + 1
        ctx.set_option("net_path", 0)
        ref = ctx.predict_pcm(pcm)
        ctx.set_option("net_path", 3)
        try:
            got = ctx.predict_pcm(pcm)
        finally:
            ctx.set_option("net_path", 0)
        assert got[0].n_frames == T and got[1].shape == (T // 16, 7)
        np.testing.assert_array_equal(got[2], ref[2])                      # overlap counts
        assert np.abs(got[1] - ref[1]).max() <= FAST_TOL                   # aggregated probabilities vs the fp32 path
        assert (got[1][23 * (n - 1) + 46:] == 0).all()                     # steps no snippet covers stay 0


def test_silent_recording_labels_nothing(ctx):
    """All-zero audio: the reference's normalisation is 0/0 = NaN everywhere (spectrogram.py:81-83), Keras propagates it and
    nothing is labelled; both network paths reproduce that (NaN probabilities, no segments, no crash)."""
    ctx.calibrate()
    for path in (0, 3):
        ctx.set_option("net_path", path)
        try:
            st, agg, cnt, lab, sta, sto = ctx.predict_pcm(np.zeros(48000 * 6, np.int16))
        finally:
            ctx.set_option("net_path", 0)
        assert len(lab) == 0 and len(sta) == 0 and cnt.max() == 2
        assert np.isnan(agg[cnt > 0]).all()


def test_block1_n_widened_variant(ctx, params):
    """block1_path 1 (net_fused_w.cuh): the three dx taps as column blocks of one MMA + shuffle epilogue; same result as the
    per-tap kernel up to fp16 storage rounding (the fp32 accumulation order differs), same tolerance against the oracle."""
    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    x = np.random.default_rng(21).random((4, 736, 171), dtype=np.float32)
    ref = network_oracle.forward(x, W)
    ctx.calibrate()
    ctx.set_option("net_path", 3)
    try:
        base = ctx.forward_host(x)
        ctx.set_option("block1_path", 1)
        wide = ctx.forward_host(x)
        ctx.set_option("chunk", 3)
        np.testing.assert_array_equal(ctx.forward_host(x), wide)
        ctx.set_option("chunk", 2048)
        # same error class as the per-tap kernel on the same input: the MEAN deviation is the stable statistic (within 10 %); the
        # maximum over 1 288 probabilities is one noise realisation per kernel (here 2.0e-3 vs 2.6e-3, on other inputs the other
        # way round) and is held to the documented whole-recording maximum of the 16-bit path plus margin
        e_base, e_wide = np.abs(base - ref), np.abs(wide - ref)
        assert e_wide.mean() <= max(FAST_MEAN_TOL, 1.10 * e_base.mean())
        assert e_wide.max() <= 3.5e-3 and np.abs(wide - base).max() <= 3.5e-3
        b0 = ctx.debug_stage(x[:2], 1)
        ctx.set_option("block1_path", 0)
        b1 = ctx.debug_stage(x[:2], 1)
        assert np.abs(b0 - b1).max() <= 2 ** -7      # one fp16 ulp at |x| < 16
    finally:
        ctx.set_option("block1_path", 0)
        ctx.set_option("net_path", 0)
        ctx.set_option("chunk", 128)


# ------------------------------------------------------------------------------------------------
# the default path: split-fp16 tensor-core GEMMs (net_path 4)
# ------------------------------------------------------------------------------------------------
def _bn_matched(W, P, seconds=20.0):
    """Weights whose BatchNorm moving statistics match their activations (as training leaves them): the harder case for 16-bit
    arithmetic."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import precision_study as ps

    spec, _, _ = so.make_spectrogram(pcm16_to_float(synth_pcm16(seconds, seed=20251018)), P["spectrogram"])
    return ps.calibrate_bn(W, po.cut_snippets(spec, 736)[:2])


def test_precise_path_against_oracle(ctx, params, golden_dir):
    """net_path 4 against the golden probabilities and the oracle: 1e-4 (gate 1e-3); chunking and batch composition change nothing."""
    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    g = np.load(golden_dir / "network_seed1234.npz")
    x = np.random.default_rng(5).random((2, 736, 171), dtype=np.float32)
    ctx.set_option("net_path", 4)
    try:
        out = ctx.forward_host(x)
        assert out.shape == (2, 46, 7) and out.dtype == np.float32 and np.isfinite(out).all()
        assert np.abs(out - g["probs"]).max() <= PRECISE_TOL
        x7 = np.random.default_rng(9).random((7, 736, 171), dtype=np.float32)
        full = ctx.forward_host(x7)
        assert np.abs(full - network_oracle.forward(x7, W)).max() <= PRECISE_TOL
        ctx.set_option("chunk", 3)
        np.testing.assert_array_equal(ctx.forward_host(x7), full)
        ctx.set_option("chunk", 1024)
        np.testing.assert_array_equal(ctx.forward_host(x7[4:5]), full[4:5])
        # stage by stage against the oracle's intermediates (un-rectified block outputs, fp32)
        _, inter = network_oracle.forward(x[:1], W, return_intermediates=True)
        for stage, key in ((1, "block1"), (2, "block2"), (3, "block3"), (4, "block4"), (5, "final")):
            want = np.transpose(inter[key], (0, 2, 3, 1))
            got = ctx.debug_stage(x[:1], stage)
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max() + 1e-5, key
    finally:
        ctx.set_option("net_path", 0)
        ctx.set_option("chunk", 128)


def test_precise_shared_interior_is_bit_identical(ctx, params):
    """Resident recordings run the trunk once over the recording as a tall image and recompute only the rows that feel a
    snippet's own border (predict.py:252-261 cuts overlapping windows).  That must not change a single bit against the
    snippet-by-snippet evaluation, for any chunking and any sub-range, and both must sit on the fp32 path."""
    P, S = params
    pcm = synth_pcm16(61.0, seed=77, calls_per_minute=40.0)
    spec, st = ctx.spectrogram(pcm)
    n = int((st.n_frames - 736) // 368 + 1)
    assert n >= 29
    ref32 = ctx.forward_resident(0, n)
    ctx.set_option("net_path", 4)
    try:
        ctx.set_option("precise_tall", 0)
        per_snippet = ctx.forward_resident(0, n)
        ctx.set_option("precise_tall", 1)
        tall = ctx.forward_resident(0, n)
        np.testing.assert_array_equal(tall, per_snippet)
        ctx.set_option("chunk", 5)                                   # ragged chunks: the tall images overlap by one snippet shift
        np.testing.assert_array_equal(ctx.forward_resident(0, n), per_snippet)
        ctx.set_option("chunk", 1024)
        np.testing.assert_array_equal(ctx.forward_resident(3, 7), per_snippet[3:10])
        np.testing.assert_array_equal(ctx.forward_resident(n - 1, 1), per_snippet[n - 1:])
        assert np.abs(tall - ref32).max() <= PRECISE_TOL
        # the strided snippet batcher == materialised copies (last: a host batch replaces the resident spectrogram)
        np.testing.assert_array_equal(ctx.forward_host(po.cut_snippets(spec, 736)), per_snippet)
    finally:
        ctx.set_option("precise_tall", 1)
        ctx.set_option("net_path", 0)
        ctx.set_option("chunk", 128)


@pytest.mark.parametrize("bn_matched", [False, True])
def test_precise_path_one_hour_inside_the_gate(ctx, params, bn_matched):
    """A 1-h recording (1 833 snippets, 590 k probabilities), seeded and BatchNorm-matched weights: the default path stays
    within 1e-3 of the fp32 path (which is held to the oracle at 1e-6) - measured 2e-5 - and labels the same segments."""
    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    if bn_matched:
        W = _bn_matched(W, P)
    pcm = synth_pcm16(3600.0, seed=20251018)
    try:
        ctx.load_weights(W)
        ctx.set_option("chunk", 1024)
        ctx.upload_pcm(pcm)
        st = ctx.spectrogram_resident(normalise=False)
        n = int((st.n_frames - 736) // 368 + 1)
        assert n == 1833
        ref32 = ctx.forward_resident(0, n)
        seg32 = ctx.predict_pcm(pcm, resident=True)
        ctx.set_option("net_path", 4)
        got = ctx.forward_resident(0, n)
        seg = ctx.predict_pcm(pcm, resident=True)
        dev = np.abs(got - ref32)
        assert dev.max() <= PROB_TOL, dev.max()
        assert dev.max() <= PRECISE_TOL, dev.max()
        assert np.abs(seg[1] - seg32[1]).max() <= PRECISE_TOL
        # identical thresholded masks -> identical segments (labels, starts, stops); a label whose aggregated probability comes
        # within 1e-5 of the threshold somewhere in the hour may flip there (both are valid fp32 evaluations of the same graph)
        same_mask = ((seg[1] > 0.25) == (seg32[1] > 0.25)).all(axis=0)
        assert same_mask.sum() >= 5
        got_s = {(int(a), int(b), int(c)) for a, b, c in zip(seg[3], seg[4], seg[5])}
        ref_s = {(int(a), int(b), int(c)) for a, b, c in zip(seg32[3], seg32[4], seg32[5])}
        for lab in np.flatnonzero(same_mask):
            assert {x for x in got_s if x[0] == lab} == {x for x in ref_s if x[0] == lab}
        assert len(got_s ^ ref_s) <= 4 and ((seg[1] > 0.25) != (seg32[1] > 0.25)).mean() <= 1e-5
    finally:
        ctx.set_option("net_path", 0)
        ctx.set_option("chunk", 128)
        ctx.load_weights(synthetic_weights(P, S, seed=1234))


def test_precise_path_edge_lengths_and_silence(ctx):
    """One snippet exactly, one snippet plus an uncovered tail, an odd multi-snippet length, and all-zero audio on the default path."""
    for n_samples in (735 * 256, 735 * 256 + 367 * 256 + 17, (736 + 368 * 4 + 5) * 256 + 3):
        pcm = synth_pcm16(n_samples / 48000.0 + 0.01, seed=123, calls_per_minute=60.0)[:n_samples]
        T = 1 + n_samples // 256
        n = (T - 736) // 368 + 1
        ctx.set_option("net_path", 0)
        ref = ctx.predict_pcm(pcm)
        ctx.set_option("net_path", 4)
        try:
            got = ctx.predict_pcm(pcm)
        finally:
            ctx.set_option("net_path", 0)
        assert got[0].n_frames == T and got[1].shape == (T // 16, 7)
        np.testing.assert_array_equal(got[2], ref[2])
        assert np.abs(got[1] - ref[1]).max() <= PRECISE_TOL
        assert (got[1][23 * (n - 1) + 46:] == 0).all()
        for i in (3, 4, 5):
            np.testing.assert_array_equal(got[i], ref[i])
    ctx.set_option("net_path", 4)
    try:
        st, agg, cnt, lab, sta, sto = ctx.predict_pcm(np.zeros(48000 * 6, np.int16))
    finally:
        ctx.set_option("net_path", 0)
    assert len(lab) == 0 and len(sta) == 0 and cnt.max() == 2 and np.isnan(agg[cnt > 0]).all()


def test_one_24_hour_recording_full_path(ctx, params):
    """BASELINE configs[2]: ONE 24-h recording (4.15 G samples, 16.2 M frames, 44 020 snippets) through the whole predict path on
    one GPU - the default network path against the fp32 path on the same device spectrogram: aggregated probabilities within 1e-4,
    overlap counts equal, and wherever the thresholded masks agree (everywhere but probabilities within 1e-5 of the threshold) the
    segments are identical."""
    one_hour = synth_pcm16(3600.0, seed=20251018)
    pcm = np.tile(one_hour, 24)
    del one_hour
    try:
        ctx.set_option("chunk", 1024)
        ctx.upload_pcm(pcm)
        ctx.set_option("net_path", 0)
        ref = ctx.predict_pcm(pcm, resident=True)
        ctx.set_option("net_path", 4)
        got = ctx.predict_pcm(pcm, resident=True)
    finally:
        ctx.set_option("net_path", 0)
        ctx.set_option("chunk", 128)
    T = 1 + pcm.size // 256
    assert got[0].n_frames == T == 16200001 and got[1].shape == (T // 16, 7)
    assert got[0].lo == ref[0].lo and got[0].hi == ref[0].hi and got[0].db_ref == ref[0].db_ref
    np.testing.assert_array_equal(got[2], ref[2])
    assert np.abs(got[1] - ref[1]).max() <= PRECISE_TOL
    thr = 0.25
    same_mask = ((got[1] > thr) == (ref[1] > thr)).all(axis=0)            # per label: every frame on the same side of the threshold
    # the recording is one hour repeated 24 times: a probability that sits within 1e-5 of the threshold flips in every repetition
    assert same_mask.sum() >= 3 and ((got[1] > thr) != (ref[1] > thr)).mean() <= 1e-5
    seg_g = {(int(a), int(b), int(c)) for a, b, c in zip(got[3], got[4], got[5])}
    seg_r = {(int(a), int(b), int(c)) for a, b, c in zip(ref[3], ref[4], ref[5])}
    for lab in np.flatnonzero(same_mask):
        assert {s for s in seg_g if s[0] == lab} == {s for s in seg_r if s[0] == lab}
    assert len(seg_g ^ seg_r) <= 200 and len(seg_g) > 100000
