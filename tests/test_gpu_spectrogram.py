"""GPU parity: K1 (STFT->dB->crop), exact percentile select, K2 normalise, through the C ABI vs the oracle."""

import numpy as np
import pytest

from oracle import spectrogram_oracle as so
from orcai_b200.synth import pcm16_to_float, synth_pcm16

pytestmark = pytest.mark.gpu

SP = {"sampling_rate": 48000, "nfft": 512, "n_overlap": 256, "freq_range": [0, 16000], "quantiles": [0.01, 0.999], "duration": 4}
DB_TOL = 1e-3        # north_star: dB spectrograms within 1e-3 dB
NORM_TOL = 1e-5      # normalised spectrogram within 1e-5 (of a [0,1] range)


def oracle(y):
    db, f, _ = so.calculate_spectrogram(y, SP)
    spec, lo, hi = so.preprocess_spectrogram(db, f, SP)
    return db[:171].T, spec, lo, hi


def check(ctx, pcm, y=None):
    y = pcm16_to_float(pcm) if y is None else y
    spec, st = ctx.spectrogram(pcm)
    db = ctx.read_db(0, spec.shape[0])
    db_ref, spec_ref, lo, hi = oracle(y)
    assert spec.shape == spec_ref.shape and st.n_frames == spec.shape[0]
    assert np.abs(db - db_ref).max() <= DB_TOL
    # the device percentiles are the exact order statistics of the device's own dB array ...
    flat = np.sort(db.ravel())
    assert flat[st.rank_lo] == st.lo and flat[st.rank_hi] == st.hi
    assert st.rank_lo == so.nearest_rank(db.size, 0.01) and st.rank_hi == so.nearest_rank(db.size, 0.999)
    # ... and within the dB tolerance of the oracle's
    assert abs(st.lo - lo) <= DB_TOL and abs(st.hi - hi) <= DB_TOL
    assert np.abs(spec - spec_ref).max() <= max(NORM_TOL, 2 * DB_TOL / float(hi - lo))
    # normalise is bit-exact given the device dB array and the device percentiles
    expect = (np.clip(db, st.lo, st.hi) - np.float32(st.lo)) / (np.float32(st.hi) - np.float32(st.lo))
    np.testing.assert_array_equal(spec, expect.astype(np.float32))
    return spec, st


def test_golden_2s(ctx, golden_dir):
    g = np.load(golden_dir / "spectrogram_2s.npz")
    spec, st = ctx.spectrogram(g["pcm"])
    db = ctx.read_db(0, spec.shape[0])
    assert np.abs(db - g["db_band"]).max() <= DB_TOL
    assert abs(st.lo - g["lo"]) <= DB_TOL and abs(st.hi - g["hi"]) <= DB_TOL
    assert np.abs(spec - g["spec"]).max() <= 1e-4
    assert abs(st.ref_power / g["ref_power"] - 1) < 1e-5


@pytest.mark.parametrize("seconds,seed", [(10.0, 1), (61.3, 2)])
def test_synthetic_recordings(ctx, seconds, seed):
    check(ctx, synth_pcm16(seconds, seed=20251018 + seed))


def test_fast_float32_fft_variant(ctx):
    """The float32-FFT variant of K1 (option stft_f64=0): same pipeline, documented looser tail error."""
    pcm = synth_pcm16(30.0, seed=20251018 + 1)
    db_ref, spec_ref, lo, hi = oracle(pcm16_to_float(pcm))
    ctx.set_option("stft_f64", 0)
    try:
        spec, st = ctx.spectrogram(pcm)
        db = ctx.read_db(0, spec.shape[0])
    finally:
        ctx.set_option("stft_f64", 1)
    err = np.abs(db - db_ref)
    assert err.max() <= 5e-3 and np.quantile(err, 0.9999) <= 3e-4 and err.mean() <= 1e-5
    flat = np.sort(db.ravel())
    assert flat[st.rank_lo] == st.lo and flat[st.rank_hi] == st.hi
    assert abs(st.lo - lo) <= 2e-3 and abs(st.hi - hi) <= 2e-3 and np.abs(spec - spec_ref).max() <= 2e-4


def test_float32_input_equals_int16_input(ctx):
    pcm = synth_pcm16(5.0, seed=4)
    a, sa = ctx.spectrogram(pcm)
    b, sb = ctx.spectrogram(pcm16_to_float(pcm))
    np.testing.assert_array_equal(a, b)  # x/32768 is exact, both paths see the same float32 samples
    assert (sa.lo, sa.hi, sa.ref_power) == (sb.lo, sb.hi, sb.ref_power)


@pytest.mark.parametrize("n", [1, 100, 255, 256, 257, 1000, 1023, 1024, 1025, 4095, 12345])
def test_ragged_lengths(ctx, n):
    """Edge frames: zero padding before sample 0 and after the last sample; T = 1 + n // 256."""
    rng = np.random.default_rng(n)
    pcm = rng.integers(-20000, 20000, n).astype(np.int16)
    spec, st = ctx.spectrogram(pcm)
    assert st.n_frames == 1 + n // 256
    db = ctx.read_db(0, spec.shape[0])
    db_ref, _, _, _ = oracle(pcm16_to_float(pcm))
    assert np.abs(db - db_ref).max() <= DB_TOL


def test_closed_form_tone_and_silence(ctx):
    n = 48000
    y = (0.5 * np.sin(2 * np.pi * 40 * 93.75 * np.arange(n) / 48000)).astype(np.float32)
    spec, st = ctx.spectrogram(y)
    np.testing.assert_allclose(st.ref_power, 64.0**2, rtol=1e-4)
    db = ctx.read_db(0, spec.shape[0])
    assert np.all(np.argmax(db[10:-10], axis=1) == 40) and db.max() == 0.0 and db.min() >= -80.0
    # digital silence: every cell sits on the amin floor = the maximum -> 0 dB everywhere, degenerate 0/0 normalisation
    spec0, st0 = ctx.spectrogram(np.zeros(4096, np.int16))
    assert st0.ref_power == 0.0 and st0.lo == 0.0 and st0.hi == 0.0
    assert np.isnan(spec0).all()  # the reference divides 0 by 0 as well (error conventions)
    assert np.all(ctx.read_db(0, spec0.shape[0]) == 0.0)


def test_loud_click_sets_global_reference(ctx):
    """ref=np.max is global over all 257 bins and all frames, including bins above the kept band."""
    rng = np.random.default_rng(0)
    y = (0.001 * rng.standard_normal(48000)).astype(np.float32)
    t = np.arange(2048)
    y[20000 : 20000 + 2048] += (0.9 * np.sin(2 * np.pi * 22000 * t / 48000)).astype(np.float32)  # 22 kHz: outside the 0-16 kHz band
    spec, st = ctx.spectrogram(y)
    db = ctx.read_db(0, spec.shape[0])
    db_ref, _, lo, hi = oracle(y)
    assert db.max() < -20.0  # the loudest kept cell is far below the out-of-band reference
    assert np.abs(db - db_ref).max() <= DB_TOL and abs(st.lo - lo) <= DB_TOL and abs(st.hi - hi) <= DB_TOL


def test_one_hour_properties(ctx):
    """BASELINE config 2 size (1 h, T = 675 001): size-independent properties + an independent exact selection."""
    import torch

    pcm = synth_pcm16(3600.0, seed=20251018)
    ctx.upload_pcm(pcm)
    st = ctx.spectrogram_resident(True)
    T = st.n_frames
    assert T == 675001 and st.rank_lo == 1154252 and st.rank_hi == 115309745
    db = torch.from_numpy(ctx.read_db(0, T)).cuda()
    flat = db.flatten()
    assert float(flat.max()) == 0.0 and float(flat.min()) >= -80.0
    lo = float(torch.kthvalue(flat, st.rank_lo + 1).values)
    hi = float(torch.kthvalue(flat, st.rank_hi + 1).values)
    assert lo == st.lo and hi == st.hi
    spec = torch.from_numpy(ctx.read_spectrogram(0, T)).cuda()
    assert float(spec.min()) == 0.0 and float(spec.max()) == 1.0
    assert int((spec == 0).sum()) >= st.rank_lo + 1 and int((spec == 1).sum()) >= flat.numel() - st.rank_hi
    # spot-check a slice in the middle of the recording against the oracle
    j0 = 300000
    seg = pcm16_to_float(pcm[(j0 - 1) * 256 : (j0 + 300 + 1) * 256])
    S = so.stft_complex64(seg)[:, 1:-1]  # frames fully inside the slice
    P = np.abs(S[:171].T) ** 2
    ref = np.maximum(10 * np.log10(np.maximum(1e-10, P)) - st.db_ref, -80.0)
    assert np.abs(db[j0 : j0 + ref.shape[0]].cpu().numpy() - ref).max() <= DB_TOL


def test_select_shared_prefixes_and_floor_mass(ctx):
    """The radix select's corner paths (select.cu): both ranks inside one digit (one histogram serves both), a rank inside the
    -80 dB floor's digit, and nearly every cell ON the floor (ranks share every prefix; the pass-0 table sees one digit)."""
    rng = np.random.default_rng(3)
    n = 48000 * 20
    t = np.arange(4096)
    burst = (32000 * np.sin(2 * np.pi * 3000 * t / 48000)).astype(np.int16)
    # (a) a short full-scale burst in digital silence: > 99.9 % of the cells sit exactly on the floor -> lo = hi = -80
    pcm = np.zeros(n, np.int16)
    pcm[100000 : 100000 + 256] = burst[:256]
    spec, st = ctx.spectrogram(pcm)
    db = ctx.read_db(0, spec.shape[0])
    flat = np.sort(db.ravel())
    assert st.lo == -80.0 and st.hi == -80.0 and flat[st.rank_lo] == st.lo and flat[st.rank_hi] == st.hi
    assert db.max() == 0.0
    # (b) faint noise under a longer burst: the low rank ON the floor, the high one at -66 dB - both inside [-80, -64) dB = one 11-bit digit
    pcm = np.round(40.0 * rng.standard_normal(n)).astype(np.int16)
    pcm[100000 : 100000 + 4096] += burst
    _, st = check(ctx, pcm)
    assert -80.0 <= st.lo < st.hi < -64.0
    # (c) louder noise: the low rank in the floor's digit, the high one in another
    pcm = np.round(300.0 * rng.standard_normal(n)).astype(np.int16)
    pcm[100000 : 100000 + 4096] += burst
    _, st = check(ctx, pcm)
    assert st.lo < -64.0 < st.hi


def test_both_float64_decompositions(ctx):
    """K1 float64: 16 threads per frame (default, stft_core16.cuh) and 8 threads per frame (stft_core.cuh) are two
    factorisations of the same 256-point FFT: both within the dB gate of the oracle, and within float rounding of each other."""
    pcm = synth_pcm16(30.0, seed=20251018 + 5, calls_per_minute=30.0)
    db_ref, _, lo, hi = oracle(pcm16_to_float(pcm))
    got = {}
    try:
        for threads in (16, 8):
            ctx.set_option("stft_threads", threads)
            spec, st = ctx.spectrogram(pcm)
            got[threads] = ctx.read_db(0, spec.shape[0])
            assert np.abs(got[threads] - db_ref).max() <= DB_TOL
            assert abs(st.lo - lo) <= DB_TOL and abs(st.hi - hi) <= DB_TOL
    finally:
        ctx.set_option("stft_threads", 16)
    assert np.abs(got[16] - got[8]).max() <= 1e-4
    # ragged edges through the 16-thread kernel's guarded loads are covered by test_ragged_lengths (it is the default)
