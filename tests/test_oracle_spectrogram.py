"""Pin the spectrogram oracle: triangulation against torch.stft (float64), closed-form signals, golden vectors."""

import numpy as np
import pytest
import torch

from oracle import spectrogram_oracle as so
from orcai_b200.synth import pcm16_to_float, synth_pcm16

SP = {"sampling_rate": 48000, "nfft": 512, "n_overlap": 256, "freq_range": [0, 16000], "quantiles": [0.01, 0.999], "duration": 4}


def torch_stft(y):
    return torch.stft(
        torch.as_tensor(y, dtype=torch.float64), n_fft=512, hop_length=256, window=torch.hann_window(512, periodic=True, dtype=torch.float64),
        center=True, pad_mode="constant", return_complex=True,
    ).numpy()


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 511, 512, 5000, 48000])
def test_stft_matches_torch_float64(n):
    rng = np.random.default_rng(n)
    y = rng.standard_normal(n).astype(np.float32) * 0.1
    S = so.stft_complex64(y)
    assert S.shape == (257, 1 + n // 256) and S.dtype == np.complex64
    Z = torch_stft(y) if n > 0 else np.zeros((257, 1))
    assert np.abs(S - Z).max() <= 4e-6 * max(1.0, np.abs(Z).max())


def test_pure_tone_at_bin_centre():
    n = 48000
    k = 40  # 3750 Hz
    y = (0.5 * np.sin(2 * np.pi * k * 93.75 * np.arange(n) / 48000)).astype(np.float32)
    S = np.abs(so.stft_complex64(y))
    mid = S[:, 10:-10]
    assert np.all(np.argmax(mid, axis=0) == k)
    # Hann: peak = A * sum(w) / 2 = 0.5 * 256 / 2
    np.testing.assert_allclose(mid[k], 64.0, rtol=1e-4)
    np.testing.assert_allclose(mid[k - 1], 32.0, rtol=1e-3)


def test_dc_and_impulse():
    y = np.full(4096, 0.25, np.float32)
    S = np.abs(so.stft_complex64(y))
    np.testing.assert_allclose(S[0, 2:-2], 0.25 * 256, rtol=1e-6)
    np.testing.assert_allclose(S[1, 2:-2], 0.25 * 128, rtol=1e-6)
    assert S[3:, 2:-2].max() < 1e-5
    imp = np.zeros(4096, np.float32)
    imp[1024] = 1.0  # centre of frame 4: window value 1
    S = np.abs(so.stft_complex64(imp))
    np.testing.assert_allclose(S[:, 4], 1.0, rtol=1e-6)


def test_db_semantics():
    y = pcm16_to_float(synth_pcm16(3.0, seed=3))
    db, f, t = so.calculate_spectrogram(y, SP)
    assert db.dtype == np.float32 and db.max() == 0.0 and db.min() >= -80.0
    assert f.shape == (257,) and f[1] == 93.75 and t[1] == 256 / 48000
    assert so.band_indices(f, [0, 16000]) == (0, 171)
    # silence: amin floor everywhere -> all zeros after the reference shift
    db0, _, _ = so.calculate_spectrogram(np.zeros(2048, np.float32), SP)
    assert np.all(db0 == 0.0)


def test_nearest_rank_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (5, 100, 1923921, 171 * 376):
        a = rng.standard_normal(n).astype(np.float32)
        srt = np.sort(a)
        for q in (0.01, 0.999):
            assert np.percentile(a, 100 * q, method="nearest") == srt[so.nearest_rank(n, q)]
    assert so.nearest_rank(19237671, 0.01) == 192377 and so.nearest_rank(19237671, 0.999) == 19218432


def test_golden_spectrogram(golden_dir):
    g = np.load(golden_dir / "spectrogram_2s.npz")
    y = pcm16_to_float(g["pcm"])
    db, f, _ = so.calculate_spectrogram(y, SP)
    spec, lo, hi = so.preprocess_spectrogram(db, f, SP)
    np.testing.assert_array_equal(db[:171].T, g["db_band"])
    np.testing.assert_array_equal(spec, g["spec"])
    assert lo == g["lo"] and hi == g["hi"]
    assert spec.min() == 0.0 and spec.max() == 1.0 and spec.shape == (376, 171)
