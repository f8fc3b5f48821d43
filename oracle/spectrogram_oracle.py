"""numpy restatement of the reference's spectrogram stage (oracle; see oracle/__init__.py).

Follows ``src/orcAI/spectrogram.py:15-87`` of the reference, with librosa 0.11.0 /
numpy semantics spelled out:

* ``librosa.load``      -> float32 samples (PCM16 / 32768), no resampling at 48 kHz   (spectrogram.py:23-31)
* ``librosa.stft``      -> centred (zero pad n_fft//2), periodic Hann in float64,
                           ``numpy.fft.rfft`` in float64, stored as complex64         (spectrogram.py:34-39)
* ``amplitude_to_db``   -> float32: 10*log10(max(1e-10, |S|^2)) - 10*log10(max(1e-10, max|S|^2)),
                           floored at (max - 80 dB)                                   (spectrogram.py:51-53)
* ``preprocess``        -> crop to the band, nearest-rank percentiles, clip, min-max  (spectrogram.py:58-87)
"""

from __future__ import annotations

import numpy as np

AMIN_POWER = np.float32(1e-10)  # librosa amplitude_to_db: amin=1e-5 -> power floor amin**2
TOP_DB = np.float32(80.0)


def hann_periodic(n_fft: int) -> np.ndarray:
    """scipy.signal.get_window("hann", n_fft, fftbins=True) in float64."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)


def num_frames(n_samples: int, hop: int) -> int:
    """librosa.stft(center=True): 1 + n // hop frames."""
    return 1 + n_samples // hop


def stft_complex64(y: np.ndarray, n_fft: int = 512, hop: int = 256, block_frames: int = 8192) -> np.ndarray:
    """Centred, zero-padded, Hann-windowed STFT; returns complex64 (1 + n_fft//2, T).

    The product window*frame and the rFFT are float64 (the window is float64 and
    numpy promotes), the result is rounded into a complex64 array - exactly what
    librosa does when handed float32 audio.
    """
    y = np.asarray(y, dtype=np.float32)
    n = y.shape[0]
    T = num_frames(n, hop)
    pad = n_fft // 2
    yp = np.zeros(n + 2 * pad, dtype=np.float32)
    yp[pad : pad + n] = y
    win = hann_periodic(n_fft)[:, None]  # (n_fft, 1) float64
    out = np.empty((1 + n_fft // 2, T), dtype=np.complex64)
    frames = np.lib.stride_tricks.sliding_window_view(yp, n_fft)[::hop]  # (T', n_fft) view
    assert frames.shape[0] >= T
    for s in range(0, T, block_frames):
        e = min(T, s + block_frames)
        blk = frames[s:e].T  # (n_fft, b) float32
        out[:, s:e] = np.fft.rfft(win * blk, axis=0)  # float64 math, cast to complex64
    return out


def amplitude_to_db_refmax(S: np.ndarray) -> np.ndarray:
    """librosa.amplitude_to_db(np.abs(S), ref=np.max) with amin=1e-5, top_db=80; float32."""
    magnitude = np.abs(S)  # float32 for complex64 input
    assert magnitude.dtype == np.float32
    ref_value = np.max(magnitude)  # float32 scalar, global over all bins and frames
    power = np.square(magnitude, out=magnitude)
    log_spec = np.float32(10.0) * np.log10(np.maximum(AMIN_POWER, power))
    log_spec -= np.float32(10.0) * np.log10(np.maximum(AMIN_POWER, ref_value * ref_value))
    log_spec = np.maximum(log_spec, log_spec.max() - TOP_DB)
    assert log_spec.dtype == np.float32
    return log_spec


def fft_frequencies(sr: int, n_fft: int) -> np.ndarray:
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def frames_to_time(T: int, sr: int, hop: int) -> np.ndarray:
    return (np.arange(T) * hop) / float(sr)


def band_indices(frequencies: np.ndarray, freq_range) -> tuple[int, int]:
    lo = int(np.argwhere(frequencies <= freq_range[0])[0][0])
    hi = int(np.argwhere(frequencies >= freq_range[1])[0][0])
    return lo, hi


def calculate_spectrogram(y: np.ndarray, spectrogram_parameter: dict):
    """(dB spectrogram (257, T) float32, frequencies (257,), times (T,))."""
    sr = spectrogram_parameter["sampling_rate"]
    n_fft = spectrogram_parameter["nfft"]
    hop = spectrogram_parameter["n_overlap"]  # the reference uses n_overlap as the hop length
    S = stft_complex64(y, n_fft, hop)
    db = amplitude_to_db_refmax(S)
    return db, fft_frequencies(sr, n_fft), frames_to_time(S.shape[1], sr, hop)


def preprocess_spectrogram(db: np.ndarray, frequencies: np.ndarray, spectrogram_parameter: dict):
    """Crop, percentile-clip, normalise, transpose -> (T, n_band) float32, plus (lo, hi)."""
    i0, i1 = band_indices(frequencies, spectrogram_parameter["freq_range"])
    band = db[i0:i1, :]
    q = spectrogram_parameter["quantiles"]
    lo = np.percentile(band, 100 * q[0], method="nearest")
    hi = np.percentile(band, 100 * q[1], method="nearest")
    clipped = np.clip(band, lo, hi)
    mn = np.min(clipped)
    mx = np.max(clipped)
    out = (clipped - mn) / (mx - mn)
    return out.T, np.float32(lo), np.float32(hi)


def make_spectrogram(y: np.ndarray, spectrogram_parameter: dict):
    """Full stage: float32 audio -> (normalised (T, n_band) float32, frequencies, times)."""
    db, freqs, times = calculate_spectrogram(y, spectrogram_parameter)
    spec, _, _ = preprocess_spectrogram(db, freqs, spectrogram_parameter)
    return spec, freqs, times


def nearest_rank(n: int, quantile: float) -> int:
    """Index into the sorted flattened array that np.percentile(..., 100*quantile, 'nearest') returns."""
    q = np.true_divide(100 * quantile, 100)
    return int(np.around((n - 1) * q))
