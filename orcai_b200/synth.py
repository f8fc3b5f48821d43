"""Deterministic synthetic recordings for tests and benchmarks.

48 kHz mono PCM16: Gaussian noise floor (sigma = 0.02 full scale) plus sparse "calls":
linear/quadratic chirps between 1 and 12 kHz lasting 0.2-2 s with amplitude U(0.1, 0.5),
about six per minute, plus short click trains.  File k uses seed 20251018 + k.
"""

from __future__ import annotations

import numpy as np

BASE_SEED = 20251018


def synth_pcm16(seconds: float, sr: int = 48000, seed: int = BASE_SEED, calls_per_minute: float = 6.0) -> np.ndarray:
    """Return int16 samples of a synthetic recording."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    y = np.empty(n, dtype=np.float32)
    blk = 1 << 22
    for s in range(0, n, blk):  # chunked so 1-h recordings do not need float64 temporaries
        e = min(n, s + blk)
        y[s:e] = rng.standard_normal(e - s, dtype=np.float32) * np.float32(0.02)
    n_calls = max(1, int(round(calls_per_minute * seconds / 60.0)))
    for _ in range(n_calls):
        dur = rng.uniform(0.2, 2.0)
        m = int(dur * sr)
        if m >= n:
            m = n // 2
        start = int(rng.integers(0, max(1, n - m)))
        f0, f1 = rng.uniform(1000.0, 12000.0, 2)
        amp = rng.uniform(0.1, 0.5)
        t = np.arange(m, dtype=np.float64) / sr
        if rng.random() < 0.5:
            phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / max(dur, 1e-3) * t * t)
        else:
            phase = 2 * np.pi * (f0 * t + (f1 - f0) / (3 * max(dur, 1e-3) ** 2) * t**3)
        env = np.sin(np.pi * np.arange(m) / m) ** 2
        y[start : start + m] += (amp * env * np.sin(phase)).astype(np.float32)
    n_trains = max(1, int(round(seconds / 30.0)))
    for _ in range(n_trains):
        k = int(rng.integers(5, 30))
        start = int(rng.integers(0, max(1, n - sr)))
        gap = int(rng.uniform(0.002, 0.02) * sr)
        for j in range(k):
            p = start + j * gap
            if p + 48 < n:
                y[p : p + 48] += (rng.uniform(0.2, 0.6) * np.hanning(48)).astype(np.float32) * (1 if j % 2 else -1)
    np.clip(y, -1.0, 32767.0 / 32768.0, out=y)
    return np.round(y * 32768.0).clip(-32768, 32767).astype(np.int16)


def pcm16_to_float(pcm: np.ndarray) -> np.ndarray:
    """What librosa.load / libsndfile hand back for PCM16: x / 32768 in float32 (exact)."""
    return pcm.astype(np.float32) / np.float32(32768.0)
