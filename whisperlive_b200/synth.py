"""Seeded synthetic 16 kHz PCM used by tests and bench.py.

speech-like = a few harmonics of a 90-250 Hz f0 with a 3-8 Hz syllable envelope
plus -30 dBFS white noise, peak 0.5.  No reference code involved.
"""
from __future__ import annotations

import numpy as np

SR = 16000


def speech_like(seconds: float, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(round(seconds * SR))
    t = np.arange(n, dtype=np.float64) / SR
    f0 = rng.uniform(90.0, 250.0)
    vib = 1.0 + 0.03 * np.sin(2 * np.pi * rng.uniform(4.0, 7.0) * t)
    phase = 2 * np.pi * np.cumsum(f0 * vib) / SR
    sig = np.zeros(n)
    for h in range(1, int(rng.integers(3, 6)) + 1):
        sig += (1.0 / h) * np.sin(h * phase + rng.uniform(0, 2 * np.pi))
    env = 0.5 * (1.0 + np.sin(2 * np.pi * rng.uniform(3.0, 8.0) * t + rng.uniform(0, 2 * np.pi)))
    sig = sig * env ** 2
    sig += 10 ** (-30 / 20) * rng.standard_normal(n)
    sig *= 0.5 / max(1e-9, np.abs(sig).max())
    return sig.astype(np.float32)


def white_noise(seconds: float, seed: int = 1234, sigma: float = 0.1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (sigma * rng.standard_normal(int(round(seconds * SR)))).astype(np.float32)


def silence(seconds: float) -> np.ndarray:
    return np.zeros(int(round(seconds * SR)), dtype=np.float32)


def chunk_durations(n_streams: int, lo: float = 5.0, hi: float = 30.0, seed: int = 1234) -> list:
    """VAD-gated chunk lengths U[lo,hi] s, one draw per stream (seed + stream)."""
    return [float(np.random.default_rng(seed + i).uniform(lo, hi)) for i in range(n_streams)]
