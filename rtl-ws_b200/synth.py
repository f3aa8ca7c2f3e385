"""Seeded synthetic 8-bit interleaved IQ captures (S1, S2, S3).

All generators return ``uint8`` arrays of shape ``[n, 2]`` (re, im) in the RTL-SDR wire
format the reference's ``cmplx_u8`` describes (common_sp.h:7-11): offset binary, 128 = 0.
Pure numpy; shared by the tests, ``bench.py`` and ``tests/golden/make_golden.py`` so that
the CPU checker and the GPU path always see the same bytes.
"""
from __future__ import annotations

import numpy as np

FS_DEFAULT = 2_048_000          # rtl_sensor.c:12


def _quantise(x: np.ndarray) -> np.ndarray:
    out = np.empty((len(x), 2), dtype=np.uint8)
    out[:, 0] = np.clip(np.rint(x.real), 0, 255).astype(np.uint8)
    out[:, 1] = np.clip(np.rint(x.imag), 0, 255).astype(np.uint8)
    return out


def s1_noise(n: int, seed: int = 0) -> np.ndarray:
    """S1: iid uniform bytes -- worst-case entropy, for the bit-exact integer stages."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, 2), dtype=np.uint8)


def s2_tones(n: int, N: int = 1024, fs: float = FS_DEFAULT, seed: int = 1,
             amplitude: float = 100.0, sigma: float = 4.0) -> np.ndarray:
    """S2: three complex tones (0, +fs/8, -fs/4 + fs/(3N)) sharing `amplitude`, plus noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    freqs = [0.0, fs / 8.0, -fs / 4.0 + fs / (3.0 * N)]
    x = np.zeros(n, dtype=np.complex128)
    for f in freqs:
        x += (amplitude / len(freqs)) * np.exp(2j * np.pi * (f / fs) * t)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return _quantise(x + (127.5 + 127.5j))


def s3_fm(n: int, fs: float = FS_DEFAULT, seed: int = 2, deviation: float = 25_000.0,
          carrier: float = 0.0, amplitude: float = 100.0, sigma: float = 2.0,
          tones=(1_000.0, 5_000.0)) -> np.ndarray:
    """S3: FM broadcast-like signal: message = equal mix of `tones`, peak deviation `deviation` Hz."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    msg = np.zeros(n, dtype=np.float64)
    for f in tones:
        msg += np.sin(2 * np.pi * f * t)
    msg /= len(tones)
    phase = 2 * np.pi * np.cumsum(deviation * msg + carrier) / fs
    x = amplitude * np.exp(1j * phase)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return _quantise(x + (127.5 + 127.5j))


def tone(n: int, k: float, N: int, amplitude: float = 127.0) -> np.ndarray:
    """A single complex exponential landing on (fractional) FFT bin k of an N-point frame."""
    t = np.arange(n, dtype=np.float64)
    x = amplitude * np.exp(2j * np.pi * k * t / N)
    return _quantise(x + (128.0 + 128.0j))


def hann(N: int) -> np.ndarray:
    """Periodic Hann window (float64); the reference itself is rectangular."""
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N, dtype=np.float64) / N)
