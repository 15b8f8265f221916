"""Seeded synthetic PCM streams (the reference ships no audio).

Recipe from SURVEY.md section 8(d): per channel c, 8 harmonics of
f0 = 110*(1+0.5c)*2^(u/12), u~U(-6,6); amplitude 0.25/(k+1); random start phases;
1 % vibrato at 0.7 Hz; 1.3 Hz tremolo in 0.6..1.0; N(0, 0.01^2) noise; clipped to
+-1 and quantised to int16 (x*32767).  Both implementations are fed int16/32768
as float32, which is exactly what the reference WAV reader produces
(/root/reference/main/wavfile.cc:736-749).
"""
from __future__ import annotations

import numpy as np


def synth_int16(seed: int, sr: int, secs: float, ch: int) -> np.ndarray:
    """Return int16 PCM of shape [ch, n]."""
    rng = np.random.default_rng(seed)
    n = int(round(sr * secs))
    t = np.arange(n, dtype=np.float64) / sr
    out = np.empty((ch, n), dtype=np.int16)
    for c in range(ch):
        u = rng.uniform(-6.0, 6.0)
        f0 = 110.0 * (1.0 + 0.5 * c) * 2.0 ** (u / 12.0)
        ph0 = rng.uniform(0.0, 2.0 * np.pi, size=8)
        vib_ph = rng.uniform(0.0, 2.0 * np.pi)
        trem_ph = rng.uniform(0.0, 2.0 * np.pi)
        # 1 % vibrato at 0.7 Hz: instantaneous frequency f0*(1+0.01 sin(2 pi 0.7 t))
        inst = 2.0 * np.pi * f0 * (t - 0.01 / (2.0 * np.pi * 0.7) * np.cos(2.0 * np.pi * 0.7 * t + vib_ph))
        x = np.zeros(n, dtype=np.float64)
        for k in range(8):
            x += 0.25 / (k + 1) * np.sin((k + 1) * inst + ph0[k])
        trem = 0.8 + 0.2 * np.sin(2.0 * np.pi * 1.3 * t + trem_ph)
        x = x * trem + rng.normal(0.0, 0.01, size=n)
        x = np.clip(x, -1.0, 1.0)
        out[c] = np.round(x * 32767.0).astype(np.int16)
    return out


def int16_to_float(pcm: np.ndarray) -> np.ndarray:
    """int16 -> float32 the way the reference WAV reader does (double multiply by 1/32768)."""
    return (pcm.astype(np.float64) * (1.0 / 32768.0)).astype(np.float32)


def synth(seed: int, sr: int, secs: float, ch: int) -> np.ndarray:
    """float32 [ch, n] stream as both implementations see it."""
    return int16_to_float(synth_int16(seed, sr, secs, ch))
