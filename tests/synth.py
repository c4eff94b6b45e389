"""Deterministic synthetic PCM shared by tests and bench.py: a few enveloped partials
with a slow chirp plus a noise floor, quantised to int16 and scaled by 1/32768 so every sample is
an exact f32 -- the same representation audio.rs:16-19 produces for 16-bit WAV files."""
import numpy as np


CELL = 1 << 20  # the tonal part repeats every CELL samples; the noise floor never repeats


def base_clip_i16(n: int, sr: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    freqs = np.geomspace(80.0, 0.45 * sr, 6)
    amps = rng.uniform(0.3, 1.0, 6)
    phases = rng.uniform(0, 2 * np.pi, 6)
    env_rate = rng.uniform(0.05, 0.4, 6)
    m = min(n, CELL)
    t = np.arange(m, dtype=np.float64) / sr
    cell = np.zeros(m, np.float64)
    for k in range(6):
        f = freqs[k]
        ph = 2 * np.pi * f * t + phases[k]
        if k in (2, 4):  # slow chirp on two partials
            ph = ph + 0.4 * f * np.sin(2 * np.pi * 0.05 * t)
        env = 0.5 + 0.5 * np.sin(2 * np.pi * env_rate[k] * t + phases[k])
        cell += amps[k] * np.sin(ph) * env
    cell = (0.20 * cell / 3.0).astype(np.float32)
    out = np.empty(n, np.float32)
    for s in range(0, n, CELL):
        e = min(n, s + CELL)
        out[s:e] = cell[: e - s] + np.float32(0.02) * rng.standard_normal(e - s, dtype=np.float32)
    return np.clip(np.rint(out * np.float32(32768.0)), -32768, 32767).astype(np.int16)


def base_clip(n: int, sr: int, seed: int) -> np.ndarray:
    return base_clip_i16(n, sr, seed).astype(np.float32) / np.float32(32768.0)


def track_gain_shift(t: int, n: int):
    """Track t of a multi-track batch: gain_t * base[(i + shift_t) mod n] (exact in f32)."""
    gain = np.float32((128 + (37 * t) % 128) / 256.0)
    shift = (7919 * t) % n
    return gain, shift


def derive_track(base: np.ndarray, t: int) -> np.ndarray:
    gain, shift = track_gain_shift(t, base.size)
    return (np.roll(base, -shift) * gain).astype(np.float32)
