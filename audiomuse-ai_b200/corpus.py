"""Synthetic corpora for parity tests and bench.py.

No network and no audio files are available, so every workload is generated from
seeds: PCM16 tracks for the analysis path (config 2), unit-norm embedding libraries
for k-NN (config 3) and clustered libraries for k-means (config 4).
"""
from __future__ import annotations

import numpy as np

SR = 48000
SEG = 480000


def synth_track(i: int, length: int = SEG) -> np.ndarray:
    """Track ``i`` as int16 PCM.  Tracks 0-5 are the edge cases of SURVEY 8(d) config 2:
    0 silence, 1 full-scale 1 kHz sine, 2 white noise N(0, 0.1), 3 over-range (clips),
    4 one sample short of 10 s, 5 a 25 s track (5 windows incl. the tail window)."""
    rng = np.random.default_rng(1000 + i)
    if i == 4:
        length = SEG - 1
    if i == 5:
        length = 1_200_000
    t = np.arange(length, dtype=np.float64) / SR
    if i == 0:
        x = np.zeros(length)
    elif i == 1:
        x = np.sin(2 * np.pi * 1000.0 * t)
    elif i == 2:
        x = rng.standard_normal(length) * 0.1
    else:
        freqs = np.exp(rng.uniform(np.log(40.0), np.log(16000.0), 8))
        amps = np.exp(rng.uniform(np.log(0.01), np.log(0.3), 8))
        phases = rng.uniform(0, 2 * np.pi, 8)
        x = np.zeros(length)
        for f, a, p in zip(freqs, amps, phases):
            x += a * np.sin(2 * np.pi * f * t + p)
        # pink-ish noise at -30 dBFS: white noise shaped by 1/sqrt(f)
        w = rng.standard_normal(length)
        spec = np.fft.rfft(w)
        fr = np.fft.rfftfreq(length, 1.0 / SR)
        fr[0] = fr[1]
        spec /= np.sqrt(fr)
        pink = np.fft.irfft(spec, n=length)
        pink *= 10 ** (-30 / 20) / (np.sqrt(np.mean(pink**2)) + 1e-12)
        x = (x + pink) * rng.uniform(0.1, 1.0)
        if i == 3:
            x *= 4.0  # exceeds +-1 before quantisation: exercises the clip
    return np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)


def synth_pcm_batch(n_tracks: int, start: int = 6, seed: int = 7) -> np.ndarray:
    """Fast bulk generator for bench workloads: int16[n_tracks, SEG] of 10 s tracks
    (sinusoid mixtures + noise).  Cheaper than synth_track (no per-track FFT shaping)."""
    rng = np.random.default_rng(seed + start)
    out = np.empty((n_tracks, SEG), dtype=np.int16)
    t = np.arange(SEG, dtype=np.float32) / np.float32(SR)
    for i in range(n_tracks):
        freqs = np.exp(rng.uniform(np.log(40.0), np.log(16000.0), 6)).astype(np.float32)
        amps = np.exp(rng.uniform(np.log(0.01), np.log(0.2), 6)).astype(np.float32)
        x = rng.standard_normal(SEG, dtype=np.float32) * np.float32(0.02)
        for f, a in zip(freqs, amps):
            x += a * np.sin(np.float32(2 * np.pi) * f * t)
        out[i] = np.clip(x * 32767.0, -32768, 32767).astype(np.int16)
    return out


def pcm16_to_float(pcm16: np.ndarray) -> np.ndarray:
    """What librosa.load yields for a PCM16 WAV: x / 32768 as float32 (analysis.py:181)."""
    return (pcm16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def knn_library(n: int = 100_000, d: int = 512, seed: int = 1234) -> np.ndarray:
    x = np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def knn_queries(x: np.ndarray, n_near: int = 10_000, n_random: int = 1_000, seed: int = 4321) -> np.ndarray:
    rng = np.random.default_rng(seed)
    d = x.shape[1]
    idx = rng.integers(0, x.shape[0], n_near)
    noise = rng.standard_normal((n_near, d), dtype=np.float32)
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)  # unit noise, scaled 0.3
    near = x[idx] + 0.3 * noise
    rnd = rng.standard_normal((n_random, d), dtype=np.float32)
    q = np.concatenate([near, rnd], 0)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def kmeans_library(n: int = 1_000_000, d: int = 512, k: int = 128, seed: int = 7):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((k, d), dtype=np.float32)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    lab = rng.integers(0, k, n)
    noise = rng.standard_normal((n, d), dtype=np.float32)
    noise *= np.float32(0.5 / np.sqrt(d))
    x = centers[lab] + noise
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32), lab.astype(np.int32), centers
