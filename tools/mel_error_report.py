"""Prints, per corpus track, the mel kernel's error against the numpy oracle WITHOUT any masking: max |dB| error over
all bins, over bins within 60 / 80 / 100 dB of the frame peak, and max |power error| / frame peak.  The bounds asserted
in tests/test_gpu_mel.py come from this report (run on the GPU box: python tools/mel_error_report.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiomuse_ai_b200 import clap_analyzer as ca, corpus  # noqa: E402
from oracle import mel as omel, segments as oseg  # noqa: E402

wins = []
for i in range(8):
    x, _ = oseg.int16_round_trip(corpus.pcm16_to_float(corpus.synth_track(i)))
    wins.append(oseg.segment_audio(x)[-1 if i == 5 else 0])
wins = np.stack(wins)
got = ca.compute_mel_spectrogram_batch(wins)
for i, w in enumerate(wins):
    want_db = omel.compute_mel_spectrogram(w)[0, 0]
    want_pow = omel.mel_power(w).astype(np.float64)
    g = got[i, 0]
    err = np.abs(g - want_db)
    peak = want_pow.max(axis=0, keepdims=True) + 1e-300
    rel = want_pow / peak
    gp = np.power(10.0, g.astype(np.float64) / 10.0)
    perr = np.abs(gp - np.maximum(want_pow, 1e-10)) / np.maximum(peak, 1e-10)
    row = [f"track {i}: max dB err all bins {err.max():.3e}"]
    for lim in (1e-6, 1e-8, 1e-10):
        m = rel >= lim
        row.append(f">= {lim:g} peak: {err[m].max() if m.any() else 0:.3e}")
    row.append(f"max |dP|/peak {perr.max():.3e}")
    print("; ".join(row), flush=True)
