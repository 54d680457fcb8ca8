"""End-to-end analysis STARTING FROM WAV FILES ON DISK: N synthetic 10 s tracks are written as PCM16
WAV files (a share of them at 44.1 kHz), then clap_analyzer.analyze_audio_files decodes them on a thread pool
(am_wav_decode_mono; 44.1 kHz files are resampled on the GPU) while B200Session.embed_tracks_stream embeds the previous
batches.  Prints one JSON line: tracks/s from files, the decode share, and the same tracks from pinned memory.
    python tools/e2e_from_wav.py [--n 1024] [--share-441 0.25] [--workers 32]"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import wave

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiomuse_ai_b200 import clap_analyzer as ca, corpus, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--share-441", type=float, default=0.25)
ap.add_argument("--workers", type=int, default=0)
a = ap.parse_args()

root = tempfile.mkdtemp(prefix="am_wav_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    base = corpus.synth_pcm_batch(64, start=500)
    paths, paths48 = [], []
    n441 = 0
    for i in range(a.n):
        pcm = base[i % 64]
        p = os.path.join(root, f"t{i:05d}.wav")
        with wave.open(p, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2)
            if (i * a.share_441) % 1.0 + a.share_441 >= 1.0 and a.share_441 > 0:
                w.setframerate(44100); w.writeframes(pcm[:441000].tobytes()); n441 += 1
            else:
                w.setframerate(48000); w.writeframes(pcm.tobytes()); paths48.append(p)
        paths.append(p)
    sess = ca.B200Session.from_state_dict(weights.random_state_dict(0))
    ca.set_clap_audio_session(sess)
    workers = a.workers or min(32, os.cpu_count() or 4)
    list(ca.analyze_audio_files(paths[:256], workers=workers))          # warm-up (page cache, workspaces)
    stats = {}
    t0 = time.perf_counter()
    res = list(ca.analyze_audio_files(paths, workers=workers, stats=stats))
    dt = time.perf_counter() - t0
    assert all(r[0] is not None for r in res)
    # the files that are already at 48 kHz alone (one library call per file, no GPU resample round trip)
    n48 = (len(paths48) // 256) * 256
    st48 = {}
    t2 = time.perf_counter()
    res48 = list(ca.analyze_audio_files(paths48[:n48], workers=workers, stats=st48))
    dt48 = time.perf_counter() - t2
    # the same number of tracks from memory (no decode): the ceiling the file path is measured against
    pcm = np.ascontiguousarray(np.stack([base[i % 64] for i in range(256)]))
    offs = np.arange(257, dtype=np.int32)
    list(sess.embed_tracks_stream([(pcm, offs)] * 2))
    t1 = time.perf_counter()
    for _ in sess.embed_tracks_stream((pcm, offs) for _ in range(a.n // 256)):
        pass
    dt_mem = time.perf_counter() - t1
    print(json.dumps({"tracks": a.n, "files_at_44100_hz": n441, "workers": workers, "host_cores": os.cpu_count(),
                      "tracks_per_s_from_wav_files": a.n / dt, "wall_s": dt,
                      "decode_thread_seconds": stats["decode_thread_seconds"],
                      "decode_share_of_wall_if_serial": stats["decode_thread_seconds"] / dt,
                      "decode_ms_per_track_per_thread": 1e3 * stats["decode_thread_seconds"] / a.n,
                      "tracks_per_s_from_48khz_files_only": n48 / dt48 if n48 else None,
                      "decode_ms_per_48khz_track_per_thread": 1e3 * st48["decode_thread_seconds"] / n48 if n48 else None,
                      "tracks_per_s_from_pageable_memory": (a.n // 256) * 256 / dt_mem,
                      "note": "decode = RIFF parse + PCM16 -> float32 (+ GPU polyphase resample for 44.1 kHz files) + the reference's "
                              "clip / int16 round trip / windowing, on a thread pool; files on tmpfs"}))
finally:
    shutil.rmtree(root, ignore_errors=True)
