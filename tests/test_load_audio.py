"""Host-side loader semantics (no GPU): the WAV fast path of clap_analyzer.load_audio against what
tasks/analysis.py:170-250 (robust_load_audio_with_fallback -> librosa.load) yields."""
import wave

import numpy as np


def test_long_wav_is_capped_like_librosa_duration(tmp_path):
    """robust_load_audio_with_fallback passes duration=AUDIO_LOAD_TIMEOUT to librosa.load (analysis.py:181): audio
    beyond that many seconds is never read.  (the WAV fast path read the whole file.)"""
    from audiomuse_ai_b200 import clap_analyzer as ca
    p = tmp_path / "long.wav"
    pcm = (np.random.default_rng(0).standard_normal(48000 * 5) * 3000).astype(np.int16)
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000); w.writeframes(np.repeat(pcm, 2).tobytes())
    old = getattr(ca.config, "AUDIO_LOAD_TIMEOUT", 600)
    try:
        ca.config.AUDIO_LOAD_TIMEOUT = 2
        x, sr = ca.load_audio(str(p))
        assert sr == 48000 and x.shape == (96000,)
        np.testing.assert_array_equal(x, pcm[:96000].astype(np.float32) / np.float32(32768.0))   # stereo mean of equal channels
        ca.config.AUDIO_LOAD_TIMEOUT = 600
        assert ca.load_audio(str(p))[0].shape == (240000,)
    finally:
        ca.config.AUDIO_LOAD_TIMEOUT = old
