"""Host-side loader semantics (no GPU): the WAV fast path of clap_analyzer.load_audio against what
tasks/analysis.py:170-250 (robust_load_audio_with_fallback -> librosa.load) yields."""
import wave

import numpy as np
import pytest


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


def _write_wav(path, data_bytes, fmt_tag, channels, sr, bits, extensible=False, extra_chunk=True):
    import struct
    block = channels * bits // 8
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, sr, sr * block, block, bits, 22, bits, 0, fmt_tag) + guid_tail
    else:
        fmt = struct.pack("<HHIIHH", fmt_tag, channels, sr, sr * block, block, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if extra_chunk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"      # odd-sized chunk + pad byte
    chunks += b"data" + struct.pack("<I", len(data_bytes)) + data_bytes
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)


def test_wav_decoder_encodings_channels_and_rates(tmp_path):
    """am_wav_decode_mono against what libsndfile's float read + librosa.to_mono yield (x / 2^(bits-1), channel mean in
    float32), for every encoding the decoder accepts, 1-3 channels, odd rates, WAVE_FORMAT_EXTENSIBLE headers."""
    from audiomuse_ai_b200 import _lib, clap_analyzer as ca
    rng = np.random.default_rng(5)
    n = 3001
    for ci, (bits, tag, ch, sr, ext) in enumerate([(16, 1, 1, 48000, False), (16, 1, 2, 44100, False), (24, 1, 2, 48000, True),
                                                   (32, 1, 3, 96000, False), (8, 1, 1, 8000, False), (32, 3, 2, 22050, False),
                                                   (64, 3, 1, 48000, True), (24, 1, 1, 44100, False)]):
        if tag == 3:
            v = rng.uniform(-1, 1, (n, ch)).astype(np.float32 if bits == 32 else np.float64)
            raw, want = v.tobytes(), v.astype(np.float32)
        elif bits == 8:
            q = rng.integers(0, 256, (n, ch)).astype(np.uint8)
            raw, want = q.tobytes(), (q.astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            q = rng.integers(-32768, 32768, (n, ch)).astype("<i2")
            raw, want = q.tobytes(), q.astype(np.float32) / np.float32(32768.0)
        elif bits == 24:
            q = rng.integers(-(1 << 23), 1 << 23, (n, ch)).astype(np.int32)
            raw = b"".join(int(x).to_bytes(4, "little", signed=True)[:3] for x in q.reshape(-1))
            want = q.astype(np.float32) / np.float32(8388608.0)
        else:
            q = rng.integers(-(1 << 31), 1 << 31, (n, ch)).astype(np.int64)
            raw, want = q.astype("<i4").tobytes(), (q.astype(np.float64) / 2147483648.0).astype(np.float32)
        p = str(tmp_path / f"c{ci}.wav")
        _write_wav(p, raw, tag, ch, sr, bits, extensible=ext)
        x, got_sr = ca.decode_wav(p)
        assert got_sr == sr and x.dtype == np.float32 and x.shape == (n,)
        mono = want[:, 0] if ch == 1 else want.mean(axis=1, dtype=np.float32)
        np.testing.assert_array_equal(x, mono.astype(np.float32))
        y, _ = ca.decode_wav(p, max_seconds=1000 / sr)
        np.testing.assert_array_equal(y, x[:1000])
    with pytest.raises(_lib.B200Error):
        ca.decode_wav(str(tmp_path / "missing.wav"))
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFF\x00\x00\x00\x00WAVEjunk")
    with pytest.raises(_lib.B200Error):
        ca.decode_wav(str(bad))
    adpcm = str(tmp_path / "adpcm.wav")
    _write_wav(adpcm, b"\x00" * 64, 2, 1, 48000, 4)          # MS ADPCM: not decoded here -> the reference loader's job
    with pytest.raises(_lib.B200Error) as e:
        ca.decode_wav(adpcm)
    assert "unsupported encoding" in str(e.value)


def test_num_segments_matches_the_reference_rule():
    """am_num_segments == len(segments) of tasks/clap_analyzer.py:510-521 (pinned by segments_golden.npz through the
    oracle)."""
    from audiomuse_ai_b200 import _lib
    from oracle import segments as oseg
    lib = _lib.load()
    for L in (1, 1000, 479_999, 480_000, 480_001, 700_000, 720_000, 720_001, 960_000, 1_199_999, 1_200_000, 1_440_000):
        assert lib.am_num_segments(L) == len(oseg.segment_audio(np.zeros(L, np.float32))), L
