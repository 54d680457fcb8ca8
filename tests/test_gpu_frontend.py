"""Front end on the GPU: the polyphase resampler against scipy.signal.resample_poly (librosa's
soxr_hq is not installable: parity is pinned against scipy only), the device clip / quantise / windowing kernel
against the host function (itself golden-pinned to the reference), and file-based bulk analysis against the
per-file call."""
import ctypes as C
import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sr_in", [44100, 22050, 96000, 16000, 32000, 11025])
def test_resample_matches_scipy_resample_poly(sr_in):
    from math import gcd
    from scipy.signal import resample_poly
    from audiomuse_ai_b200 import clap_analyzer as ca
    rng = np.random.default_rng(sr_in)
    t = np.arange(3 * sr_in) / sr_in
    x = (0.4 * np.sin(2 * np.pi * 440.0 * t) + 0.2 * np.sin(2 * np.pi * 0.31 * sr_in * t) + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
    y = ca.resample(x, sr_in, 48000)
    g = gcd(sr_in, 48000)
    want = resample_poly(x.astype(np.float64), 48000 // g, sr_in // g)
    assert y.dtype == np.float32 and y.shape == want.shape
    err = np.abs(y - want).max()
    print(f"[resample] {sr_in} -> 48000 (up/down {48000 // g}/{sr_in // g}): max |err| vs scipy = {err:.2e}")
    assert err <= 5e-7
    np.testing.assert_array_equal(ca.resample(x, 48000, 48000), x)


@pytest.mark.parametrize("L", [1000, 479_999, 480_000, 480_001, 720_000, 1_199_999, 1_440_000])
def test_device_segmentation_equals_host_function(L):
    import torch
    from audiomuse_ai_b200 import _lib, clap_analyzer as ca
    lib = _lib.load()
    rng = np.random.default_rng(L)
    x = (rng.standard_normal(L) * 0.5).astype(np.float32)
    x[::997] *= 4.0                                  # beyond +-1: exercises the clip
    want = ca.pcm_to_segments(x)
    xd = torch.from_numpy(x).cuda()
    n = C.c_int(0)
    _lib.check(lib.am_audio_to_segments_dev(C.c_void_p(xd.data_ptr()), L, None, 0, C.byref(n), None))
    assert n.value == len(want) == lib.am_num_segments(L)
    seg = torch.empty((n.value, 480000), dtype=torch.int16, device="cuda")
    _lib.check(lib.am_audio_to_segments_dev(C.c_void_p(xd.data_ptr()), L, C.c_void_p(seg.data_ptr()), n.value, C.byref(n),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(seg.cpu().numpy(), want)


def _wav(path, pcm, sr, ch=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(ch); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.repeat(pcm, ch).astype("<i2").tobytes() if ch > 1 else pcm.astype("<i2").tobytes())


def test_load_audio_resamples_44k1_and_bulk_file_analysis(tmp_path):
    from math import gcd
    from scipy.signal import resample_poly
    from audiomuse_ai_b200 import clap_analyzer as ca, corpus, weights
    sess = ca.B200Session.from_state_dict(weights.random_state_dict(0))
    ca.set_clap_audio_session(sess)
    try:
        paths = []
        for i in range(5):
            pcm48 = corpus.synth_track(20 + i, length=480000 if i != 3 else 700000)
            p = tmp_path / f"t{i}.wav"
            if i % 2 == 0:
                _wav(p, pcm48, 48000, ch=1 + (i == 2))
            else:  # a 44.1 kHz file: what most real libraries hold
                n441 = int(len(pcm48) * 44100 / 48000)
                _wav(p, pcm48[:n441], 44100)
            paths.append(str(p))
        x, sr = ca.load_audio(paths[1])
        raw, sr0 = ca.decode_wav(paths[1])
        assert sr == 48000 and sr0 == 44100
        want = resample_poly(raw.astype(np.float64), 160, 147)
        assert x.shape == want.shape and np.abs(x - want).max() <= 5e-7
        x0, _ = ca.load_audio(paths[0])
        np.testing.assert_array_equal(x0, corpus.pcm16_to_float(corpus.synth_track(20)))
        bad = tmp_path / "broken.wav"
        bad.write_bytes(b"not a wav file")
        all_paths = paths[:2] + [str(bad)] + paths[2:]
        stats = {}
        bulk = list(ca.analyze_audio_files(all_paths, batch_tracks=2, workers=3, stats=stats))
        assert len(bulk) == 6 and bulk[2] == (None, 0, 0)
        for p, (emb, dur, nseg) in zip(all_paths, bulk):
            one = ca.analyze_audio_file(p)
            if one[0] is None:
                assert emb is None
                continue
            np.testing.assert_allclose(emb, one[0], atol=2e-6)
            assert dur == one[1] and nseg == one[2]
        assert bulk[4][2] == 2 and stats["wall_seconds"] > 0 and stats["decode_thread_seconds"] > 0   # the 700 000-sample file: 2 windows
    finally:
        ca.set_clap_audio_session(None)
