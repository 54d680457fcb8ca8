"""K1 parity: CUDA log-mel (through the C ABI) vs the numpy oracle on the synthetic corpus."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mel as omel
from oracle import segments as oseg


def _windows():
    from audiomuse_ai_b200 import corpus
    wins = []
    for i in range(8):
        pcm = corpus.synth_track(i)
        x, _ = oseg.int16_round_trip(corpus.pcm16_to_float(pcm))
        wins.append(oseg.segment_audio(x)[-1 if i == 5 else 0])
    return np.stack(wins)


def _check(got_db, want_db, want_pow, tag=""):
    """Bounds set from tools/mel_error_report.py on the B200 (profiles/r02_mel_error_report.txt): the kernel's worst
    error over ALL bins of the corpus is 7.6e-5 dB except under the full-scale sine, whose side bins sit 80-100+ dB
    below the frame peak (2.7e-4 dB at >= 1e-8 of the peak, 4.6e-3 dB at >= 1e-10).  Asserted:
      * <= 1e-3 dB on every bin within 80 dB of its frame's strongest bin;
      * everywhere (no mask): |P_got - P_want| <= 1e-3 * P_want + 1e-9 * frame peak -- a bin 80 dB down may be off
        by at most 10 %, one 60 dB down by 0.1 % (the round-1 bound allowed 2e-6 * peak: 100 % at -57 dB)."""
    assert got_db.shape == want_db.shape
    peak = want_pow.max(axis=0, keepdims=True).astype(np.float64)
    err_db = np.abs(got_db - want_db)
    strong = want_pow >= 1e-8 * np.maximum(peak, 1e-300)
    print(f"[mel parity] {tag} max |dB| error: all bins {err_db.max():.2e}, bins >= 1e-8 peak {err_db[strong].max() if strong.any() else 0:.2e}")
    if strong.any():
        assert err_db[strong].max() <= 1e-3, f"max dB error on bins within 80 dB of the peak {err_db[strong].max()}"
    got_pow = np.power(10.0, got_db.astype(np.float64) / 10.0)
    floor = np.maximum(want_pow.astype(np.float64), 1e-10)
    abs_err = np.abs(got_pow - floor)
    assert (abs_err <= 1e-3 * floor + 1e-9 * peak).all(), "power error beyond fp32 FFT accuracy"


def test_mel_matches_oracle_on_corpus():
    from audiomuse_ai_b200 import clap_analyzer as ca
    wins = _windows()
    got = ca.compute_mel_spectrogram_batch(wins)
    assert got.shape == (len(wins), 1, 128, 1001) and got.dtype == np.float32
    for i, w in enumerate(wins):
        _check(got[i, 0], omel.compute_mel_spectrogram(w)[0, 0], omel.mel_power(w), tag=f"corpus track {i}")


def test_mel_matches_torchaudio_second_oracle():
    """A second, independent implementation (torchaudio's MelSpectrogram with slaney scale / norm, fp32 STFT) so a
    bug shared by the kernel and the numpy restatement cannot hide (two oracles agree to 2.7e-4 dB)."""
    import torch
    import torchaudio
    from audiomuse_ai_b200 import clap_analyzer as ca
    x = _windows()[6]
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=48000, n_fft=2048, hop_length=480, f_min=0.0, f_max=14000.0,
                                              n_mels=128, window_fn=torch.hann_window, power=2.0, center=True,
                                              pad_mode="reflect", norm="slaney", mel_scale="slaney")
    want = 10.0 * torch.log10(torch.clamp(ms(torch.from_numpy(x)), min=1e-10)).numpy()
    got = ca.compute_mel_spectrogram(x)[0, 0]
    err = np.abs(got - want)
    print(f"[mel parity] vs torchaudio: max |dB| error {err.max():.2e}, mean {err.mean():.2e}")
    assert err.max() <= 5e-3 and err.mean() <= 1e-4


def test_silence_is_minus_100_db():
    from audiomuse_ai_b200 import clap_analyzer as ca
    out = ca.compute_mel_spectrogram(np.zeros(480000, np.float32))
    assert out.shape == (1, 1, 128, 1001)
    assert np.abs(out + 100.0).max() < 2e-5


def test_int16_input_path_equals_float_path():
    from audiomuse_ai_b200 import clap_analyzer as ca, corpus
    pcm = np.stack([corpus.synth_track(i) for i in (1, 2, 7)])
    seg16 = np.stack([ca.pcm_to_segments(corpus.pcm16_to_float(p))[0] for p in pcm])
    # ... plus a window that contains every int16 value (the kernel scales without a division: all 65 536 inputs
    # must give the reference's (q / 32767.0).astype(float32), so both paths see identical samples)
    allv = np.resize(np.random.default_rng(0).permutation(np.arange(-32768, 32768)).astype(np.int16), seg16.shape[1])
    seg16 = np.concatenate([seg16, allv[None, :]], 0)
    a = ca.compute_mel_spectrogram_batch(seg16)
    b = ca.compute_mel_spectrogram_batch((seg16 / 32767.0).astype(np.float32))
    np.testing.assert_array_equal(a, b)


def test_single_call_signature_and_transposed_layout():
    from audiomuse_ai_b200 import clap_analyzer as ca
    x = _windows()[2]
    a = ca.compute_mel_spectrogram(x)
    old = ca.config.CLAP_AUDIO_MEL_TRANSPOSE
    try:
        ca.config.CLAP_AUDIO_MEL_TRANSPOSE = True
        b = ca.compute_mel_spectrogram(x)
    finally:
        ca.config.CLAP_AUDIO_MEL_TRANSPOSE = old
    assert a.shape == (1, 1, 128, 1001) and b.shape == (1, 1, 1001, 128)
    np.testing.assert_array_equal(a[0, 0].T, b[0, 0])


@pytest.mark.parametrize("n", [2048, 48000, 96001])
def test_other_window_lengths(n):
    from audiomuse_ai_b200 import clap_analyzer as ca
    rng = np.random.default_rng(n)
    x = (0.1 * rng.standard_normal(n)).astype(np.float32)
    got = ca.compute_mel_spectrogram(x)[0, 0]
    _check(got, omel.compute_mel_spectrogram(x)[0, 0], omel.mel_power(x))


def test_bad_config_is_reported():
    from audiomuse_ai_b200 import _lib, clap_analyzer as ca
    old = ca.config.CLAP_AUDIO_N_FFT
    try:
        ca.config.CLAP_AUDIO_N_FFT = 1000
        with pytest.raises(_lib.B200Error):
            ca.compute_mel_spectrogram(np.zeros(48000, np.float32))
    finally:
        ca.config.CLAP_AUDIO_N_FFT = old


@pytest.mark.parametrize("n_fft,n_mels,fmin,transpose", [(1024, 64, 50, True), (512, 40, 0, False), (1024, 128, 0, False)])
def test_other_fft_sizes_teacher_config(n_fft, n_mels, fmin, transpose):
    """config.py:377-392: the teacher model's mel is CLAP_AUDIO_N_FFT=1024, N_MELS=64, FMIN=50, transposed.
    compute_mel_spectrogram reads these at call time (clap_analyzer.py:431-436); shorter frames run as
    zero-padded 2048-point transforms whose every 2nd / 4th bin is the n_fft-point spectrum."""
    from audiomuse_ai_b200 import clap_analyzer as ca
    cfg = ca.config
    old = (cfg.CLAP_AUDIO_N_FFT, cfg.CLAP_AUDIO_N_MELS, cfg.CLAP_AUDIO_FMIN, cfg.CLAP_AUDIO_MEL_TRANSPOSE)
    wins = _windows()[[1, 2, 5]]
    try:
        cfg.CLAP_AUDIO_N_FFT, cfg.CLAP_AUDIO_N_MELS, cfg.CLAP_AUDIO_FMIN = n_fft, n_mels, fmin
        cfg.CLAP_AUDIO_MEL_TRANSPOSE = transpose
        got = ca.compute_mel_spectrogram_batch(wins)
        one = ca.compute_mel_spectrogram(wins[0])
    finally:
        (cfg.CLAP_AUDIO_N_FFT, cfg.CLAP_AUDIO_N_MELS, cfg.CLAP_AUDIO_FMIN, cfg.CLAP_AUDIO_MEL_TRANSPOSE) = old
    assert got.shape == ((3, 1, 1001, n_mels) if transpose else (3, 1, n_mels, 1001))
    np.testing.assert_array_equal(one[0], got[0])
    for i, w in enumerate(wins):
        want = omel.compute_mel_spectrogram(w, n_fft=n_fft, n_mels=n_mels, fmin=fmin)[0, 0]
        g = got[i, 0].T if transpose else got[i, 0]
        _check(g, want, omel.mel_power(w, n_fft=n_fft, n_mels=n_mels, fmin=fmin))


@pytest.mark.parametrize("seconds", [3.5, 30.0])
def test_musicnn_front_end_matches_oracle(seconds):
    """tasks/analysis.py:368-391: mel 96 / n_fft 512 / hop 256 at 16 kHz, center=False, log10(1 + 10000 x), patches of
    187 frames -- on the same kernel in its second framing / compression mode (am_mel_batch_ex)."""
    from audiomuse_ai_b200 import analysis_frontend as af
    rng = np.random.default_rng(int(seconds * 10))
    n = int(seconds * 16000)
    t = np.arange(n) / 16000.0
    x = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * np.sin(2 * np.pi * 3000 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    got = af.musicnn_patches(x)
    want = omel.musicnn_patches(x)
    assert got.shape == want.shape and got.dtype == np.float32 and got.shape[1:] == (187, 96)
    err = np.abs(got - want)
    print(f"[musicnn mel] {got.shape[0]} patches, max |err| = {err.max():.2e} (values up to {want.max():.2f})")
    assert err.max() <= 2e-4                       # log10(1 + 1e4 x): absolute, values in [0, ~8]
    assert af.musicnn_patches(x[: 187 * 256]) is None and omel.musicnn_patches(x[: 187 * 256]) is None   # one frame short
