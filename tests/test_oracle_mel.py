"""Cross-check the mel oracle against an independent implementation (torchaudio) and
its own float64 evaluation.  (librosa itself is not installable: parity unpinned.)"""
import numpy as np
import pytest
import torch

from oracle import mel as omel


def _clip(seed, n=96000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    x = 0.2 * np.sin(2 * np.pi * 440 * t) + 0.05 * np.sin(2 * np.pi * 9000 * t) + 0.02 * rng.standard_normal(n)
    return x.astype(np.float32)


def test_filterbank_shape_and_support():
    w = omel.mel_filterbank()
    assert w.shape == (128, 1025) and w.dtype == np.float32
    nz = np.nonzero(w.sum(0))[0]
    assert nz.max() == 597            # fmax = 14 kHz: bins above 597 carry no weight
    assert int((w > 0).sum()) == 1176  # SURVEY 8(a) probe
    assert (w.sum(1) > 0).all()       # no empty filters


def test_filterbank_matches_torchaudio():
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(1025, 0.0, 14000.0, 128, 48000, norm="slaney", mel_scale="slaney")
    np.testing.assert_allclose(omel.mel_filterbank(), fb.numpy().T, atol=2e-7)


def test_mel_matches_torchaudio():
    import torchaudio
    x = _clip(0)
    ours = omel.compute_mel_spectrogram(x)[0, 0]
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=48000, n_fft=2048, hop_length=480, f_min=0.0,
                                              f_max=14000.0, n_mels=128, power=2.0, center=True,
                                              pad_mode="reflect", norm="slaney", mel_scale="slaney")
    ref = 10.0 * torch.log10(torch.clamp(ms(torch.from_numpy(x)), min=1e-10)).numpy()
    assert ours.shape == ref.shape == (128, 201)
    assert np.abs(ours - ref).max() < 2e-3


def test_mel_close_to_float64_truth():
    x = _clip(1)
    p32 = omel.mel_power(x)
    p64 = omel.mel_power_f64(x)
    rel = np.abs(p32 - p64) / (np.abs(p64) + 1e-30)
    assert rel.max() < 1e-5


def test_silence_is_minus_100_db():
    """amin=1e-10 floor: -100 dB up to the float32 log10 ulp (numpy's log10f gives -100.00001)."""
    out = omel.compute_mel_spectrogram(np.zeros(48000, np.float32))
    assert out.shape == (1, 1, 128, 101)
    assert np.abs(out + 100.0).max() < 2e-5


def test_layouts():
    x = _clip(2, 480000)
    a = omel.compute_mel_spectrogram(x)
    b = omel.compute_mel_spectrogram(x, transpose=True)
    assert a.shape == (1, 1, 128, 1001) and b.shape == (1, 1, 1001, 128)
    np.testing.assert_array_equal(a[0, 0].T, b[0, 0])
