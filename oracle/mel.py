"""Oracle: log-mel spectrogram.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``tasks/clap_analyzer.py:417-464`` (compute_mel_spectrogram), i.e.

    librosa.feature.melspectrogram(y, sr=48000, n_fft=2048, hop_length=480,
        win_length=2048, window='hann', center=True, pad_mode='reflect', power=2.0,
        n_mels=128, fmin=0, fmax=14000)
    librosa.power_to_db(mel, ref=1.0, amin=1e-10, top_db=None)

librosa==0.11.0 (requirements/common.txt:26) is a third-party dependency that is
absent from /root/reference and not installable here; its published algorithm is
restated below with librosa's own dtype discipline:

  * reflect-pad n_fft//2 on both sides (np.pad mode='reflect'), frames of n_fft at
    stride hop, n_frames = 1 + len(y)//hop;
  * periodic Hann (scipy.signal.get_window('hann', N, fftbins=True)) in float64;
    window * frame is a float64 product, the rFFT runs in float64 and the result is
    stored as complex64 (librosa.stft with float32 input -> complex64 output);
  * power = np.abs(complex64)**2 in float32;
  * mel basis = librosa.filters.mel(htk=False, norm='slaney', dtype=float32): 130
    Slaney-mel-spaced edge frequencies, triangles from np.subtract.outer ramps,
    each row scaled by 2/(f[i+2]-f[i]);
  * mel = basis(float32) @ power(float32); dB = 10*log10(max(1e-10, mel)) in float32.

PARITY UNPINNED: no reference unit test touches this function and librosa cannot
be imported here.  ``tests/test_oracle_mel.py`` cross-checks this restatement against
torchaudio's independent MelSpectrogram(norm='slaney', mel_scale='slaney').
"""
from __future__ import annotations

import numpy as np

# Defaults: config.py:386-392 (student model, "EfficientAT epoch 36")
SR = 48000
N_FFT = 2048
HOP = 480
N_MELS = 128
FMIN = 0.0
FMAX = 14000.0


def hz_to_mel(f):
    """Slaney mel scale (librosa.hz_to_mel, htk=False)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (f - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(m):
    """Slaney mel scale inverse (librosa.mel_to_hz, htk=False)."""
    m = np.asanyarray(m, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * m
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -> f32[n_mels, 1+n_fft//2]."""
    if fmax is None:
        fmax = sr / 2.0
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True) in float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def stft_power(y, n_fft=N_FFT, hop=HOP):
    """|STFT|^2 as librosa computes it for float32 input: f32[1+n_fft//2, n_frames]."""
    y = np.asarray(y, dtype=np.float32)
    ypad = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = ypad[idx]  # [T, n_fft] float32
    win = hann_periodic(n_fft)
    spec = np.fft.rfft(win[None, :] * frames, axis=1)  # float64 compute
    spec = spec.astype(np.complex64)  # librosa stores complex64
    mag = np.abs(spec)  # float32
    return (mag**2).T.astype(np.float32, copy=False)  # [bins, T]


def mel_power(y, sr=SR, n_fft=N_FFT, hop=HOP, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """Mel power spectrogram f32[n_mels, T] (before dB)."""
    power = stft_power(y, n_fft, hop)
    basis = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    return basis @ power


def power_to_db(s, amin=1e-10):
    """librosa.power_to_db(ref=1.0, amin=1e-10, top_db=None) on float32."""
    s = np.asarray(s, dtype=np.float32)
    log_spec = 10.0 * np.log10(np.maximum(np.float32(amin), s))
    log_spec -= 10.0 * np.log10(np.maximum(np.float32(amin), np.float32(1.0)))
    return log_spec.astype(np.float32, copy=False)


def compute_mel_spectrogram(audio_data, sr=SR, n_fft=N_FFT, hop=HOP, n_mels=N_MELS,
                            fmin=FMIN, fmax=FMAX, transpose=False):
    """Restates clap_analyzer.compute_mel_spectrogram: f32 (1,1,n_mels,T) or (1,1,T,n_mels)."""
    mel = power_to_db(mel_power(audio_data, sr, n_fft, hop, n_mels, fmin, fmax))
    if transpose:
        mel = mel.T
    return np.ascontiguousarray(mel[np.newaxis, np.newaxis, :, :], dtype=np.float32)


def mel_power_f64(y, sr=SR, n_fft=N_FFT, hop=HOP, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """All-float64 mel power: the numerically 'true' value, used to size tolerances."""
    y = np.asarray(y, dtype=np.float64)
    ypad = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(hann_periodic(n_fft)[None, :] * ypad[idx], axis=1)
    power = (spec.real**2 + spec.imag**2).T
    return mel_filterbank(sr, n_fft, n_mels, fmin, fmax).astype(np.float64) @ power


def musicnn_patches(audio, sr=16000, n_mels=96, hop=256, n_fft=512, frame_size=187):
    """Restates tasks/analysis.py:368-391 (the MusiCNN front end): librosa.feature.melspectrogram(y, sr, n_fft=512,
    hop_length=256, n_mels=96, window='hann', center=False, power=2.0, norm='slaney', htk=False) ->
    log10(1 + 10000 x) -> non-overlapping patches of 187 frames, transposed to (n_patches, 187, 96) float32.
    Returns None when the track is too short for one patch (the reference logs and returns None)."""
    y = np.asarray(audio, dtype=np.float32)
    if len(y) < n_fft:
        return None
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(hann_periodic(n_fft)[None, :] * y[idx], axis=1).astype(np.complex64)
    power = (np.abs(spec) ** 2).T.astype(np.float32, copy=False)            # [bins, T]
    mel = mel_filterbank(sr, n_fft, n_mels, 0.0, sr / 2.0) @ power          # float32
    log_mel = np.log10(1 + 10000 * mel)
    patches = [log_mel[:, i:i + frame_size] for i in range(0, log_mel.shape[1] - frame_size + 1, frame_size)]
    if not patches:
        return None
    return np.array(patches).transpose(0, 2, 1).astype(np.float32)
