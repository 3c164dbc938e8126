"""Oracle (TEST INFRASTRUCTURE): the MFCC statistics of the drift classifier's feature vector, numpy.

The reference's `extract_features` (validation/classifier/trainer.py:49-52) computes
    y, sr = librosa.load(path, sr=16000)
    mfcc = librosa.feature.mfcc(y=y, sr=sr, n_mfcc=13);  mfcc_mean = mean(mfcc, axis=1);  mfcc_std = std(mfcc, axis=1)
librosa (pyproject.toml:53, `librosa>=0.10`) is a third-party dependency that is NOT under /root/reference and NOT in
this image: PARITY UNPINNED against librosa itself.  This file restates its published algorithm (librosa 0.10:
feature/spectral.py `mfcc` -> `melspectrogram` -> core/spectrum.py `stft`, `power_to_db`, filters.py `mel`):
    stft      n_fft 2048, hop 512, periodic hann, center=True with ZERO padding (pad_mode="constant" since 0.10)
    power     |X|^2
    mel       128 slaney-scale, slaney-normalised triangles, 0 .. sr/2
    dB        10 log10(max(1e-10, S)), then max(., max(.) - 80)            (ref = 1.0, top_db = 80)
    dct       scipy.fftpack.dct(type=2, norm="ortho") along the mel axis, first 13 rows
and is pinned on the pieces of that chain that ARE in this image: transformers.audio_utils.spectrogram (a port of
librosa's stft / mel / power_to_db, which transformers tests against librosa) for the dB mel spectrogram, and
scipy.fft.dct for the transform (tests/golden/make_golden_mfcc.py).  The 16 kHz signal is an input here: librosa.load
resamples with soxr, which is not restated.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
N_FFT = 2048
HOP = 512
N_MELS = 128
N_MFCC = 13
SR = 16000


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), 200.0 * m / 3.0)


def mel_filterbank(n_mels: int = N_MELS, n_fft: int = N_FFT, sr: int = SR) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels) (slaney scale and norm), [n_mels, n_fft/2 + 1] fp32."""
    n_bins = n_fft // 2 + 1
    freqs = np.linspace(0.0, sr / 2.0, n_bins)
    centres = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    diff = np.diff(centres)
    ramps = centres[:, None] - freqs[None, :]
    lower = -ramps[:-2] / diff[:-1, None]
    upper = ramps[2:] / diff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (centres[2:] - centres[:-2]))[:, None]
    return w.astype(F32)


def hann(n: int = N_FFT) -> np.ndarray:
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(F32)


def dct_matrix(n_out: int = N_MFCC, n_in: int = N_MELS) -> np.ndarray:
    """Rows of the orthonormal DCT-II: y[k] = s_k sum_m x[m] cos(pi k (2m + 1) / (2N))."""
    m = np.arange(n_in, dtype=np.float64)
    k = np.arange(n_out, dtype=np.float64)[:, None]
    d = np.cos(np.pi * k * (2.0 * m + 1.0) / (2.0 * n_in)) * math.sqrt(2.0 / n_in)
    d[0] *= math.sqrt(0.5)
    return d.astype(F32)


def mel_db(y: np.ndarray) -> np.ndarray:
    """power_to_db(melspectrogram(y)) -> [128, T] fp32, T = 1 + len(y) // 512."""
    y = np.asarray(y, dtype=F32).reshape(-1)
    yp = np.pad(y, (N_FFT // 2, N_FFT // 2))
    T = 1 + y.size // HOP
    fr = np.lib.stride_tricks.sliding_window_view(yp, N_FFT)[::HOP][:T]
    spec = np.fft.rfft((fr * hann()[None, :]).astype(F32), axis=1)
    power = (spec.real.astype(F32) ** 2 + spec.imag.astype(F32) ** 2).astype(F32)          # [T, 1025]
    mel = (power @ mel_filterbank().T).astype(F32)                                         # [T, 128]
    db = (10.0 * np.log10(np.maximum(F32(1e-10), mel))).astype(F32)
    db = np.maximum(db, db.max() - F32(80.0)).astype(F32)
    return db.T


def mfcc(y: np.ndarray) -> np.ndarray:
    """librosa.feature.mfcc(y=y, sr=16000, n_mfcc=13) -> [13, T] fp32."""
    return (dct_matrix() @ mel_db(y)).astype(F32)


def mfcc_stats(y: np.ndarray) -> np.ndarray:
    """[mfcc_mean (13), mfcc_std (13)] as trainer.py:51-52 computes them (np.std: population)."""
    c = mfcc(y).astype(np.float64)
    return np.concatenate([c.mean(axis=1), c.std(axis=1)]).astype(F32)
