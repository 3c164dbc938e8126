"""Oracle (TEST INFRASTRUCTURE): Whisper-style log-mel front end, numpy.

The reference reaches this stage through its STT validator
(/root/reference/src/rho_tts/validation/stt/stt_validator.py:78-107 -> transformers
ASR pipeline -> WhisperFeatureExtractor).  transformers is a third-party
dependency that is NOT under /root/reference (pyproject.toml:42,
`transformers>=4.40`, unpinned; installed and used as the pin: 5.5.0).  This
file restates its published algorithm:
  transformers/models/whisper/feature_extraction_whisper.py:135-164 (torch STFT path),
  :296-303 (pad / truncate to 480 000 samples),
  transformers/audio_utils.py:263-375, 453-544 (slaney mel scale, slaney norm).

Golden vectors made with the real WhisperFeatureExtractor are in tests/golden/.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

N_FFT = 400
HOP = 160
N_BINS = N_FFT // 2 + 1       # 201
SR16 = 16000
N_SAMPLES_30S = 480000
N_FRAMES_30S = 3000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    out = mel.copy()
    out[hi] = 15.0 + np.log(f[hi] / 1000.0) * logstep
    return out


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    out = f.copy()
    out[hi] = 1000.0 * np.exp(logstep * (m[hi] - 15.0))
    return out


def slaney_mel_filterbank(n_mels: int, n_bins: int = N_BINS, sr: int = SR16,
                          f_min: float = 0.0, f_max: float = 8000.0) -> np.ndarray:
    """(n_bins, n_mels) float64 triangular bank: slaney mel scale, slaney area norm,
    bin centres linspace(0, sr//2, n_bins)."""
    m_lo, m_hi = _hz_to_mel_slaney([f_min, f_max])
    centres = _mel_to_hz_slaney(np.linspace(m_lo, m_hi, n_mels + 2))
    bins = np.linspace(0, sr // 2, n_bins)
    gaps = np.diff(centres)
    rel = centres[None, :] - bins[:, None]                 # (n_bins, n_mels + 2)
    falling = -rel[:, :-2] / gaps[:-1]
    rising = rel[:, 2:] / gaps[1:]
    bank = np.maximum(0.0, np.minimum(falling, rising))
    bank = bank * (2.0 / (centres[2:n_mels + 2] - centres[:n_mels]))[None, :]
    return bank


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """torch.hann_window(n) (periodic) in fp32."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(F32)


def stft_power(w: np.ndarray) -> np.ndarray:
    """|STFT|^2 of a 16 kHz signal: n_fft 400, hop 160, periodic Hann, center=True
    with reflect padding; the last frame torch.stft produces is dropped.
    Returns (201, len(w)//160) fp32."""
    w = np.asarray(w, dtype=F32).reshape(-1)
    n = w.size
    half = N_FFT // 2
    if n <= half:
        raise ValueError("reflect padding needs more than n_fft/2 samples")
    p = np.concatenate([w[half:0:-1], w, w[n - 2:n - 2 - half:-1]]).astype(F32)
    n_frames = n // HOP              # (1 + n // HOP) frames from torch.stft, minus the dropped one
    frames = np.lib.stride_tricks.sliding_window_view(p, N_FFT)[::HOP][:n_frames]
    spec = np.fft.rfft((frames * hann_periodic()[None, :]).astype(F32), axis=1)
    spec = spec.astype(np.complex64)
    power = (np.abs(spec).astype(F32) ** 2).astype(F32)
    return power.T


def log_mel(w16: np.ndarray, n_mels: int = 80, pad_to_30s: bool = True) -> np.ndarray:
    """(n_mels, T) fp32 features for one clip.

    pad_to_30s=True : zero-pad / truncate to 480 000 samples -> T = 3000
                      (WhisperFeatureExtractor default, padding="max_length").
    pad_to_30s=False: T = len(w16) // 160 (padding="longest" on a single clip).
    """
    w = np.asarray(w16, dtype=F32).reshape(-1)
    if pad_to_30s:
        buf = np.zeros(N_SAMPLES_30S, dtype=F32)
        m = min(w.size, N_SAMPLES_30S)
        buf[:m] = w[:m]
        w = buf
    power = stft_power(w)                                              # (201, T)
    bank = slaney_mel_filterbank(n_mels).astype(F32)                    # (201, n_mels)
    mel = (bank.T @ power).astype(F32)
    ls = np.log10(np.maximum(mel, F32(1e-10))).astype(F32)
    ls = np.maximum(ls, ls.max() - F32(8.0))
    return ((ls + F32(4.0)) / F32(4.0)).astype(F32)


def log_mel_truth64(y24: np.ndarray, n_mels: int = 80) -> np.ndarray:
    """What the reference chain resample(24k -> 16k) -> WhisperFeatureExtractor(30 s pad) computes, evaluated in float64
    with the SAME fp32 tables (taps, window, filterbank): the value every fp32 implementation -- torch / MKL in the
    reference, pocketfft in this oracle, the GPU kernels -- approximates.  On bins 70..80 dB below a frame's peak an fp32
    FFT is itself up to ~1e-4 (after the log) from this value, so parity tests at the 1e-4 level judge the GPU against
    this truth and allow its distance to another fp32 implementation to exceed 1e-4 only by that implementation's own
    distance to the truth (tests/test_full_parity.py)."""
    from .resample import sinc_resample_kernel
    taps, width, orig, new = sinc_resample_kernel(24000, 16000)
    y = np.asarray(y24, dtype=np.float64).reshape(-1)
    L = y.size
    xp = np.zeros(L + 2 * width + orig)
    xp[width:width + L] = y
    fr = np.lib.stride_tricks.sliding_window_view(xp, taps.shape[1])[::orig]
    w = (fr @ taps.astype(np.float64).T).reshape(-1)[:-(-new * L // orig)]
    buf = np.zeros(N_SAMPLES_30S)
    buf[:min(w.size, N_SAMPLES_30S)] = w[:N_SAMPLES_30S]
    p = np.concatenate([buf[200:0:-1], buf, buf[-2:-202:-1]])
    frames = np.lib.stride_tricks.sliding_window_view(p, 400)[::160][:3000]
    power = np.abs(np.fft.rfft(frames * hann_periodic().astype(np.float64)[None, :], axis=1)) ** 2
    mel = slaney_mel_filterbank(n_mels).astype(F32).astype(np.float64).T @ power.T
    ls = np.log10(np.maximum(mel, 1e-10))
    ls = np.maximum(ls, ls.max() - 8.0)
    return (ls + 4.0) / 4.0
