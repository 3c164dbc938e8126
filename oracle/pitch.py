"""Oracle (TEST INFRASTRUCTURE): pitch shift, numpy.

The reference's `_apply_speed_pitch` (base_tts.py:639-648) calls
`torchaudio.functional.pitch_shift(audio, sample_rate, pitch_semitones)`.  torchaudio is a third-party dependency
that is NOT under /root/reference (pyproject.toml:31, unpinned; installed and used as the pin: 2.11.0).  This file
restates its published algorithm, torchaudio/functional/functional.py:

  pitch_shift        :1596-1641   stretch -> resample(int(sr / rate), sr) -> crop / zero-pad to the input length
  _stretch_waveform  :1644-1694   stft(512, hop 128, periodic hann, center, reflect) -> phase_vocoder -> istft
  phase_vocoder      :723-803
  resample           :1305-1432   (oracle/resample.py restates the tap formula; here only the taps inside the window
                                   are evaluated -- outside it torchaudio's fp32 taps are ~1e-23, and for a ratio like
                                   26939:24000 the dense table would hold 6.5e8 of them)

Every fp32 operation of the phase vocoder is restated in fp32 IN TORCH'S ORDER, because the result depends on it:
the phase accumulator reaches ~1e6 rad (ulp 0.06) on the top bins, so fp32 pitch_shift differs from the same
algorithm in fp64 by 1e-4..5e-4 -- "parity" is with the fp32 reference's own rounding:
  * `torch.arange(0, T, rate, dtype=float32)` on CPU is evaluated by the vectorised range kernel
    (ATen/native/cpu/RangeFactoriesKernel.cpp): blocks of 2 x V lanes as float(float(rate * i0) + k * rate), the tail
    (J mod 2V elements) as float(rate * i); V = 8 in the torch 2.11 wheels (checked element by element against
    torch.arange for 48 rates x 15 lengths; exact while J <= 32768, the kernel's parallel grain).  `arange_f32`
    restates that; V is a parameter (0 = float(rate * i)).
  * `torch.cumsum` of fp32 on CPU accumulates in double and rounds every element to fp32.
  * `phase / (2 * math.pi)` is a true fp32 division by float(2 pi).

Golden vectors made with the real torchaudio function (and through the reference method) are in tests/golden/
(make_golden_pitch.py).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
N_FFT = 512
HOP = 128
N_FREQ = N_FFT // 2 + 1


def hann512() -> np.ndarray:
    """torch.hann_window(512) (periodic), fp32."""
    n = np.arange(N_FFT, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)).astype(F32)


def pitch_rate(n_steps: float, bins_per_octave: int = 12) -> float:
    """functional.py:1638."""
    return 2.0 ** (-float(n_steps) / bins_per_octave)


def arange_f32(count: int, step: float, vec: int = 8) -> np.ndarray:
    """torch.arange(0, end, step, dtype=float32) on CPU, `count` = ceil(end / step) elements."""
    i = np.arange(count, dtype=np.float64)
    if vec <= 0:
        return (i * step).astype(F32)
    blk = 2 * vec
    full = (count // blk) * blk
    i0 = (np.arange(count) // vec) * vec
    base = (i0.astype(np.float64) * step).astype(F32).astype(np.float64)
    out = (base + (i - i0) * step).astype(F32)
    out[full:] = (i[full:] * step).astype(F32)
    return out


def linspace_f32(end: float, n: int) -> np.ndarray:
    """torch.linspace(0, end, n) (float32, CPU): step = float(end) / float(n - 1) in fp32; the first half is
    float(step * i), the second half end - step * (n - 1 - i) as ONE fused multiply-subtract
    (RangeFactoriesKernel.cpp linspace; checked element by element against torch for n = 201, 257, 513)."""
    e = F32(end)
    step = F32(e / F32(n - 1))
    i = np.arange(n)
    lo = (step * i.astype(F32)).astype(F32)
    hi = (np.float64(e) - np.float64(step) * (n - 1 - i)).astype(F32)     # exact product, one rounding = fma
    return np.where(i < n // 2, lo, hi).astype(F32)


def stft512(x: np.ndarray) -> np.ndarray:
    """torch.stft(x, 512, 128, 512, hann, center=True, pad_mode='reflect', onesided) -> complex64 [257, T]."""
    x = np.asarray(x, dtype=F32).reshape(-1)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + (xp.size - N_FFT) // HOP
    fr = np.lib.stride_tricks.sliding_window_view(xp, N_FFT)[::HOP][:n_frames]
    return np.fft.rfft((fr * hann512()[None, :]).astype(F32), axis=1).astype(np.complex64).T


def phase_vocoder(spec: np.ndarray, rate: float, vec: int = 8):
    """functional.py:723-803 on one [257, T] complex64 spectrogram -> (mag fp32 [257, J], phase_acc fp32 [257, J])."""
    n_freq, T = spec.shape
    J = int(math.ceil(T / rate))
    ts = arange_f32(J, rate, vec)
    alphas = np.fmod(ts, F32(1.0)).astype(F32)
    phase_advance = linspace_f32(math.pi * HOP, n_freq)[:, None]
    phase_0 = np.angle(spec[:, :1]).astype(F32)
    sp = np.concatenate([spec, np.zeros((n_freq, 2), np.complex64)], axis=1)
    i0 = ts.astype(np.int64)
    i1 = (ts + F32(1.0)).astype(F32).astype(np.int64)
    s0, s1 = sp[:, i0], sp[:, i1]
    a0 = np.arctan2(s0.imag, s0.real).astype(F32)
    a1 = np.arctan2(s1.imag, s1.real).astype(F32)
    n0 = np.abs(s0).astype(F32)
    n1 = np.abs(s1).astype(F32)
    phase = ((a1 - a0).astype(F32) - phase_advance).astype(F32)
    two_pi = F32(2 * math.pi)
    phase = (phase - (two_pi * np.round((phase / two_pi).astype(F32)).astype(F32)).astype(F32)).astype(F32)
    phase = (phase + phase_advance).astype(F32)
    phase = np.concatenate([phase_0, phase[:, :-1]], axis=1)
    phase_acc = np.cumsum(phase.astype(np.float64), axis=1).astype(F32)
    mag = ((alphas[None, :] * n1).astype(F32) + ((F32(1.0) - alphas)[None, :] * n0).astype(F32)).astype(F32)
    return mag, phase_acc


def istft512(spec: np.ndarray, length: int) -> np.ndarray:
    """torch.istft(spec, 512, 128, 512, hann, length=length) for a [257, J] complex64 spectrogram."""
    w = hann512()
    J = spec.shape[1]
    fr = np.fft.irfft(spec.T.astype(np.complex64), n=N_FFT, axis=1).astype(F32) * w[None, :]
    full = N_FFT + HOP * (J - 1)
    y = np.zeros(full, dtype=F32)
    env = np.zeros(full, dtype=F32)
    w2 = (w * w).astype(F32)
    for j in range(J):
        y[j * HOP:j * HOP + N_FFT] += fr[j]
        env[j * HOP:j * HOP + N_FFT] += w2
    start = N_FFT // 2
    end = start + length
    ys, es = y[start:end], env[start:end]
    out = (ys / es).astype(F32)
    if out.size < length:
        out = np.concatenate([out, np.zeros(length - out.size, F32)])
    return out


def resample_windowed(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """torchaudio resample (sinc_interp_hann, width 6, rolloff 0.99), only the taps inside the window."""
    x = np.asarray(x, dtype=F32).reshape(-1)
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    if orig == new:
        return x.copy()
    lpw = 6
    base = min(orig, new) * 0.99
    width = math.ceil(lpw * orig / base)
    L = x.size
    target = int(math.ceil(new * L / orig))
    o = np.arange(target, dtype=np.int64)
    p, q = o % new, o // new
    # tap i (0 <= i < 2*width + orig) of phase p sits at t = (-p/new + (i - width)/orig) * base; |t| < lpw
    centre = p.astype(np.float64) * orig / new + width
    W = 2 * width + 2
    i_lo = np.maximum(0, np.floor(centre - lpw * orig / base).astype(np.int64))
    i = i_lo[:, None] + np.arange(W)[None, :]
    i = np.minimum(i, 2 * width + orig - 1)
    ph = (-(p.astype(F32)) / F32(new)).astype(F32)
    idx = ((i - width).astype(F32) / F32(orig)).astype(F32)
    t = (ph[:, None] + idx).astype(F32)
    t = (t * F32(base)).astype(F32)
    t = np.clip(t, F32(-lpw), F32(lpw)).astype(F32)
    window = np.cos((t * F32(math.pi) / F32(lpw) / F32(2)).astype(F32), dtype=F32) ** 2
    t = (t * F32(math.pi)).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0, F32(1.0), np.sin(t, dtype=F32) / t).astype(F32)
    taps = (sinc * (window.astype(F32) * F32(base / orig))).astype(F32)
    dup = np.concatenate([np.zeros((target, 1), bool), i[:, 1:] == i[:, :-1]], axis=1)      # clamped repeats
    taps[dup] = 0
    src = q[:, None] * orig + i - width                                                       # index into x
    ok = (src >= 0) & (src < L)
    xv = np.where(ok, x[np.clip(src, 0, max(L - 1, 0))], F32(0))
    return (taps.astype(np.float64) * xv).sum(axis=1).astype(F32)


def pitch_shift(x: np.ndarray, sample_rate: int, n_steps: float, vec: int = 8) -> np.ndarray:
    """torchaudio.functional.pitch_shift(x[None], sample_rate, n_steps)[0] for a 1-D fp32 clip."""
    x = np.asarray(x, dtype=F32).reshape(-1)
    L = x.size
    rate = pitch_rate(n_steps)
    spec = stft512(x)
    mag, pacc = phase_vocoder(spec, rate, vec)
    st = (mag * np.cos(pacc.astype(np.float64))).astype(F32) + 1j * (mag * np.sin(pacc.astype(np.float64))).astype(F32)
    len_stretch = int(round(L / rate))
    w = istft512(st.astype(np.complex64), len_stretch)
    y = resample_windowed(w, int(sample_rate / rate), sample_rate)
    if y.size > L:
        return y[:L].copy()
    return np.concatenate([y, np.zeros(L - y.size, F32)])
