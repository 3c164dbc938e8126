"""Oracle (TEST INFRASTRUCTURE): polyphase windowed-sinc resampler, numpy.

The reference calls `torchaudio.functional.resample` (base_tts.py:632; the
24 kHz -> 16 kHz instance happens in its dependencies).
torchaudio is a third-party dependency that is NOT under /root/reference
(pyproject.toml:31, `torchaudio>=2.0`, unpinned; installed and used as the pin:
2.11.0).  This file restates its published algorithm
(torchaudio/functional/functional.py:1305-1432, `_get_sinc_resample_kernel` and
`_apply_sinc_resample_kernel`, method "sinc_interp_hann", width 6, rolloff 0.99).

Golden vectors made with the real torchaudio function are in tests/golden/.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

F32 = np.float32


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6,
                         rolloff: float = 0.99) -> Tuple[np.ndarray, int, int, int]:
    """Returns (taps[new, K] fp32, width, orig, new) with orig/new reduced by their gcd.

    For an fp32 waveform torchaudio builds the taps in fp32 (dtype=waveform.dtype),
    so every step below is carried in fp32 as well.
    """
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff                      # python double
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = (np.arange(-width, width + orig, dtype=F32) / F32(orig)).astype(F32)      # [K]
    phase = (np.arange(0, -new, -1, dtype=F32) / F32(new)).astype(F32)               # [new]
    t = (phase[:, None] + idx[None, :]).astype(F32)
    t = (t * F32(base)).astype(F32)
    t = np.clip(t, F32(-lowpass_filter_width), F32(lowpass_filter_width)).astype(F32)
    window = np.cos((t * F32(math.pi) / F32(lowpass_filter_width) / F32(2)).astype(F32), dtype=F32) ** 2
    t = (t * F32(math.pi)).astype(F32)
    scale = F32(base / orig)
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0, F32(1.0), np.sin(t, dtype=F32) / t).astype(F32)
    taps = (sinc * (window.astype(F32) * scale)).astype(F32)
    return taps, width, orig, new


def resample(x: np.ndarray, orig_freq: int = 24000, new_freq: int = 16000) -> np.ndarray:
    """y[new*m + p] = sum_i xpad[orig*m + i] * taps[p, i], xpad = x zero-padded by
    (width, width + orig); truncated to ceil(new * L / orig) samples."""
    x = np.asarray(x, dtype=F32).reshape(-1)
    if orig_freq == new_freq:
        return x.copy()
    taps, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    L = x.size
    K = taps.shape[1]
    xp = np.zeros(L + 2 * width + orig, dtype=F32)
    xp[width:width + L] = x
    n_blocks = (xp.size - K) // orig + 1
    frames = np.lib.stride_tricks.sliding_window_view(xp, K)[::orig][:n_blocks]     # [M, K]
    y = (frames @ taps.T).astype(F32).reshape(-1)                                   # [M*new]
    target = int(math.ceil(new * L / orig))
    return y[:target]
