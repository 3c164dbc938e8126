"""Oracle (TEST INFRASTRUCTURE): QwenTTS._post_process_audio, numpy restatement.

Follows /root/reference/src/rho_tts/providers/qwen.py:268-378 step by step:
  :283-289  squeeze; overall RMS (fp32); < 1e-8 -> unchanged
  :291-299  window = int(sr * 2.0); windowed correction only when n > 2 * window
  :326-378  per-window RMS (fp32 -> python float), gain = first / rms capped at +18 dB, skipped when the
            gain range is < 0.05; two passes of a 3-tap moving average with fixed end points (python floats);
            np.interp over sample indices (float64) from the window centres -> float32 envelope; multiply
  :301-307  global gain to -23 dBFS: 20*log10(rms) in fp32, the dB arithmetic and 10**(g/20) in python floats
  :309-311  tanh(x / 0.95) * 0.95

Pinned by golden vectors produced by the reference's own method (tests/golden/make_golden_qwen.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

TARGET_RMS_DB = -23.0
WINDOW_SEC = 2.0
MAX_GAIN_DB = 18.0
MAX_AMPLITUDE = 0.95


def _rms32(x: np.ndarray) -> np.float32:
    return np.sqrt(np.mean(x * x, dtype=F32), dtype=F32)


def windowed_gains(x: np.ndarray, window: int):
    """(apply, smoothed gains as python floats) of _apply_windowed_normalization (qwen.py:326-370)."""
    n = x.size
    n_windows = n // window
    if n_windows < 2:
        return False, []
    rms = [float(_rms32(x[i * window:(i + 1) * window])) for i in range(n_windows)]
    ref = rms[0]
    if ref < 1e-8:
        return False, []
    cap = 10 ** (MAX_GAIN_DB / 20)
    gains = [1.0 if r < 1e-8 else min(ref / r, cap) for r in rms]
    if max(gains) - min(gains) < 0.05:
        return False, gains
    sm = list(gains)
    for _ in range(2):
        nxt = list(sm)
        for i in range(1, len(sm) - 1):
            nxt[i] = (sm[i - 1] + sm[i] + sm[i + 1]) / 3
        sm = nxt
    return True, sm


def gain_envelope(n: int, window: int, smoothed) -> np.ndarray:
    centres = np.array([(i + 0.5) * window for i in range(len(smoothed))])
    return np.interp(np.arange(n, dtype=np.float64), centres, smoothed).astype(F32)


def post_process(audio: np.ndarray, sr: int = 24000) -> np.ndarray:
    """QwenTTS._post_process_audio on one clip (any shape with one non-trivial axis); returns the same shape."""
    a = np.asarray(audio, dtype=F32)
    shape = a.shape
    x = a.reshape(-1).copy()
    if x.size == 0 or _rms32(x) < 1e-8:
        return a.copy()
    window = int(sr * WINDOW_SEC)
    if x.size > 2 * window:
        apply, sm = windowed_gains(x, window)
        if apply:
            x = (x * gain_envelope(x.size, window, sm)).astype(F32)
    rms = _rms32(x)
    if rms > 1e-8:
        current_db = float(F32(20.0) * np.log10(rms, dtype=F32))
        gain_linear = 10 ** ((TARGET_RMS_DB - current_db) / 20)
        x = (x * F32(gain_linear)).astype(F32)
    x = (np.tanh(x / F32(MAX_AMPLITUDE), dtype=F32) * F32(MAX_AMPLITUDE)).astype(F32)
    return x.reshape(shape)
