"""Inputs of the Qwen loudness post-process fixtures: regenerated (numpy Generator streams are stable) instead of
stored, so that tests/golden/golden_qwen_v1.npz only holds a checksum of every input and the sub-sampled outputs."""
import numpy as np

STRIDE = 61      # stored output samples: every 61st, plus the first and last 512


def make_inputs():
    rng = np.random.default_rng(20261019)
    sr = 24000

    def tone(seconds, a0, a1, f=220.0, noise=1e-3, shape="lin"):
        n = int(round(seconds * sr))
        t = np.arange(n) / sr
        env = np.linspace(a0, a1, n) if shape == "lin" else a0 * (a1 / a0) ** (t / max(seconds, 1e-9))
        return (env * np.sin(2 * np.pi * f * t) * (0.7 + 0.3 * np.sin(2 * np.pi * 3.1 * t))
                + rng.normal(0, noise, n)).astype(np.float32)

    clips = [
        tone(10.0, 1.0, 0.2),                       # the reference test's decaying clip: windowed correction applies
        tone(6.0, 0.5, 0.5),                        # constant: gain range < 0.05 -> global normalisation only
        tone(3.9, 0.8, 0.1),                        # n <= 2 windows: no windowed pass
        tone(4.0 + 1 / 24000, 0.8, 0.1),            # n = 2 windows + 1 sample: windowed pass with 2 windows
        tone(13.37, 0.05, 0.9, shape="exp"),        # rising level: gains < 1, ragged tail after the last window
        tone(30.0, 0.9, 0.004, shape="exp"),        # deep decay: the +18 dB cap is hit
        np.zeros(24000 * 5, np.float32),            # silence: returned unchanged
        (rng.normal(0, 1e-9, 24000 * 5)).astype(np.float32),   # below the 1e-8 RMS gate
        np.concatenate([np.zeros(48000, np.float32), tone(8.0, 0.5, 0.3)]),   # silent first window: ref_rms < 1e-8
        np.concatenate([tone(4.0, 0.5, 0.4), np.zeros(48000, np.float32), tone(4.0, 0.3, 0.2)]),  # a silent window mid-clip
        tone(1.0, 2.5, 2.0),                        # loud and short: the tanh soft clip matters
        tone(0.001, 0.5, 0.5),                      # 24 samples
    ]
    return clips




def keep_index(n: int) -> np.ndarray:
    idx = np.unique(np.concatenate([np.arange(0, n, STRIDE), np.arange(min(n, 512)), np.arange(max(0, n - 512), n)]))
    return idx.astype(np.int64)
