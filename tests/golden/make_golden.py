"""Generates tests/golden/golden_v1.npz by running the REFERENCE itself (authoring container only).

  in-tree stages : rho_tts.base_tts.BaseTTS methods imported from /root/reference/src
                   (_trim_silence, _smooth_segment_join, _validate_sound_decay; base_tts.py:297-536)
  resample       : torchaudio.functional.resample(x, 24000, 16000)        (torchaudio 2.11.0)
  log-mel        : transformers.WhisperFeatureExtractor(feature_size=80|128) (transformers 5.5.0)
  cosine         : the numpy expression of base_tts.py:341-344

The reference ships no golden vectors of its own, so these outputs ARE the pin.
Inputs are stored in the file as well, so the fixtures do not depend on any RNG being reproducible.

    python tests/golden/make_golden.py
"""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import torch  # noqa: E402
import torchaudio  # noqa: E402
import transformers  # noqa: E402
from rho_tts.base_tts import BaseTTS  # noqa: E402

logging.getLogger("rho_tts.base_tts").setLevel(logging.ERROR)


class Ref:
    """The reference's own test idiom: borrow the unbound methods onto a plain object
    (tests/test_audio_processing.py:7-29 in the reference)."""

    def __init__(self, sr=24000):
        self.device = "cpu"
        self.silence_threshold_db = -50.0
        self.crossfade_duration_sec = 0.05
        self.trim_silence = True
        self.fade_duration_sec = 0.02
        self.force_sentence_split = True
        self.inter_sentence_pause_sec = 0.1
        self.sound_decay_threshold = 0.3
        self._sr = sr

    @property
    def sample_rate(self):
        return self._sr


for _n in ("_trim_silence", "_remove_dc_offset", "_apply_fades", "_smooth_segment_join", "_validate_sound_decay"):
    setattr(Ref, _n, getattr(BaseTTS, _n))


def make_inputs():
    rng = np.random.default_rng(20261018)
    sr = 24000

    def tone(L, lead, trail, rho, amp=0.3, dc=1e-3):
        t = np.arange(L) / sr
        x = amp * np.sin(2 * np.pi * rng.uniform(90, 300) * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * t))
        x *= np.linspace(1.0, rho, L)
        x[:lead] = 0
        if trail:
            x[L - trail:] = 0
        return (x + rng.normal(0, 1e-3, L) + dc).astype(np.float32)

    clips = [
        tone(36000, 2500, 4100, 0.9),
        tone(24000, 0, 3000, 0.05),                       # decays: rejected
        tone(30001, 5000, 0, 0.5),
        tone(12000, 1300, 1700, 1.1),
        (rng.normal(0, 0.2, 6000)).astype(np.float32),     # loud everywhere
        (rng.normal(0, 1e-4, 2400)).astype(np.float32),    # all silent
        tone(700, 100, 100, 1.0),
        tone(240, 0, 0, 1.0),
        tone(5000, 400, 300, 0.2, dc=2e-3),
    ]
    items = [[0, 1, 2], [3, 4], [2, 5, 3], [5, 5], [5, 5, 5], [6, 7, 8, 0], [1], [5], [4, 6]]
    return clips, items


def main():
    ref = Ref()
    clips, items = make_inputs()
    out = {"n_clips": np.int32(len(clips)), "n_items": np.int32(len(items)),
           "versions": np.array([torch.__version__, torchaudio.__version__, transformers.__version__])}
    for i, x in enumerate(clips):
        out[f"clip{i}"] = x
        bounds = []
        for tf in range(4):
            r = ref._trim_silence(torch.from_numpy(x.copy()), bool(tf & 1), bool(tf & 2))
            # recover start from the view's storage offset
            bounds.append((r.storage_offset(), r.storage_offset() + r.shape[-1], r.dim()))
        out[f"trim{i}"] = np.asarray(bounds, dtype=np.int32)           # [flags][start, end, dim]
        y = ref._smooth_segment_join([torch.from_numpy(x.copy())])
        out[f"post{i}"] = y.numpy().reshape(-1).copy()
        out[f"post_dim{i}"] = np.int32(y.dim())
        ratio, ok = ref._validate_sound_decay(y)
        out[f"decay{i}"] = np.asarray([ratio, float(ok)], dtype=np.float64)
    for k, idx in enumerate(items):
        y = ref._smooth_segment_join([torch.from_numpy(clips[j].copy()) for j in idx])
        out[f"item{k}_idx"] = np.asarray(idx, dtype=np.int32)
        out[f"item{k}"] = y.numpy().reshape(-1).copy()
        out[f"item_dim{k}"] = np.int32(y.dim())
        ratio, ok = ref._validate_sound_decay(y)
        out[f"item_decay{k}"] = np.asarray([ratio, float(ok)], dtype=np.float64)
    # resample (torchaudio) on the post-processed clips 0..3 and three odd lengths
    for i in range(4):
        y = torchaudio.functional.resample(torch.from_numpy(out[f"post{i}"])[None], 24000, 16000)[0]
        out[f"rs{i}"] = y.numpy().copy()
    # log-mel (transformers) on the 16 kHz version of clip 0 and clip 2
    from transformers import WhisperFeatureExtractor
    for nm in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=nm)
        for i in (0, 2):
            w = out[f"rs{i}"]
            pad = fe(w, sampling_rate=16000, return_tensors="np")["input_features"][0]
            nop = fe(w, sampling_rate=16000, return_tensors="np", padding="longest", truncation=False)["input_features"][0]
            t_keep = w.size // 160 + 8
            out[f"mel{nm}_{i}_pad_head"] = pad[:, :t_keep].astype(np.float32).copy()
            tail = pad[:, t_keep:]
            assert np.all(tail == tail[0, 0]), "padding frames are expected to be one constant"
            out[f"mel{nm}_{i}_pad_fill"] = np.float32(tail[0, 0])
            out[f"mel{nm}_{i}_nopad"] = nop.astype(np.float32).copy()
    # cosine
    rng = np.random.default_rng(4321)
    e = np.maximum(rng.normal(size=(17, 256)), 0).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    out["emb"] = e
    refe = e[0]
    out["cos"] = np.asarray([np.dot(refe, g) / (np.linalg.norm(refe) * np.linalg.norm(g)) for g in e[1:]], dtype=np.float32)
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
