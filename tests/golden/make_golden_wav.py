"""Generates tests/golden/golden_wav_v1.npz: the int16 payload BaseTTS._save_wav writes
(/root/reference/src/rho_tts/base_tts.py:652-667) through the REFERENCE method itself.  torchaudio.save has no
backend in this image (torchcodec is absent), so the method takes its in-tree `wave` fallback -- the branch the GPU op
mirrors; the script asserts that this is the branch taken.  Authoring container only.

    python tests/golden/make_golden_wav.py
"""
import os
import sys
import tempfile
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/src")
import torch  # noqa: E402
import torchaudio  # noqa: E402
from rho_tts.base_tts import BaseTTS  # noqa: E402
from wav_inputs import wav_input  # noqa: E402


class Ref:
    pass


Ref._save_wav = BaseTTS._save_wav


def main():
    x = wav_input()
    fallback = False
    try:
        with tempfile.NamedTemporaryFile(suffix=".wav") as f:
            torchaudio.save(f.name, torch.from_numpy(x)[None], 24000)
    except Exception:   # noqa: BLE001
        fallback = True
    assert fallback, "torchaudio.save works here: the reference would not take its wave fallback"
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a.wav")
        Ref()._save_wav(path, torch.from_numpy(x)[None], 24000)
        with wave.open(path, "rb") as wf:
            assert (wf.getnchannels(), wf.getsampwidth(), wf.getframerate()) == (1, 2, 24000)
            pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16).copy()
        raw = open(path, "rb").read()
    assert pcm.size == x.size
    out = {"pcm": pcm, "file_bytes": np.frombuffer(raw, dtype=np.uint8).copy(),
           "versions": np.array([torch.__version__, torchaudio.__version__])}
    p = os.path.join(HERE, "golden_wav_v1.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p) // 1024, "KiB;", pcm.size, "samples, min/max", pcm.min(), pcm.max())


if __name__ == "__main__":
    main()
