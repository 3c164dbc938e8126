"""Generates tests/golden/golden_speaker_v1.npz: resemblyzer's wav_to_mel_spectrogram (audio.py:
librosa.feature.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=40).T) reached from
/root/reference/src/rho_tts/base_tts.py:335-339.  Neither resemblyzer nor librosa is in this image, so the vectors come
from the piece of that chain that is: transformers.audio_utils.spectrogram / mel_filter_bank (a port of librosa's stft /
filters.mel that transformers tests against librosa; float64 inside) with librosa 0.10's defaults spelled out.
Authoring container only.

    python tests/golden/make_golden_speaker.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import transformers  # noqa: E402
from transformers import audio_utils as AU  # noqa: E402
from speaker_inputs import SPEAKER_LENGTHS, speaker_input  # noqa: E402


def mel40(y):
    fb = AU.mel_filter_bank(201, 40, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    return AU.spectrogram(y, AU.window_function(400, "hann"), 400, 160, 400, power=2.0, center=True,
                          pad_mode="constant", mel_filters=fb).T


def main():
    out = {"versions": np.array([transformers.__version__, np.__version__]), "lengths": np.asarray(SPEAKER_LENGTHS)}
    for i, n in enumerate(SPEAKER_LENGTHS):
        if n == 0:
            continue
        y = speaker_input(i)
        m = mel40(y)
        assert m.shape == (1 + n // 160, 40), m.shape
        out[f"mel_sub{i}"] = m[::3].astype(np.float64) if m.shape[0] > 64 else m.astype(np.float64)
        out[f"mel_max{i}"] = np.float64(m.max())
        print(i, n, m.shape, float(m.max()))
    out["filterbank"] = AU.mel_filter_bank(201, 40, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
    path = os.path.join(HERE, "golden_speaker_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
