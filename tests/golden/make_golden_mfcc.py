"""Generates tests/golden/golden_mfcc_v1.npz: mean / std of librosa.feature.mfcc(y, sr=16000, n_mfcc=13)
(/root/reference/src/rho_tts/validation/classifier/trainer.py:50-52).  librosa is NOT in this image, so the vectors
come from the pieces of that chain that are: transformers.audio_utils.spectrogram / mel_filter_bank (a port of
librosa's stft / filters.mel / power_to_db that transformers tests against librosa; float64 inside) with librosa
0.10's defaults spelled out, and scipy.fft.dct (what librosa calls).  Authoring container only.

    python tests/golden/make_golden_mfcc.py
"""
import os
import sys

import numpy as np
import scipy
import scipy.fft

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import transformers  # noqa: E402
from transformers import audio_utils as AU  # noqa: E402
from mfcc_inputs import MFCC_LENGTHS, mfcc_input  # noqa: E402


def mel_db(y):
    fb = AU.mel_filter_bank(1025, 128, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    return AU.spectrogram(y, AU.window_function(2048, "hann"), 2048, 512, 2048, power=2.0, center=True,
                          pad_mode="constant", mel_filters=fb, log_mel="dB", reference=1.0, min_value=1e-10, db_range=80.0)


def main():
    out = {"versions": np.array([transformers.__version__, scipy.__version__, np.__version__]),
           "lengths": np.asarray(MFCC_LENGTHS)}
    for i, n in enumerate(MFCC_LENGTHS):
        y = mfcc_input(n, i)
        db = mel_db(y)
        c = scipy.fft.dct(db.astype(np.float64), type=2, norm="ortho", axis=0)[:13]
        out[f"stats{i}"] = np.concatenate([c.mean(axis=1), c.std(axis=1)])
        out[f"db_sub{i}"] = db[::5, ::3].astype(np.float32)
        out[f"mfcc_sub{i}"] = c[:, ::3].astype(np.float32)
        print(i, n, db.shape, out[f"stats{i}"][:2], out[f"stats{i}"][13:15])
    out["filterbank_sub"] = AU.mel_filter_bank(1025, 128, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T[::4, ::7].astype(np.float32)
    path = os.path.join(HERE, "golden_mfcc_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
