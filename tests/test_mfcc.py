"""MFCC statistics of the drift classifier's feature vector (validation/classifier/trainer.py:50-52; SURVEY.md 8f NEXT-3):
mean / std over the frames of librosa.feature.mfcc(y, sr=16000, n_mfcc=13).

PARITY UNPINNED against librosa itself (absent from the image): the golden vectors come from transformers' port of
librosa's stft / mel / power_to_db and scipy's DCT (tests/golden/make_golden_mfcc.py), evaluated in float64.
CPU: oracle/mfcc.py (fp32) against them.  GPU: rho_b200_mfcc_stats against the golden vectors and the oracle; tolerance
1e-4 relative to max(1, |value|) on the 26 statistics (values range over +-400)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mfcc as OM
from tests.util import assert_close

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from mfcc_inputs import MFCC_LENGTHS, mfcc_input  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_mfcc_v1.npz"))


def test_tables_vs_golden():
    assert list(G["lengths"]) == MFCC_LENGTHS
    assert float(np.abs(OM.mel_filterbank()[::4, ::7] - G["filterbank_sub"]).max()) <= 1e-8
    d = OM.dct_matrix(13, 128).astype(np.float64)
    assert float(np.abs(d @ d.T - np.eye(13)).max()) <= 1e-6            # orthonormal rows


@pytest.mark.parametrize("i", range(len(MFCC_LENGTHS)))
def test_oracle_vs_golden(i):
    y = mfcc_input(MFCC_LENGTHS[i], i)
    db = OM.mel_db(y)
    assert db.shape == (128, 1 + MFCC_LENGTHS[i] // 512)
    assert_close(db[::5, ::3], G[f"db_sub{i}"], tol=2e-4, what="dB mel spectrogram")
    assert_close(OM.mfcc(y)[:, ::3], G[f"mfcc_sub{i}"], what="mfcc")
    assert_close(OM.mfcc_stats(y), G[f"stats{i}"], what="mfcc mean / std")


def test_silence_and_constant():
    """All-zero input: every band sits at the 1e-10 floor (-100 dB), c0 = -100 sqrt(128), the rest 0, std 0."""
    s = OM.mfcc_stats(np.zeros(5000, np.float32))
    assert abs(s[0] + 100.0 * np.sqrt(128.0)) <= 1e-3 and float(np.abs(s[1:]).max()) <= 1e-3


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_mfcc_vs_golden_and_oracle(cuda_device):
    import rho_tts_b200 as R
    ys = [mfcc_input(n, i) for i, n in enumerate(MFCC_LENGTHS)] + [np.zeros(5000, np.float32)]
    rb = R.RaggedBatch.from_list([torch.from_numpy(y) for y in ys], cuda_device)
    out = R.mfcc_stats_batch(rb).cpu().numpy()
    assert out.shape == (len(ys), 26)
    for i in range(len(MFCC_LENGTHS)):
        assert_close(out[i], G[f"stats{i}"], what=f"gpu vs golden, n = {MFCC_LENGTHS[i]}")
        assert_close(out[i], OM.mfcc_stats(ys[i]), what=f"gpu vs oracle, n = {MFCC_LENGTHS[i]}")
    assert_close(out[-1], OM.mfcc_stats(ys[-1]), what="silence")


@pytest.mark.gpu
def test_gpu_mfcc_of_the_resampled_pipeline_output(cuda_device):
    """The front end as the classifier would use it: post-processed 24 kHz clips -> 16 kHz on the device -> MFCC."""
    import oracle
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    lens = [240000, 100001, 36000]
    clips = [c.numpy() for c in synth.make_clips(lens, 51)]
    rb = R.RaggedBatch.from_list([torch.from_numpy(c) for c in clips], cuda_device)
    rb16 = R.resample_batch(rb)
    out = R.mfcc_stats_batch(rb16, lengths=rb16.lengths).cpu().numpy()
    for i, x in enumerate(clips):
        assert_close(out[i], OM.mfcc_stats(oracle.resample(x)), what=f"clip {i}")


@pytest.mark.gpu
def test_gpu_mfcc_rejects_empty(cuda_device):
    import rho_tts_b200 as R
    rb = R.RaggedBatch.from_list([torch.zeros(0), torch.zeros(100)], cuda_device)
    with pytest.raises(RuntimeError):
        R.mfcc_stats_batch(rb)
