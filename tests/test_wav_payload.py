"""The caller after the path: BaseTTS._save_wav (base_tts.py:652-667) writes 16-bit PCM.  rho_b200_pcm16 makes that
payload on the device -- (clip(x, -1, 1) * 32767) truncated toward zero, the method's in-tree `wave` writer (the branch
it takes when torchaudio.save has no backend) -- so a clip leaves the GPU as 2 bytes per sample.  Integer output:
bit-exact against the bytes the reference method wrote (tests/golden/make_golden_wav.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from wav_inputs import wav_input  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_wav_v1.npz"))


def test_numpy_formula_equals_the_reference_payload():
    x = wav_input()
    assert np.array_equal((np.clip(x, -1.0, 1.0) * 32767).astype(np.int16), G["pcm"])


@pytest.mark.gpu
def test_gpu_pcm16_is_bit_exact(cuda_device):
    import rho_tts_b200 as R
    x = wav_input()
    clips = [x, x[:1], x[3:1004], x[:8], x[5:13]]           # odd offsets / lengths: the scalar and the 128-bit path
    rb = R.RaggedBatch.from_list([torch.from_numpy(c.copy()) for c in clips], cuda_device)
    pcm = R.pcm16_batch(rb).cpu().numpy()
    for i, c in enumerate(clips):
        o = int(rb.h_offsets[i])
        assert np.array_equal(pcm[o:o + c.size], (np.clip(c, -1.0, 1.0) * 32767).astype(np.int16)), i
    assert np.array_equal(pcm[int(rb.h_offsets[0]):int(rb.h_offsets[0]) + x.size], G["pcm"])


@pytest.mark.gpu
def test_gpu_write_wav_equals_the_reference_file(cuda_device, tmp_path):
    import rho_tts_b200 as R
    path = str(tmp_path / "a.wav")
    R.write_wav(path, torch.from_numpy(wav_input()).unsqueeze(0), 24000, device=cuda_device.index or 0)
    assert np.array_equal(np.frombuffer(open(path, "rb").read(), dtype=np.uint8), G["file_bytes"])
