"""Speed control of BaseTTS._apply_speed_pitch (base_tts.py:618-650; SURVEY.md 8f NEXT-4, the resample half):
torchaudio.functional.resample(audio, int(sr * speed), sr) for any ratio.

CPU: oracle (oracle/resample.py is ratio-generic) and the library's host tap builder against golden vectors made by
the reference method / torchaudio (tests/golden/make_golden_speed.py).
GPU: rho_b200_resample and the mixin's _apply_speed_pitch against the same vectors and the oracle; the reference's
own speed tests (tests/test_speed_pitch.py:48-99)."""
import ctypes
import math
import os
import sys

import numpy as np
import pytest
import torch

import oracle
from tests.util import assert_close

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from qwen_inputs import keep_index  # noqa: E402
from speed_inputs import SPEEDS, speed_input  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_speed_v1.npz"))
X = speed_input()
SR = 24000


def test_input_regenerates_bit_identically():
    n, s1, s2 = G["in_sum"]
    assert X.size == int(n) and X.astype(np.float64).sum() == s1 and (X.astype(np.float64) ** 2).sum() == s2
    assert list(G["speeds"]) == SPEEDS


@pytest.mark.parametrize("i", range(len(SPEEDS)))
def test_oracle_vs_golden(i):
    y = oracle.resample(X, int(SR * SPEEDS[i]), SR)
    assert y.size == int(G[f"len{i}"])
    assert_close(y[keep_index(y.size)], G[f"out{i}"], tol=1e-5, what=f"speed {SPEEDS[i]}")


def test_host_taps_match_torchaudio():
    """The 11:10 tap table (speed 1.1) of the fp32 recipe equals torchaudio's own kernel to an ulp; the library's
    host builder follows the same recipe and is checked end to end by the GPU tests below."""
    from rho_tts_b200 import _lib
    lib = _lib.load()
    want = G["taps_11_10"]
    taps, width, orig, new = oracle.sinc_resample_kernel(26400, 24000)
    assert (orig, new, width) == (11, 10, int(G["width_11_10"])) and taps.shape == want.shape
    assert float(np.abs(taps - want).max()) <= 1.2e-7
    assert int(lib.rho_b200_resample_out_len(36001, 26400, 24000)) == math.ceil(10 * 36001 / 11)
    assert int(lib.rho_b200_resample_out_len(0, 26400, 24000)) == 0


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(SPEEDS)))
def test_gpu_resample_vs_golden_and_oracle(cuda_device, i):
    import rho_tts_b200 as R
    orig = int(SR * SPEEDS[i])
    rb = R.RaggedBatch.from_list([torch.from_numpy(X), torch.from_numpy(X[:1234]), torch.from_numpy(X[:1])], cuda_device)
    out = R.resample_any_batch(rb, orig, SR)
    lens = out.lengths.cpu().numpy()
    y = out.clip(0, int(lens[0])).cpu().numpy()
    assert y.size == int(G[f"len{i}"])
    assert_close(y[keep_index(y.size)], G[f"out{i}"], what=f"gpu vs golden, speed {SPEEDS[i]}")
    for j, n in enumerate((X.size, 1234, 1)):
        want = oracle.resample(X[:n], orig, SR)
        assert int(lens[j]) == want.size
        assert_close(out.clip(j, want.size).cpu().numpy(), want, what=f"gpu vs oracle, speed {SPEEDS[i]}, n={n}")


@pytest.mark.gpu
def test_gpu_resample_3to2_agrees_with_specialised_kernel(cuda_device):
    import rho_tts_b200 as R
    rb = R.RaggedBatch.from_list([torch.from_numpy(X)], cuda_device)
    a = R.resample_any_batch(rb, 24000, 16000)
    b = R.resample_batch(rb)
    n = int(a.lengths[0])
    assert n == int(b.lengths[0])
    assert float((a.clip(0, n) - b.clip(0, n)).abs().max()) <= 2e-6


@pytest.mark.gpu
class TestSpeedOnMixin:
    """tests/test_speed_pitch.py:48-99 of the reference, on the mixin's _apply_speed_pitch."""

    @staticmethod
    def _tts(sr=16000):
        import rho_tts_b200 as R

        class T(R.B200AudioMixin):
            device = "cpu"
            sample_rate = sr
        return T()

    def test_speed_2x_halves_duration(self, cuda_device):
        t = torch.linspace(0, 1, 16000)
        x = torch.sin(2 * 3.14159 * 440 * t)
        y = self._tts()._apply_speed_pitch(x, 2.0, 0.0)
        assert 0.3 < y.numel() / x.numel() < 0.7 and y.dim() == 1 and y.device.type == "cpu"

    def test_speed_05x_doubles_duration(self, cuda_device):
        t = torch.linspace(0, 1, 16000)
        x = torch.sin(2 * 3.14159 * 440 * t)
        y = self._tts()._apply_speed_pitch(x, 0.5, 0.0)
        assert 1.5 < y.numel() / x.numel() < 2.5

    def test_apply_speed_pitch_noop(self, cuda_device):
        audio = torch.randn(16000)
        result = self._tts()._apply_speed_pitch(audio, speed=1.0, pitch_semitones=0.0)
        assert result is audio                                     # the reference returns the input object (:95-99)

    def test_golden_through_the_method(self, cuda_device):
        tts = self._tts(24000)
        for i in (2, 5):
            y = tts._apply_speed_pitch(torch.from_numpy(X.copy()), SPEEDS[i], 0.0).numpy()
            assert y.size == int(G[f"len{i}"])
            assert_close(y[keep_index(y.size)], G[f"out{i}"], what=f"mixin speed {SPEEDS[i]}")
        y2 = tts._apply_speed_pitch(torch.from_numpy(X.copy()).unsqueeze(0), 1.1, 0.0)
        assert y2.dim() == 1                                        # (1, L) is squeezed (:636-637)

