"""Pitch control of BaseTTS._apply_speed_pitch (base_tts.py:639-648; SURVEY.md 8f NEXT-4, the pitch half):
torchaudio.functional.pitch_shift = stft -> phase vocoder -> istft -> resample -> crop / pad.

CPU: oracle/pitch.py against golden vectors made by the reference method / torchaudio
(tests/golden/make_golden_pitch.py), stage by stage (time steps, phase_advance: bit-exact; stft, vocoder, istft) and
end to end.  GPU: rho_b200_pitch_shift and the mixin's _apply_speed_pitch against the same vectors and the oracle.

Tolerance: 1e-4 absolute on waveforms of peak ~0.4 (the task's float tolerance) for clips up to 1.5 s.  The fp32
reference differs from its own float64 evaluation by 1e-4 .. 9e-4 (phase accumulator ~1e6 rad, ulp 0.06), so this only
holds because the oracle and the kernels repeat torch's fp32 roundings; typical errors are 1e-5 .. 4e-5.  On long
tonal clips the accumulated phase is rounded to fp32 at magnitudes where one ulp is 0.016 rad on an audible partial:
last-bit differences of the FFT / atan2 (MKL, Sleef) flip ~1 % of those roundings, and no independent implementation
-- the numpy oracle included -- tracks the fp32 run to 1e-4 there (2.6e-4 on the 10 s case).  That case is bounded
by the reference's own fp32 noise instead (LONG_CASE)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pitch as OP
from tests.util import assert_close

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from pitch_inputs import COMBINED_CASES, LONG_CASE, PITCH_CASES, pitch_input  # noqa: E402
from qwen_inputs import keep_index  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_pitch_v1.npz"))
SR = 24000
TOL = 1e-4


def test_arange_and_linspace_models_are_bit_exact():
    for i, r in enumerate(G["arange_rates"]):
        want = G[f"arange{i}"]
        assert np.array_equal(OP.arange_f32(want.size, float(r)), want)
        assert want.size == math.ceil((1500 + 7 * i) / float(r))
    assert np.array_equal(OP.linspace_f32(math.pi * 128, 257), G["linspace257"])


def test_oracle_stages_vs_torch():
    x = pitch_input(9000, 50)
    rate = float(G["arange_rates"][0])
    spec = OP.stft512(x)
    assert_close(np.abs(spec)[::8, ::5], G["stft_abs_sub"], tol=2e-5, what="stft magnitude")
    mag, pacc = OP.phase_vocoder(spec, rate)
    pv = mag * np.cos(pacc.astype(np.float64)) + 1j * mag * np.sin(pacc.astype(np.float64))
    want = G["pv_sub"]
    err = np.abs(pv[::8, ::5] - want).max() / np.abs(want).max()
    assert err <= 5e-4, err                      # phase noise of atan2 implementations, relative to the largest bin
    w = OP.istft512(pv.astype(np.complex64), int(round(9000 / rate)))
    assert_close(w[::7], G["istft_sub"], tol=5e-5, what="istft")


@pytest.mark.parametrize("i", range(len(PITCH_CASES)))
def test_oracle_vs_golden(i):
    n, steps = PITCH_CASES[i]
    x = pitch_input(n, i)
    y = OP.pitch_shift(x, SR, steps)
    assert y.size == n
    assert_close(y[keep_index(n)], G[f"out{i}"], tol=TOL, what=f"n_steps {steps}")
    s1, s2 = G[f"out_sum{i}"]
    assert abs((y.astype(np.float64) ** 2).sum() - s2) <= 2e-3 * s2


def _long_case_check(y, what):
    """10 s tonal clip: the distance to the fp32 reference is bounded by the reference's own fp32 noise (its distance
    to the float64 run of the same code), and the distance to the float64 run is no worse than the reference's."""
    n = LONG_CASE[0]
    k = keep_index(n)
    noise_max, noise_rms = G["long_noise"]
    d32 = np.abs(y[k].astype(np.float64) - G["long32"])
    d64 = np.abs(y[k].astype(np.float64) - G["long64"])
    ref64 = np.abs(G["long32"].astype(np.float64) - G["long64"])
    assert noise_max > TOL                                        # the premise: the reference is noisier than 1e-4 here
    assert d32.max() <= noise_max, f"{what}: {d32.max():.2e} from the fp32 reference, its own noise is {noise_max:.2e}"
    assert np.sqrt((d32 ** 2).mean()) <= noise_rms
    assert d64.max() <= 1.25 * ref64.max() and np.sqrt((d64 ** 2).mean()) <= 1.25 * np.sqrt((ref64 ** 2).mean())


def test_oracle_long_tonal_clip_within_reference_noise():
    n, steps = LONG_CASE
    _long_case_check(OP.pitch_shift(pitch_input(n, 200), SR, steps), "oracle")


def test_windowed_resample_equals_dense():
    """The taps outside the window do not matter: the windowed evaluation equals oracle.resample (dense taps)."""
    import oracle
    x = pitch_input(5000, 60)
    for orig, new in ((26400, 24000), (20000, 24000), (48000, 24000), (12000, 24000)):
        a = OP.resample_windowed(x, orig, new)
        b = oracle.resample(x, orig, new)
        assert a.size == b.size
        assert_close(a, b, tol=2e-6, what=f"{orig}->{new}")


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_pitch_vs_golden_and_oracle(cuda_device):
    import rho_tts_b200 as R
    by_steps = {}
    for i, (n, steps) in enumerate(PITCH_CASES):
        by_steps.setdefault(steps, []).append(i)
    for steps, idx in by_steps.items():
        xs = [pitch_input(PITCH_CASES[i][0], i) for i in idx]
        rb = R.RaggedBatch.from_list([torch.from_numpy(x) for x in xs], cuda_device)
        out = R.pitch_shift_batch(rb, SR, steps)
        for j, i in enumerate(idx):
            n = PITCH_CASES[i][0]
            y = out.clip(j, n).cpu().numpy()
            assert_close(y[keep_index(n)], G[f"out{i}"], tol=TOL, what=f"gpu vs golden, n_steps {steps}, n {n}")
            assert_close(y, OP.pitch_shift(xs[j], SR, steps), tol=TOL, what=f"gpu vs oracle, n_steps {steps}, n {n}")


@pytest.mark.gpu
def test_gpu_pitch_ragged_batch_vs_oracle(cuda_device):
    """One call, ragged lengths (hop multiples, +-1, the 257-sample minimum): every clip equals its own oracle run."""
    import rho_tts_b200 as R
    lens = [257, 384, 511, 512, 513, 1280, 4095, 9999, 24000]
    xs = [pitch_input(n, 70 + k) for k, n in enumerate(lens)]
    rb = R.RaggedBatch.from_list([torch.from_numpy(x) for x in xs], cuda_device)
    for steps in (3.0, -4.0):
        out = R.pitch_shift_batch(rb, SR, steps)
        for k, n in enumerate(lens):
            assert_close(out.clip(k, n).cpu().numpy(), OP.pitch_shift(xs[k], SR, steps), tol=TOL,
                         what=f"n_steps {steps}, n {n}")


@pytest.mark.gpu
def test_gpu_long_tonal_clip_within_reference_noise(cuda_device):
    import rho_tts_b200 as R
    n, steps = LONG_CASE
    rb = R.RaggedBatch.from_list([torch.from_numpy(pitch_input(n, 200))], cuda_device)
    _long_case_check(R.pitch_shift_batch(rb, SR, steps).clip(0).cpu().numpy(), "gpu")


@pytest.mark.gpu
def test_gpu_pitch_tiny_step_skips_the_resample(cuda_device):
    """int(sr / rate) == sr for |n_steps| < ~7e-4: torchaudio's resample returns its input, so only the vocoder acts."""
    import rho_tts_b200 as R
    x = pitch_input(9000, 90)
    for steps in (3e-4, -2e-4):                       # -2e-4: 23999 -> 24000, a ratio with 24000 phases
        assert int(SR / OP.pitch_rate(steps)) == (SR if steps > 0 else SR - 1)
        rb = R.RaggedBatch.from_list([torch.from_numpy(x)], cuda_device)
        y = R.pitch_shift_batch(rb, SR, steps).clip(0).cpu().numpy()
        assert_close(y, OP.pitch_shift(x, SR, steps), tol=TOL, what=f"n_steps {steps}")


@pytest.mark.gpu
def test_gpu_pitch_errors(cuda_device):
    import rho_tts_b200 as R
    rb = R.RaggedBatch.from_list([torch.from_numpy(pitch_input(256, 1)), torch.from_numpy(pitch_input(4000, 2))], cuda_device)
    with pytest.raises(RuntimeError):          # torch.stft: reflect padding needs more than 256 samples
        R.pitch_shift_batch(rb, SR, 2.0)
    rb2 = R.RaggedBatch.from_list([torch.from_numpy(pitch_input(4000, 2))], cuda_device)
    assert R.pitch_shift_batch(rb2, SR, 0.0) is rb2


@pytest.mark.gpu
class TestPitchOnMixin:
    """tests/test_speed_pitch.py of the reference (the pitch cases), on the mixin's _apply_speed_pitch."""

    @staticmethod
    def _tts(sr=16000):
        import rho_tts_b200 as R

        class T(R.B200AudioMixin):
            device = "cpu"
            sample_rate = sr
        return T()

    def test_pitch_shift_preserves_length(self, cuda_device):
        t = torch.linspace(0, 1, 16000)
        x = torch.sin(2 * 3.14159 * 440 * t)
        y = self._tts()._apply_speed_pitch(x, 1.0, 2.0)
        assert y.shape == x.shape and y.device.type == "cpu"
        y2 = self._tts()._apply_speed_pitch(x.unsqueeze(0), 1.0, -2.0)
        assert y2.shape == (1, 16000)

    def test_pitch_moves_the_spectral_peak(self, cuda_device):
        t = torch.arange(16000) / 16000.0
        x = torch.sin(2 * math.pi * 440 * t)
        y = self._tts()._apply_speed_pitch(x, 1.0, 12.0).numpy()
        peak = np.abs(np.fft.rfft(y[2000:14000] * np.hanning(12000))).argmax() * 16000 / 12000
        assert abs(peak - 880) < 15

    def test_golden_through_the_method(self, cuda_device):
        tts = self._tts(24000)
        for i in (1, 3):
            n, steps = PITCH_CASES[i]
            y = tts._apply_speed_pitch(torch.from_numpy(pitch_input(n, i)), 1.0, steps).numpy()
            assert_close(y[keep_index(n)], G[f"out{i}"], tol=TOL, what=f"mixin n_steps {steps}")

    def test_speed_then_pitch(self, cuda_device):
        tts = self._tts(24000)
        for i, (n, speed, steps) in enumerate(COMBINED_CASES):
            y = tts._apply_speed_pitch(torch.from_numpy(pitch_input(n, 100 + i)), speed, steps).numpy()
            assert y.size == int(G[f"comb_len{i}"])
            assert_close(y[keep_index(y.size)], G[f"comb{i}"], tol=TOL, what=f"speed {speed} + n_steps {steps}")
