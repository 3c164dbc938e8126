"""GPU: the drop-in boundary itself.  `B200AudioMixin` is driven exactly as the reference drives BaseTTS:
the cases of the reference's own tests (tests/test_audio_processing.py:33-142, tests/test_sound_decay.py:49-101,
same inputs, same assertions), then the committed golden vectors (outputs of the reference's methods,
tests/golden/make_golden.py) through the single-clip methods, including their aliasing contract
(trim returns a view, DC removal a new tensor, fades mutate in place)."""
import os

import numpy as np
import pytest
import torch

import oracle
from tests.util import assert_close

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
N_CLIPS, N_ITEMS = int(G["n_clips"]), int(G["n_items"])


@pytest.fixture(scope="module")
def TTS(cuda_device):
    import rho_tts_b200 as R

    class ConcreteTTS(R.B200AudioMixin):
        """The reference's stand-in (test_audio_processing.py:7-22, test_sound_decay.py:11-45) with the
        audio methods supplied by the mixin instead of being borrowed from BaseTTS."""

        def __init__(self, sr=16000, device="cpu"):
            self.device = device
            self.silence_threshold_db = -50.0
            self.crossfade_duration_sec = 0.05
            self.trim_silence = True
            self.fade_duration_sec = 0.02
            self.force_sentence_split = True
            self.inter_sentence_pause_sec = 0.1
            self.sound_decay_threshold = 0.3
            self._sample_rate = sr

        @property
        def sample_rate(self):
            return self._sample_rate

    return ConcreteTTS


# ------------------------------------------------------------------ reference: test_audio_processing.py
class TestDCOffsetRemoval:
    def test_removes_offset(self, TTS):
        audio = torch.randn(16000) + 0.5
        result = TTS()._remove_dc_offset(audio)
        assert abs(result.mean().item()) < 0.01

    def test_empty_audio(self, TTS):
        assert TTS()._remove_dc_offset(torch.tensor([])).numel() == 0

    def test_zero_audio_unchanged(self, TTS):
        audio = torch.zeros(100)
        assert torch.allclose(TTS()._remove_dc_offset(audio), audio)


class TestFades:
    def test_fade_in_starts_at_zero(self, TTS):
        result = TTS()._apply_fades(torch.ones(16000), fade_in=True, fade_out=False)
        assert abs(result[0].item()) < 0.01

    def test_fade_out_ends_at_zero(self, TTS):
        result = TTS()._apply_fades(torch.ones(16000), fade_in=False, fade_out=True)
        assert abs(result[-1].item()) < 0.01

    def test_no_fade_unchanged(self, TTS):
        audio = torch.ones(16000)
        assert torch.allclose(TTS()._apply_fades(audio, fade_in=False, fade_out=False), audio)

    def test_short_audio_not_faded(self, TTS):
        short_audio = torch.ones(10)
        assert torch.allclose(TTS()._apply_fades(short_audio, fade_in=True, fade_out=True), short_audio)

    def test_empty_audio(self, TTS):
        assert TTS()._apply_fades(torch.tensor([]), fade_in=True, fade_out=True).numel() == 0


class TestSilenceTrimming:
    def test_trims_leading_silence(self, TTS):
        audio = torch.cat([torch.zeros(16000), torch.randn(8000) * 0.5])
        result = TTS()._trim_silence(audio, from_start=True, from_end=False)
        assert result.shape[-1] < audio.shape[-1]

    def test_trims_trailing_silence(self, TTS):
        audio = torch.cat([torch.randn(8000) * 0.5, torch.zeros(16000)])
        result = TTS()._trim_silence(audio, from_start=False, from_end=True)
        assert result.shape[-1] < audio.shape[-1]

    def test_disabled_trimming(self, TTS):
        tts = TTS()
        tts.trim_silence = False
        audio = torch.zeros(16000)
        assert tts._trim_silence(audio, from_start=True, from_end=True).shape == audio.shape

    def test_all_silent_audio(self, TTS):
        assert TTS()._trim_silence(torch.zeros(16000), from_start=True, from_end=True).numel() > 0


class TestSegmentJoining:
    def test_single_segment(self, TTS):
        result = TTS()._smooth_segment_join([torch.randn(16000) * 0.3])
        assert result is not None and result.numel() > 0

    def test_two_segments(self, TTS):
        result = TTS()._smooth_segment_join([torch.randn(16000) * 0.3, torch.randn(16000) * 0.3])
        assert result is not None and result.shape[-1] > 16000

    def test_empty_list_returns_none(self, TTS):
        assert TTS()._smooth_segment_join([]) is None


# ------------------------------------------------------------------ reference: test_sound_decay.py:49-101
class TestValidateSoundDecay:
    @staticmethod
    def _tone(env_to):
        sr = 24000
        t = torch.linspace(0, 3, 3 * sr)
        return torch.sin(2 * 3.14159 * 440 * t) * torch.linspace(1.0, env_to, 3 * sr)

    def test_constant_volume_passes(self, TTS):
        t = torch.linspace(0, 3, 3 * 24000)
        ratio, is_ok = TTS(24000)._validate_sound_decay(torch.sin(2 * 3.14159 * 440 * t) * 0.5)
        assert is_ok and ratio > 0.9

    def test_severe_decay_fails(self, TTS):
        ratio, is_ok = TTS(24000)._validate_sound_decay(self._tone(0.01))
        assert not is_ok and ratio < 0.3

    def test_mild_decay_passes(self, TTS):
        ratio, is_ok = TTS(24000)._validate_sound_decay(self._tone(0.7))
        assert is_ok and ratio > 0.3

    def test_empty_audio_passes(self, TTS):
        ratio, is_ok = TTS(24000)._validate_sound_decay(torch.tensor([]))
        assert is_ok and ratio == 1.0

    def test_silent_audio_passes(self, TTS):
        _, is_ok = TTS(24000)._validate_sound_decay(torch.zeros(24000))
        assert is_ok

    def test_custom_threshold(self, TTS):
        tts = TTS(24000)
        tts.sound_decay_threshold = 0.8
        _, is_ok = tts._validate_sound_decay(self._tone(0.6))
        assert not is_ok

    def test_return_types_match_reference(self, TTS):
        ratio, is_ok = TTS(24000)._validate_sound_decay(self._tone(0.7))
        assert type(ratio) is float and type(is_ok) is bool          # base_tts.py:322-323


# ------------------------------------------------------------------ golden vectors through the single-clip methods
@pytest.mark.parametrize("where", ["cpu", "cuda"])
@pytest.mark.parametrize("i", range(N_CLIPS))
def test_trim_golden_bounds_and_view(TTS, cuda_device, i, where):
    """start / end sample-exact for all four flag combinations; the result aliases the caller's tensor."""
    tts = TTS(24000)
    for tf in range(4):
        audio = torch.from_numpy(G[f"clip{i}"].copy()).to(where)
        r = tts._trim_silence(audio, bool(tf & 1), bool(tf & 2))
        start, end, dim = (int(v) for v in G[f"trim{i}"][tf])
        assert r.device == audio.device and r.dim() == dim
        assert (r.storage_offset(), r.storage_offset() + r.shape[-1]) == (start, end), (i, tf)
        assert r.untyped_storage().data_ptr() == audio.untyped_storage().data_ptr()       # a view (:392)


@pytest.mark.parametrize("i", range(N_CLIPS))
def test_post_process_golden(TTS, cuda_device, i):
    """_smooth_segment_join([clip]) == trim -> DC -> fades of the reference, then its decay verdict."""
    tts = TTS(24000)
    y = tts._smooth_segment_join([torch.from_numpy(G[f"clip{i}"].copy())])
    want = G[f"post{i}"]
    assert y.dim() == int(G[f"post_dim{i}"]) and y.numel() == want.size
    assert_close(y.cpu().numpy().reshape(-1), want, what=f"post {i}")
    ratio, ok = tts._validate_sound_decay(y)
    gr, gok = G[f"decay{i}"]
    assert ok == bool(gok)
    assert abs(ratio - gr) <= 1e-4 * max(1.0, abs(gr))


@pytest.mark.parametrize("k", range(N_ITEMS))
def test_join_golden(TTS, cuda_device, k):
    """Crossfade joins incl. the all-silent fallback and the 2-D result (base_tts.py:435-536)."""
    tts = TTS(24000)
    segs = [torch.from_numpy(G[f"clip{j}"].copy()) for j in G[f"item{k}_idx"]]
    y = tts._smooth_segment_join(segs)
    want = G[f"item{k}"]
    assert y.dim() == int(G[f"item_dim{k}"]) and y.numel() == want.size, k
    assert_close(y.cpu().numpy().reshape(-1), want, what=f"item {k}")
    ratio, ok = tts._validate_sound_decay(y)
    gr, gok = G[f"item_decay{k}"]
    assert ok == bool(gok) and abs(ratio - gr) <= 1e-4 * max(1.0, abs(gr))


def test_dc_returns_new_tensor_and_fades_mutate_in_place(TTS, cuda_device):
    tts = TTS(24000)
    for dev in ("cpu", cuda_device):
        x = torch.from_numpy(G["clip0"].copy()).to(dev)
        keep = x.clone()
        y = tts._remove_dc_offset(x)
        assert y.data_ptr() != x.data_ptr() and torch.equal(x, keep) and y.device == x.device      # :394-399
        assert_close(y.cpu().numpy(), oracle.remove_dc_offset(G["clip0"]), what="dc")
        z = tts._apply_fades(y)
        assert z.data_ptr() == y.data_ptr()                                                          # :401-433
        assert float(z[0].abs()) < 1e-6 and float(z[-1].abs()) < 1e-6
        c = oracle.derive_constants()
        assert_close(z.cpu().numpy(), oracle.apply_fades(oracle.remove_dc_offset(G["clip0"]), c), what="fades")
    two_d = torch.ones(1, 24000, device=cuda_device)
    out = tts._apply_fades(two_d)
    assert out.shape == (1, 24000) and float(out[0, 0]) == 0.0                                       # view(original_shape)


def test_cosine_golden(TTS, cuda_device):
    tts = TTS(24000)
    emb = G["emb"]
    for j in range(16):
        got = tts._b200_cosine(emb[0], emb[j + 1])
        assert type(got) is np.float32                                                               # :344
        assert abs(float(got) - float(G["cos"][j])) <= 1e-4 * max(1.0, abs(float(G["cos"][j])))


def test_errors_are_runtime_errors(TTS, cuda_device):
    """Multi-channel input is refused with RuntimeError, never ValueError (base_tts.py:786-787)."""
    tts = TTS(24000)
    with pytest.raises(RuntimeError):
        tts._trim_silence(torch.zeros(2, 24000))
    with pytest.raises(RuntimeError):
        tts._smooth_segment_join([torch.zeros(2, 24000), torch.zeros(24000)])


def test_attributes_are_read_at_call_time(TTS, cuda_device):
    """silence_threshold_db / fade_duration_sec / trim_silence changes take effect on the next call."""
    tts = TTS(24000)
    x = torch.from_numpy(G["clip0"].copy())
    a = tts._trim_silence(x).shape[-1]
    tts.silence_threshold_db = -20.0
    b = tts._trim_silence(x).shape[-1]
    c = oracle.derive_constants(silence_db=-20.0)
    tr = oracle.trim_bounds(G["clip0"], c, True, True)
    assert b == tr.end - tr.start and b < a
    tts.trim_silence = False
    assert tts._trim_silence(x) is x
