"""CPU, authoring container: the oracle against the LIVE reference / torchaudio / transformers.
Skipped on boxes where /root/reference is absent (the committed golden vectors cover those)."""
import numpy as np
import pytest
import torch

import oracle
from tests.util import TOL, assert_close, tone_clip


def _ref_obj(BaseTTS, sr=24000, **over):
    class Ref:
        def __init__(self):
            self.device = "cpu"; self.silence_threshold_db = -50.0; self.crossfade_duration_sec = 0.05
            self.trim_silence = True; self.fade_duration_sec = 0.02; self.force_sentence_split = True
            self.inter_sentence_pause_sec = 0.1; self.sound_decay_threshold = 0.3
            for k, v in over.items():
                setattr(self, k, v)

        @property
        def sample_rate(self):
            return sr
    for n in ("_trim_silence", "_remove_dc_offset", "_apply_fades", "_smooth_segment_join", "_validate_sound_decay"):
        setattr(Ref, n, getattr(BaseTTS, n))
    return Ref()


@pytest.mark.parametrize("sr", [24000, 16000, 22050, 44100])
def test_frame_mean_square_is_bit_identical_to_avg_pool(sr):
    """The sequential fp32 chain of oracle.frame_energy IS torch's CPU avg_pool1d (bit for bit on the
    pooled mean-square; torch's vectorised CPU sqrt may differ from IEEE sqrt by 1 ulp)."""
    rng = np.random.default_rng(sr)
    c = oracle.derive_constants(sr=sr)
    for L in (1, c.hop, c.window + 1, 1000, 24001, 71999):
        x = rng.normal(0, 0.1, L).astype(np.float32)
        pooled = torch.nn.functional.avg_pool1d(torch.from_numpy(x)[None] ** 2, kernel_size=c.window,
                                                stride=c.window // 2, padding=c.window // 2)[0].numpy()
        e = oracle.frame_energy(x, c)
        assert e.shape == pooled.shape
        assert np.array_equal(e.view(np.uint32), np.sqrt(pooled, dtype=np.float32).view(np.uint32))


@pytest.mark.parametrize("sr", [24000, 16000])
def test_trim_and_join_match_reference(reference_basetts, sr):
    rng = np.random.default_rng(1 + sr)
    ref = _ref_obj(reference_basetts, sr)
    c = oracle.derive_constants(sr=sr)
    for L in (1, 100, c.window, 1000, sr, 3 * sr + 1):
        for fs in (True, False):
            for fe in (True, False):
                x = tone_clip(rng, L, min(L // 4, 3000), min(L // 5, 5000), sr=sr)
                r = ref._trim_silence(torch.from_numpy(x.copy()), fs, fe)
                y, tr = oracle.trim_silence(x, c, fs, fe)
                assert r.numel() == y.size and (r.dim() == 2) == tr.all_silent
                assert np.array_equal(r.numpy().reshape(-1), y)
    n_fb = 0
    for _ in range(150):
        segs = []
        for _ in range(int(rng.integers(1, 6))):
            L = int(rng.choice([0, 5, 200, 600, 1500, 5000, 30000]))
            segs.append(rng.normal(0, 1e-4, L).astype(np.float32) if rng.integers(0, 4) == 0
                        else tone_clip(rng, L, min(L // 4, 2000), min(L // 5, 2000), sr=sr))
        r = ref._smooth_segment_join([torch.from_numpy(s.copy()) for s in segs])
        o = oracle.smooth_segment_join(segs, c)
        assert r.numel() == o.audio.size and (r.dim() == 2) == o.two_d, [s.size for s in segs]
        assert_close(o.audio, r.numpy().reshape(-1), tol=1e-6, what="join")
        rr, ok = ref._validate_sound_decay(r)
        r2, ok2, fr, _ = oracle.sound_decay(o.audio)
        assert ok == ok2 and abs(rr - r2) <= 1e-5 * max(1, abs(rr))
        n_fb += o.fallback
    assert n_fb > 10


def test_join_without_pause_and_custom_threshold(reference_basetts):
    rng = np.random.default_rng(9)
    ref = _ref_obj(reference_basetts, inter_sentence_pause_sec=0.0, sound_decay_threshold=0.8, silence_threshold_db=-40.0)
    c = oracle.derive_constants(pause_sec=0.0, silence_db=-40.0)
    segs = [tone_clip(rng, 20000, 900, 1500) for _ in range(4)]
    r = ref._smooth_segment_join([torch.from_numpy(s.copy()) for s in segs])
    o = oracle.smooth_segment_join(segs, c)
    assert r.numel() == o.audio.size
    assert_close(o.audio, r.numpy(), tol=1e-6, what="join no pause")
    x = tone_clip(rng, 48000, decay_to=0.6)
    assert ref._validate_sound_decay(torch.from_numpy(x))[1] == oracle.sound_decay(x, 0.8)[1]


@pytest.mark.parametrize("xfade_sec", [0.0, 0.0004, 0.001])
def test_join_with_crossfade_disabled_matches_reference(reference_basetts, xfade_sec):
    """crossfade_samples == 0: `current_segment[..., :-0]` (base_tts.py:485) is empty, the reference drops segment 0;
    crossfade_samples <= 10: no crossfade is made but segment 0 still loses its tail."""
    rng = np.random.default_rng(31)
    ref = _ref_obj(reference_basetts, crossfade_duration_sec=xfade_sec)
    c = oracle.derive_constants(xfade_sec=xfade_sec)
    for _ in range(40):
        segs = []
        for _ in range(int(rng.integers(2, 5))):
            L = int(rng.choice([0, 5, 20, 600, 5000, 20000]))
            segs.append(rng.normal(0, 1e-4, L).astype(np.float32) if rng.integers(0, 5) == 0
                        else tone_clip(rng, L, min(L // 4, 2000), min(L // 5, 2000)))
        r = ref._smooth_segment_join([torch.from_numpy(s.copy()) for s in segs])
        o = oracle.smooth_segment_join(segs, c)
        assert r.numel() == o.audio.size and (r.dim() == 2) == o.two_d, ([s.size for s in segs], o.fallback)
        assert_close(o.audio, r.numpy().reshape(-1), tol=1e-6, what="join, crossfade disabled")


def test_resample_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(2)
    for L in (1, 2, 3, 4, 5, 7, 1000, 24001, 240000):
        x = rng.normal(0, 0.3, L).astype(np.float32)
        want = ta.functional.resample(torch.from_numpy(x)[None], 24000, 16000)[0].numpy()
        got = oracle.resample(x)
        assert got.shape == want.shape
        assert_close(got, want, tol=1e-6, what=f"resample {L}")


@pytest.mark.parametrize("n_mels", [80, 128])
def test_log_mel_matches_transformers(n_mels):
    tr = pytest.importorskip("transformers")
    from rho_tts_b200 import synth
    fe = tr.WhisperFeatureExtractor(feature_size=n_mels)
    w = oracle.resample(synth.make_clip_block(1, 240000, 5)[0].numpy())
    want = fe(w, sampling_rate=16000, return_tensors="np")["input_features"][0]
    assert_close(oracle.log_mel(w, n_mels, True), want, tol=TOL, what="log-mel 30 s pad")
    want = fe(w, sampling_rate=16000, return_tensors="np", padding="longest", truncation=False)["input_features"][0]
    assert_close(oracle.log_mel(w, n_mels, False), want, tol=TOL, what="log-mel unpadded")


def test_pitch_shift_matches_torchaudio_on_random_cases():
    """oracle/pitch.py against the live torchaudio.functional.pitch_shift (what base_tts.py:644 calls) on random
    lengths and fractional steps; clips <= 1.3 s, where the fp32 reference is reproducible to 1e-4 (tests/test_pitch_shift.py
    explains the long-clip case)."""
    ta = pytest.importorskip("torchaudio")
    from oracle import pitch as OP
    rng = np.random.default_rng(77)
    for case in range(6):
        L = int(rng.integers(257, 31000))
        steps = float(np.round(rng.uniform(-6, 6), 2)) or 1.0
        x = tone_clip(rng, L, amp=0.35) + rng.normal(0, 0.003, L).astype(np.float32)
        want = ta.functional.pitch_shift(torch.from_numpy(x)[None], 24000, steps)[0].numpy()
        got = OP.pitch_shift(x, 24000, steps)
        assert got.shape == want.shape == x.shape
        assert_close(got, want, tol=TOL, what=f"pitch_shift L={L} n_steps={steps}")


def test_pitch_time_steps_match_torch_arange_for_all_semitones():
    """The fp32 time steps of the phase vocoder, bit for bit, for every half-semitone in +-12 and 20 frame counts."""
    from oracle import pitch as OP
    for half in range(-24, 25):
        if half == 0:
            continue
        rate = OP.pitch_rate(half / 2.0)
        for T in range(3, 2400, 123):
            want = torch.arange(0, T, rate, dtype=torch.float32).numpy()
            assert np.array_equal(OP.arange_f32(want.size, rate), want), (half, T)
