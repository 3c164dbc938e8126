"""GPU: the drop-in driven through the REFERENCE's own callers.

`class FakeB200(B200AudioMixin, FakeTTS)` is built exactly like the reference's test double
(/root/reference/tests/test_pipeline.py:14-49) and run through BaseTTS._run_pipeline (base_tts.py:708-956: join ->
hook -> decay at :912-926), stream() (:1132-1184: hook BEFORE trim) and generate(speed, pitch) (:1020-1021), next to the
same class without the mixin (the reference's CPU path): identical segment counts, lengths and decay decisions, audio
within 1e-4.  Needs the reference package: baseline/_ref (installed by __graft_entry__.build()) or /root/reference/src.
"""
import hashlib
import logging
import os
import sys
import threading
import types

import numpy as np
import pytest
import torch

from tests.util import TOL, assert_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 24000


def _import_reference():
    for path in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(path, "rho_tts")):
            if path not in sys.path:
                sys.path.insert(0, path)
            try:
                import rho_tts.base_tts as bt                                   # noqa: F401
                from rho_tts.cancellation import CancellationToken              # noqa: F401
                logging.getLogger("rho_tts").setLevel(logging.CRITICAL)
                return path
            except Exception as e:      # noqa: BLE001
                return f"import failed from {path}: {e!r}"
    return None


REF_PATH = _import_reference()
if REF_PATH is None or REF_PATH.startswith("import failed"):
    pytest.skip(f"SKIP test_reference_callers: the reference package rho_tts is not importable on this box "
                f"(looked in baseline/_ref and /root/reference/src: {REF_PATH})", allow_module_level=True)

from rho_tts.base_tts import BaseTTS                     # noqa: E402
from rho_tts.cancellation import CancellationToken       # noqa: E402


def _segment_audio(text: str, sr: int = SR) -> torch.Tensor:
    """Deterministic stand-in for a TTS model: 0.5 .. 1.3 s of decaying harmonic tone with leading / trailing silence,
    noise and a DC offset, all derived from the text."""
    seed = int(hashlib.sha256(text.encode()).hexdigest()[:8], 16)
    rng = np.random.default_rng(seed)
    L = int(rng.integers(sr // 2, int(1.3 * sr)))
    t = np.arange(L) / sr
    f0 = rng.uniform(100, 280)
    x = 0.3 * np.sin(2 * np.pi * f0 * t) + 0.1 * np.sin(2 * np.pi * 2 * f0 * t)
    x *= np.linspace(1.0, rng.uniform(0.05, 1.1), L) * (0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * t))
    lead, trail = int(rng.integers(0, sr // 8)), int(rng.integers(0, sr // 6))
    x[:lead] = 0
    if trail:
        x[L - trail:] = 0
    x = x + rng.normal(0, 1e-3, L) + rng.uniform(-2e-3, 2e-3)
    return torch.from_numpy(x.astype(np.float32))


class FakeTTS(BaseTTS):
    """The reference's own test double (tests/test_pipeline.py:14-49), with a text-dependent generator."""

    def __init__(self, sr=SR):
        self.device = "cpu"
        self.seed = 42
        self.deterministic = False
        self.phonetic_mapping = {}
        self.silence_threshold_db = -50.0
        self.crossfade_duration_sec = 0.05
        self.trim_silence = True
        self.fade_duration_sec = 0.02
        self.force_sentence_split = True
        self.inter_sentence_pause_sec = 0.1
        self._voice_encoder = None
        self.reference_embedding = None
        self._sample_rate = sr
        self.max_chars_per_segment = 40
        self.max_iterations = 1
        self.accent_drift_threshold = 0.17
        self.text_similarity_threshold = 0.85
        self.sound_decay_threshold = 0.3
        self.max_decay_retries = 2
        self.voice_id = None
        self.drift_model_path = None
        self._max_chars_explicit = True
        self._max_model_chars = 3000

    def _generate_audio(self, text, **kwargs):
        return _segment_audio(text, self._sample_rate)

    @property
    def sample_rate(self):
        return self._sample_rate


def _qwen_hook():
    """QwenTTS._post_process_audio (providers/qwen.py:268-378) with a stub `qwen_tts` module, as the reference's own
    tests import it (tests/test_sound_decay.py:113-121)."""
    if "qwen_tts" not in sys.modules:
        stub = types.ModuleType("qwen_tts")
        stub.Qwen3TTSModel = object
        sys.modules["qwen_tts"] = stub
    from rho_tts.providers.qwen import QwenTTS
    return QwenTTS._post_process_audio


TEXTS = [
    "Hello there. This is the first item of the batch. It has three sentences in it.",
    "One short sentence only.",
    "The quick brown fox jumps over the lazy dog. Pack my box with five dozen liquor jugs. How vexingly quick daft "
    "zebras jump! Sphinx of black quartz, judge my vow.",
]


@pytest.fixture(scope="module")
def classes(cuda_device):
    import rho_tts_b200 as R

    class FakeB200(R.B200AudioMixin, FakeTTS):
        pass

    class FakeQwen(FakeTTS):
        qwen3_sr = SR
        _post_process_audio = _qwen_hook()

    class FakeQwenB200(R.B200QwenAudioMixin, FakeTTS):
        qwen3_sr = SR

    return {"base": (FakeTTS, FakeB200), "qwen": (FakeQwen, FakeQwenB200)}


def _pipeline(tts):
    return tts._run_pipeline(TEXTS, CancellationToken())


def _compare_pipeline(ref_res, got_res):
    assert len(ref_res) == len(got_res) == len(TEXTS)
    n_multi = 0
    for r, g in zip(ref_res, got_res):
        assert (r is None) == (g is None)
        (ra, rn, rm), (ga, gn, gm) = r, g
        assert rn == gn and ra.shape == ga.shape and ga.device.type == "cpu"           # counts, lengths, rank, device
        assert_close(ga.numpy(), ra.numpy(), what="pipeline audio")
        assert abs(rm["decay_ratio"] - gm["decay_ratio"]) <= TOL * max(1.0, abs(rm["decay_ratio"]))
        assert (rm["decay_ratio"] >= 0.3) == (gm["decay_ratio"] >= 0.3)
        n_multi += rn > 1
    assert n_multi >= 2         # the texts really split into several segments: the join / crossfade path is exercised


@pytest.mark.parametrize("kind", ["base", "qwen"])
def test_run_pipeline_matches_reference(classes, kind):
    Ref, B200 = classes[kind]
    _compare_pipeline(_pipeline(Ref()), _pipeline(B200()))


@pytest.mark.parametrize("kind", ["base", "qwen"])
def test_stream_matches_reference(classes, kind):
    """stream(): hook BEFORE trim -> DC -> fades, one result per segment, no join (base_tts.py:1160-1178)."""
    Ref, B200 = classes[kind]
    ref = list(Ref().stream(TEXTS[2]))
    got = list(B200().stream(TEXTS[2]))
    assert len(ref) == len(got) >= 3
    for r, g in zip(ref, got):
        assert r.audio.shape == g.audio.shape and r.duration_sec == g.duration_sec and g.segments_count == 1
        assert_close(g.audio.numpy(), r.audio.numpy(), what="stream audio")


def test_generate_with_speed_and_pitch_matches_reference(classes):
    """generate(speed, pitch): _apply_speed_pitch after the pipeline (base_tts.py:1020-1021).  One-sentence items keep the
    clips short enough (<= 1.3 s) for the fp32 phase vocoder of the reference to be reproducible to 1e-4."""
    Ref, B200 = classes["base"]
    texts = ["One short sentence only.", "Another one, a little longer."]
    for speed, pitch in ((1.1, 0.0), (0.9, 0.0), (1.0, 2.0), (1.25, -3.0)):
        ref = Ref().generate(texts, speed=speed, pitch_semitones=pitch)
        got = B200().generate(texts, speed=speed, pitch_semitones=pitch)
        assert ref is not None and got is not None and len(ref) == len(got) == 2
        for r, g in zip(ref, got):
            assert r.audio.shape == g.audio.shape and r.duration_sec == g.duration_sec
            assert r.segments_count == g.segments_count
            assert abs(r.decay_ratio - g.decay_ratio) <= TOL * max(1.0, abs(r.decay_ratio))
            assert_close(g.audio.numpy(), r.audio.numpy(), what=f"generate speed={speed} pitch={pitch}")
    # speed 1.0 / pitch 0.0 never enters _apply_speed_pitch; the result object carries the pipeline's audio
    r0, g0 = Ref().generate(texts[0]), B200().generate(texts[0])
    assert r0.audio.shape == g0.audio.shape and r0.sample_rate == g0.sample_rate == SR


def test_shared_instance_from_four_threads(classes):
    """One provider instance shared by concurrent sessions (ui/state.py:85-87): four threads run _run_pipeline and
    stream() on the SAME object at once; every result equals the single-threaded one."""
    _, B200 = classes["qwen"]
    tts = B200()
    want = _pipeline(tts)
    want_stream = [r.audio.clone() for r in tts.stream(TEXTS[0])]
    errs, got = [], [None] * 4

    def work(k):
        try:
            for _ in range(3):
                res = _pipeline(tts)
                st = [r.audio for r in tts.stream(TEXTS[0])]
            got[k] = (res, st)
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for res, st in got:
        for (wa, wn, wm), (ga, gn, gm) in zip(want, res):
            assert wn == gn and torch.equal(wa, ga) and wm["decay_ratio"] == gm["decay_ratio"]
        assert len(st) == len(want_stream) and all(torch.equal(a, b) for a, b in zip(want_stream, st))


def test_factory_registration_round_trip(classes):
    """TTSFactory.register_provider / get_tts_instance with a B200 class (factory.py:75-122)."""
    from rho_tts import TTSFactory
    _, B200 = classes["base"]
    TTSFactory.register_provider("fake_b200", B200)
    try:
        assert "fake_b200" in TTSFactory.list_providers() if hasattr(TTSFactory, "list_providers") else True
        tts = TTSFactory.get_tts_instance(provider="fake_b200")
        assert isinstance(tts, B200)
        y = tts._smooth_segment_join([_segment_audio("a"), _segment_audio("b")])
        assert y.dim() == 1 and y.numel() > 0
    finally:
        getattr(TTSFactory, "_providers", {}).pop("fake_b200", None)
