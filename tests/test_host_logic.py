"""CPU: host-side logic of the package (layout, sharding, shim behaviour without a GPU)."""
import os
import re

import numpy as np
import pytest
import torch

import rho_tts_b200 as R
from rho_tts_b200 import dist as rdist
from rho_tts_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ragged_layout_alignment():
    lens = [0, 1, 31, 32, 33, 1000, 240000]
    off = R.RaggedBatch.plan_offsets(lens)
    assert off[0] == 0 and np.all(off % R.ALIGN == 0)
    assert np.all(off[1:] >= off[:-1] + np.asarray(lens[:-1]))
    rb = R.RaggedBatch.from_list([torch.arange(n, dtype=torch.float32) for n in lens], "cpu")
    for i, n in enumerate(lens):
        assert torch.equal(rb.clip(i), torch.arange(n, dtype=torch.float32))
    assert rb.max_len == 240000 and rb.total_samples == sum(lens)


def test_from_dense_is_zero_copy():
    x = torch.zeros(4, 64)
    rb = R.RaggedBatch.from_dense(x)
    assert rb.data.data_ptr() == x.data_ptr() and rb.h_offsets.tolist() == [0, 64, 128, 192]
    rb2 = R.RaggedBatch.from_dense(torch.zeros(3, 50))
    assert rb2.h_offsets.tolist() == [0, 64, 128] and rb2.h_lengths.tolist() == [50, 50, 50]


def test_params_mirror_basetts_defaults():
    p = R.make_params()
    assert (p.sr, p.trim_enabled, p.silence_db, p.fade_sec, p.xfade_sec, p.pause_sec, p.decay_thr) == \
        (24000, 1, -50.0, 0.02, 0.05, 0.1, 0.3)

    class T:
        sample_rate = 16000
        trim_silence = False
        silence_threshold_db = -42.5
        fade_duration_sec = 0.01
        crossfade_duration_sec = 0.03
        inter_sentence_pause_sec = 0.0
        sound_decay_threshold = 0.8
    q = R.params_from_tts(T())
    assert (q.sr, q.trim_enabled, q.silence_db, q.fade_sec, q.xfade_sec, q.pause_sec, q.decay_thr) == \
        (16000, 0, -42.5, 0.01, 0.03, 0.0, 0.8)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000, 64000):
        for w in (1, 2, 3, 8):
            spans = [rdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_samples_balanced_and_item_aligned():
    lens = synth.make_ragged_lengths(4000, 3)
    first = synth.make_item_partition(4000, 5)
    for w in (2, 4, 8):
        b = rdist.shard_by_samples(lens, first, w)
        assert b[0] == 0 and b[-1] == len(first) - 1 and np.all(np.diff(b) >= 0)
        pre = np.concatenate([[0], np.cumsum(lens.astype(np.int64))])[first]
        per_rank = np.diff(pre[b])
        assert per_rank.max() / per_rank.mean() < 1.02          # items of <= 6 clips out of 4000


def test_item_partition_and_synth_determinism():
    first = synth.make_item_partition(100, 1)
    assert first[0] == 0 and first[-1] == 100 and np.all(np.diff(first) >= 1) and np.all(np.diff(first)[:-1] >= 2)
    a = synth.make_clip_block(2, 4800, 7)
    b = synth.make_clip_block(2, 4800, 7)
    assert torch.equal(a, b) and a.abs().max() < 1.0


@pytest.mark.skipif(torch.cuda.is_available(), reason="box has a GPU")
def test_shim_raises_runtime_error_without_gpu():
    class Fake(R.B200AudioMixin):
        device = "cpu"
        sample_rate = 24000
        trim_silence = True
        silence_threshold_db = -50.0
        fade_duration_sec = 0.02
        crossfade_duration_sec = 0.05
        inter_sentence_pause_sec = 0.1
        sound_decay_threshold = 0.3
    t = Fake()
    x = torch.randn(24000) * 0.1
    for call in (lambda: t._trim_silence(x), lambda: t._remove_dc_offset(x), lambda: t._apply_fades(x),
                 lambda: t._smooth_segment_join([x, x]), lambda: t._validate_sound_decay(x)):
        with pytest.raises(RuntimeError):            # never ValueError (base_tts.py:786-787)
            call()
    # the early-outs of the reference need no device
    assert t._smooth_segment_join([]) is None
    assert t._validate_sound_decay(torch.tensor([])) == (1.0, True)
    e = torch.tensor([])
    assert t._remove_dc_offset(e) is e and t._apply_fades(e) is e and t._trim_silence(e) is e
    with pytest.raises(RuntimeError):
        R.post_process_batch(R.RaggedBatch.from_dense(torch.zeros(2, 64)), R.make_params())


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under rho_tts_b200/ may import, call or link it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle\.", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rho_tts_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(text), f"{f} references the oracle"


def test_mixin_registers_with_reference_factory(reference_basetts):
    """With the reference importable: the B200 twin is a BaseTTS subclass TTSFactory accepts
    (factory.py:110-122) and its hot-path methods resolve to the mixin."""
    from rho_tts import TTSFactory

    class Dummy(reference_basetts):
        def __init__(self, device="cpu", **kw):
            super().__init__(device=device, **kw)

        def _generate_audio(self, text, **kw):
            return torch.zeros(24000)

        @property
        def sample_rate(self):
            return 24000

    cls = R.make_b200_provider(Dummy)
    TTSFactory.register_provider("dummy_b200", cls)
    tts = TTSFactory.get_tts_instance("dummy_b200")
    for name in ("_trim_silence", "_remove_dc_offset", "_apply_fades", "_smooth_segment_join", "_validate_sound_decay"):
        assert getattr(type(tts), name) is getattr(R.B200AudioMixin, name)
    with pytest.raises(TypeError):
        TTSFactory.register_provider("bad", R.B200AudioMixin)


def test_only_tests_bench_and_smoke_touch_the_oracle():
    """oracle/ is test infrastructure: only tests/, bench.py (its CPU baseline / reference arm) and
    __graft_entry__.py (smoke) may import it -- not the package, not tools/."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(import oracle|from oracle)", re.M)
    offenders = []
    for sub in ("rho_tts_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith(".py") and pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.join(sub, f))
    assert offenders == []


def test_bind_to_gpu_numa_is_harmless_without_nvml():
    """No GPU / NVML here: the helper must leave the affinity alone and say so."""
    import os
    import rho_tts_b200 as R
    before = os.sched_getaffinity(0)
    assert R.dist.bind_to_gpu_numa(0) is None or isinstance(R.dist.bind_to_gpu_numa(0), list)
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_host_item_layout_matches_join_capacity_rule():
    """Output offsets / capacities of the ragged host entry point: sum of the item's segment lengths + max(0, n - 2) pauses
    (SURVEY App. A.5), every offset a multiple of 32 samples."""
    import numpy as np
    import rho_tts_b200 as R
    p = R.make_params()                       # pause 0.1 s -> 2400 samples
    seg = np.array([1000, 24000, 5, 0, 48017, 300, 301], np.int32)
    first = np.array([0, 1, 4, 4, 7], np.int32)      # items of 1, 3, 0 and 3 segments
    off, cap, total = R.host_item_layout(seg, first, p)
    assert cap.tolist() == [1000, 24000 + 5 + 0 + 2400, 0, 48017 + 300 + 301 + 2400]
    assert np.all(off % 32 == 0) and np.all(np.diff(off) >= cap[:-1]) and total >= off[-1] + cap[-1]
    off0, cap0, _ = R.host_item_layout(seg, first, R.make_params(inter_sentence_pause_sec=0.0))
    assert cap0.tolist() == [1000, 24005, 0, 48618]


def test_tensor_validation_sibling_has_no_cpu_fallback():
    """validate_audio_text_match_tensor needs the STT model and a B200: both are loud errors (RuntimeError, never
    ValueError: base_tts.py:786-787), a failing transcription is (True, 0.0, None) like stt_validator.py:251-253."""
    import pytest
    import torch
    import rho_tts_b200 as R
    x = torch.zeros(24000)
    with pytest.raises(RuntimeError, match="STT model"):
        R.validate_audio_text_match_tensor(x, 24000, "hello")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            R.whisper_features(x, 24000)

        class M:
            def generate(self, **kw):
                return [[1]]
        # ... also through the validation entry: only the STT model's own failures are "validation skipped"
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            R.validate_audio_text_match_tensor(x, 24000, "hello", model=M(), tokenizer=object(), similarity_fn=lambda a, b: 1.0)
