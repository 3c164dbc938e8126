"""QwenTTS._post_process_audio (providers/qwen.py:268-378; SURVEY.md 8f NEXT-1).

CPU: the numpy oracle against golden vectors made by the reference's own method (tests/golden/make_golden_qwen.py),
and live against the reference when /root/reference is present.
GPU: rho_b200_qwen_postprocess and the B200QwenAudioMixin hook against the golden vectors and the oracle, plus the
reference's own two tests (tests/test_sound_decay.py:104-179) run on the mixin."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import qwen as oq
from tests.conftest import have_reference
from tests.util import assert_close

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from qwen_inputs import make_inputs, keep_index  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_qwen_v1.npz"))
N = int(G["n_clips"])
CLIPS = make_inputs()


def test_inputs_regenerate_bit_identically():
    assert len(CLIPS) == N
    for i, x in enumerate(CLIPS):
        n, s1, s2 = G[f"in_sum{i}"]
        assert x.size == int(n) and x.astype(np.float64).sum() == s1 and (x.astype(np.float64) ** 2).sum() == s2


@pytest.mark.parametrize("i", range(N))
def test_oracle_vs_golden(i):
    y = oq.post_process(CLIPS[i])
    assert_close(y[keep_index(y.size)], G[f"out{i}"], tol=1e-6, what=f"qwen clip {i}")
    s1, s2 = G[f"out_sum{i}"]
    assert abs(float((y.astype(np.float64) ** 2).sum()) - s2) <= 1e-5 * max(1.0, s2)


def test_oracle_covers_every_branch():
    unchanged = [np.array_equal(oq.post_process(x), x) for x in CLIPS]
    assert unchanged[6] and unchanged[7] and not unchanged[0]                         # the 1e-8 RMS gate
    w = 48000
    applied = [x.size > 2 * w and oq.windowed_gains(x, w)[0] for x in CLIPS]
    assert applied[0] and applied[3] and applied[4] and applied[5] and applied[9]
    assert not applied[1] and not applied[2] and not applied[8]                        # flat / short / silent first window
    assert max(oq.windowed_gains(CLIPS[5], w)[1]) > 7.0                                # the +18 dB cap (7.94) region
    y16 = oq.post_process(CLIPS[0], 16000)
    assert_close(y16[keep_index(y16.size)], G["out0_sr16k"], tol=1e-6, what="sr 16000")


@pytest.mark.skipif(not have_reference(), reason="/root/reference not present on this box")
def test_oracle_vs_live_reference():
    from unittest.mock import MagicMock
    sys.path.insert(0, "/root/reference/src")
    sys.modules.setdefault("qwen_tts", MagicMock())
    from rho_tts.providers.qwen import QwenTTS
    tts = QwenTTS.__new__(QwenTTS)
    tts.qwen3_sr = 24000
    tts.device = "cpu"
    rng = np.random.default_rng(3)
    for k in range(12):
        n = int(rng.integers(1000, 400000))
        x = (rng.normal(0, 0.1, n) * np.linspace(1.0, rng.uniform(0.05, 1.5), n)).astype(np.float32)
        want = tts._post_process_audio(torch.from_numpy(x.copy())).numpy()
        assert_close(oq.post_process(x), want, tol=1e-6, what=f"live {k}")


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_batch_vs_golden_and_oracle(cuda_device):
    import rho_tts_b200 as R
    rb = R.RaggedBatch.from_list([torch.from_numpy(x) for x in CLIPS], cuda_device)
    out = R.qwen_post_process_batch(rb, 24000)
    assert rb.data.data_ptr() != out.data.data_ptr()
    for i, x in enumerate(CLIPS):
        y = out.clip(i).cpu().numpy()
        assert_close(y[keep_index(y.size)], G[f"out{i}"], what=f"gpu vs golden {i}")
        assert_close(y, oq.post_process(x), what=f"gpu vs oracle {i}")
        if i in (6, 7):
            assert np.array_equal(y, x)                                                 # gated clips are copied bit for bit
    # in place, and a different sample rate (window = 2 * sr)
    out16 = R.qwen_post_process_batch(rb, 16000, in_place=True)
    assert out16.data.data_ptr() == rb.data.data_ptr()
    y16 = out16.clip(0).cpu().numpy()
    assert_close(y16[keep_index(y16.size)], G["out0_sr16k"], what="gpu sr 16000")


@pytest.mark.gpu
def test_gpu_random_ragged_vs_oracle(cuda_device):
    import rho_tts_b200 as R
    rng = np.random.default_rng(11)
    clips = []
    for k in range(40):
        n = int(rng.integers(1, 500000))
        env = np.linspace(1.0, rng.uniform(0.02, 2.0), n)
        clips.append((rng.normal(0, rng.uniform(0.01, 0.5), n) * env).astype(np.float32))
    rb = R.RaggedBatch.from_list([torch.from_numpy(x) for x in clips], cuda_device)
    out = R.qwen_post_process_batch(rb, 24000)
    for i, x in enumerate(clips):
        assert_close(out.clip(i).cpu().numpy(), oq.post_process(x), what=f"ragged {i} (n={x.size})")


@pytest.mark.gpu
def test_gpu_lengths_from_records_and_full_size(cuda_device):
    """C2-sized batch: the hook runs on the joined audio with the lengths taken from the records (device,
    48-byte stride); properties that hold without the oracle, and three clips against it."""
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    x = synth.make_clip_block(1000, 240000, 0xB200, device=cuda_device)
    rb = R.RaggedBatch.from_dense(x)
    post = R.post_process_batch(rb, R.make_params())
    rec = post.records_host()
    out = R.qwen_post_process_batch(post.audio, 24000, lengths=post.records[:, 8:12].contiguous().view(torch.int32),
                                    len_stride=4)
    for i in (0, 499, 999):
        L = int(rec["out_len"][i])
        assert_close(out.clip(i, L).cpu().numpy(), oq.post_process(post.audio.clip(i, L).cpu().numpy()),
                     what=f"full-size clip {i}")
    assert float(out.data.abs().max()) <= 0.95                                          # tanh soft clip bound
    rms = torch.stack([out.clip(i, int(rec["out_len"][i])).double().pow(2).mean().sqrt() for i in range(0, 1000, 37)])
    assert float((20 * torch.log10(rms)).max()) <= -22.9                               # -23 dBFS before the soft clip


@pytest.mark.gpu
class TestWindowedNormalizationOnMixin:
    """The reference's tests/test_sound_decay.py:104-179, with the hook supplied by B200QwenAudioMixin."""

    @staticmethod
    def _tts():
        import rho_tts_b200 as R

        class T(R.B200QwenAudioMixin):
            qwen3_sr = 24000
            device = "cpu"
            sample_rate = 24000
        return T()

    def test_corrects_decaying_audio(self, cuda_device):
        sr, duration = 24000, 10
        n = sr * duration
        t = torch.linspace(0, duration, n)
        audio = torch.sin(2 * 3.14159 * 440 * t) * torch.linspace(1.0, 0.2, n)
        third = n // 3
        before = (audio[-third:].pow(2).mean().sqrt() / audio[:third].pow(2).mean().sqrt()).item()
        result = self._tts()._post_process_audio(audio.clone())
        after = (result[-third:].pow(2).mean().sqrt() / result[:third].pow(2).mean().sqrt()).item()
        assert before < 0.4 and after > 0.6

    def test_does_not_alter_constant_audio(self, cuda_device):
        sr, duration = 24000, 6
        n = sr * duration
        t = torch.linspace(0, duration, n)
        result = self._tts()._post_process_audio((torch.sin(2 * 3.14159 * 440 * t) * 0.5).clone())
        third = n // 3
        ratio = (result[-third:].pow(2).mean().sqrt() / result[:third].pow(2).mean().sqrt()).item()
        assert 0.85 < ratio < 1.15

    def test_shapes_and_golden(self, cuda_device):
        tts = self._tts()
        y = tts._post_process_audio(torch.from_numpy(CLIPS[0].copy()).unsqueeze(0))      # (1, L) stays (1, L)
        assert tuple(y.shape) == (1, CLIPS[0].size) and y.device.type == "cpu"
        yn = y.numpy()[0]
        assert_close(yn[keep_index(yn.size)], G["out0"], what="mixin vs golden")
        assert tts._post_process_audio(torch.tensor([])).numel() == 0
        tts.qwen3_sr = 16000                                                             # read at call time
        y16 = tts._post_process_audio(torch.from_numpy(CLIPS[0].copy())).numpy()
        assert_close(y16[keep_index(y16.size)], G["out0_sr16k"], what="mixin sr 16000")


@pytest.mark.gpu
def test_gpu_qwen_pipeline_join_hook_decay(cuda_device):
    """The Qwen provider's order of operations (base_tts.py:911-926): join -> loudness hook -> decay check on the
    HOOKED audio, batched, against the oracle chain; the accept / reject decision must be the same."""
    import oracle
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    lens = synth.make_ragged_lengths(24, 9, 1.0, 7.0)
    clips = [c.numpy() for c in synth.make_clips(lens, 21)]
    first = synth.make_item_partition(24, 5, 1, 4)
    p = R.make_params()
    rb = R.RaggedBatch.from_list([torch.from_numpy(c) for c in clips], cuda_device)
    out = R.qwen_pipeline_batch(rb, first, p, 24000)
    rec = out.records_host()
    c = oracle.derive_constants()
    flips = 0
    for i in range(len(first) - 1):
        segs = clips[first[i]:first[i + 1]]
        j = oracle.smooth_segment_join(segs, c).audio
        hooked = oq.post_process(j)
        ratio, ok, fr, lr = oracle.sound_decay(hooked, 0.3)
        L = int(rec["out_len"][i])
        assert L == hooked.size
        assert_close(out.audio.clip(i, L).cpu().numpy(), hooked, what=f"item {i}")
        assert bool(rec["ok"][i]) == ok and abs(rec["decay_ratio"][i] - ratio) <= 1e-4 * max(1.0, abs(ratio))
        flips += int(abs(ratio - oracle.sound_decay(j, 0.3)[0]) > 1e-3 * max(1.0, abs(ratio)))
    assert flips > 0            # the ratio is the hooked audio's, not the joined audio's: the order of operations matters


@pytest.mark.gpu
def test_gpu_qwen_validate_features_of_hooked_audio(cuda_device):
    """qwen_validate_batch: the log-mel features and the cosine belong to the audio AFTER the loudness hook."""
    import oracle
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    lens = synth.make_ragged_lengths(10, 31, 1.5, 6.0)
    clips = [c.numpy() for c in synth.make_clips(lens, 33)]
    first = synth.make_item_partition(10, 7, 1, 3)
    n_items = len(first) - 1
    emb, ref = synth.make_embeddings(n_items, device=cuda_device)
    p = R.make_params()
    rb = R.RaggedBatch.from_list([torch.from_numpy(c) for c in clips], cuda_device)
    v = R.qwen_validate_batch(rb, first, p, emb, ref, n_mels=80, pad_to_30s=False)
    rec = v.records_host()
    c = oracle.derive_constants()
    for i in range(n_items):
        j = oracle.smooth_segment_join(clips[first[i]:first[i + 1]], c).audio
        hooked = oq.post_process(j)
        L = int(rec["out_len"][i])
        got_audio = v.audio.clip(i, L).cpu().numpy()
        assert_close(got_audio, hooked, what=f"hooked item {i}")
        # stage parity on identical inputs: the oracle's features of the GPU's hooked audio.  End to end the bound is
        # looser: a 1e-7 relative difference of the waveform is a noise floor at -140 dB, i.e. 1e-3 relative on bins just
        # above the normaliser's -80 dB clamp, 4e-4 in log10 units.
        want = oracle.log_mel(oracle.resample(got_audio), 80, False)
        T = want.shape[1]
        assert_close(v.mel[i, :, :T].cpu().numpy(), want, what=f"log-mel of the hooked item {i}")
        assert_close(v.mel[i, :, :T].cpu().numpy(), oracle.log_mel(oracle.resample(hooked), 80, False), tol=4e-4,
                     what=f"log-mel end to end, item {i}")
        assert abs(rec["cosine"][i] - oracle.cosine_similarity(ref.cpu().numpy(), emb[i].cpu().numpy())) <= 1e-5
        assert bool(rec["ok"][i]) == oracle.sound_decay(hooked, 0.3)[1]
