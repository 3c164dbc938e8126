"""GPU: memory-safety checks of our own.  compute-sanitizer is closed on this GPU pool (profiles/sanitizer_r02_closed.log:
"runs under it have left GPUs needing a reset ... find a bad access with bounds checks and asserts of your own, small
cases, and a comparison with the CPU reference"), so the memcheck / initcheck part of SURVEY.md 5.1 is done here:

  * out-of-bounds WRITES: every output buffer is pre-filled with a sentinel bit pattern; after the call every byte outside
    the documented output ranges (beyond an item's out_len inside its slot, the alignment gaps between items, unused
    feature columns) must still hold the sentinel;
  * out-of-bounds READS that matter, and reads of uninitialised memory: the gaps between the input clips are poisoned
    with NaN and with huge values -- if any kernel consumed a sample outside [off, off + len) the outputs would change
    (they must be bit-identical to a run whose gaps are zero);
over the ragged / edge-case inputs of the parity tests (lengths 0, 1, hop +- 1, fade limits, all-silent, > 30 s).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SENT = 0x7FC0DEAD          # a quiet-NaN bit pattern no kernel produces


def _sentinel_fill(t):
    t.view(torch.int32).fill_(SENT) if t.dtype == torch.float32 else t.fill_(0xAB)


def _untouched(t):
    return t.view(torch.int32) == SENT


def _layout(R, lens, dev, poison):
    """Ragged batch whose clips are 96 samples apart from each other's end; the gaps hold `poison`."""
    from rho_tts_b200 import synth
    lens = np.asarray(lens, dtype=np.int32)
    off = np.zeros(len(lens), dtype=np.int64)
    pos = 64
    for i, L in enumerate(lens):
        off[i] = pos
        pos += (int(L) + 31) // 32 * 32 + 96
    data = torch.full((pos + 64,), poison, dtype=torch.float32, device=dev)
    rng = np.random.default_rng(12)
    for i, L in enumerate(lens):
        if L:
            if i % 5 == 4:
                c = torch.from_numpy(rng.normal(0, 1e-4, int(L)).astype(np.float32))          # all silent
            else:
                c = synth.make_clip_block(1, int(L), 800 + i)[0]
            data[off[i]:off[i] + int(L)] = c.to(dev)
    return R.RaggedBatch(data, torch.from_numpy(off).to(dev), torch.from_numpy(lens).to(dev), off, lens)


LENS = [0, 1, 2, 119, 120, 121, 240, 959, 960, 961, 2000, 7680, 24000, 24001, 100003, 184320, 240000, 730001]


@pytest.mark.parametrize("mode", ["one_seg_fused", "one_seg_stages", "joined_fused", "compact_128"])
def test_outputs_stay_inside_their_ranges_and_gaps_are_never_consumed(cuda_device, mode):
    import rho_tts_b200 as R
    p = R.make_params()
    n = len(LENS)
    first = np.arange(n + 1, dtype=np.int32) if mode != "joined_fused" else np.array([0, 1, 4, 9, 12, 15, 17, 18], np.int32)
    n_items = len(first) - 1
    kw = dict(n_mels=128 if mode == "compact_128" else 80, pad_to_30s=True, fuse=(mode != "one_seg_stages"),
              compact=(mode == "compact_128"))
    results = []
    for poison in (0.0, float("nan"), 3.0e38):
        rb = _layout(R, LENS, cuda_device, poison)
        plan = R.ValidatePlan(rb, first, p, kw["n_mels"], kw["pad_to_30s"], kw["fuse"], kw["compact"])
        for t in (plan.out.data, plan.mel_buf, plan.rec):
            _sentinel_fill(t)
        if plan.pad_value is not None:
            _sentinel_fill(plan.pad_value)
        if plan.scratch16 is not None:
            _sentinel_fill(plan.scratch16)
        guard_before = rb.data.clone()
        out = plan.run(rb, None, None)
        torch.cuda.synchronize()
        assert torch.equal(guard_before.view(torch.int32), rb.data.view(torch.int32)), "the input buffer was written"
        rec = out.records_host()
        # y: only [off_i, off_i + out_len_i) of every item may be written
        y = plan.out.data
        keep = torch.ones(y.numel(), dtype=torch.bool, device=cuda_device)
        for i in range(n_items):
            o, L = int(plan.out.h_offsets[i]), int(rec["out_len"][i])
            assert L <= int(plan.out.h_lengths[i])
            keep[o:o + L] = False
            assert not bool(_untouched(y[o:o + L]).any()), f"item {i}: unwritten samples inside its output"
        assert bool(_untouched(y)[keep].all()), f"{mode}: samples outside the items' output ranges were written"
        # features: every row complete up to the row length, finite; records fully written
        mel = plan.mel_buf
        assert not bool(_untouched(mel).any()) and bool(torch.isfinite(mel).all())
        assert not bool((plan.rec == 0xAB).all(dim=1).any())
        if plan.pad_value is not None:
            assert not bool(_untouched(plan.pad_value).any())
        results.append((rec.copy(), y.clone(), mel.clone()))
    for r in results[1:]:               # poisoned gaps change nothing: no sample outside a clip is ever consumed
        assert results[0][0].tobytes() == r[0].tobytes()
        assert torch.equal(results[0][1].view(torch.int32), r[1].view(torch.int32))
        assert torch.equal(results[0][2].view(torch.int32), r[2].view(torch.int32))


def test_join_resample_logmel_qwen_guards(cuda_device):
    """The stand-alone entry points on the same poisoned layout: results independent of the gap contents."""
    import rho_tts_b200 as R
    p = R.make_params()
    first = np.array([0, 3, 4, 9, 12, 15, 17, 18], np.int32)
    res = []
    for poison in (0.0, float("nan")):
        rb = _layout(R, LENS, cuda_device, poison)
        j = R.join_batch(rb, first, p)
        r16 = R.resample_batch(rb)
        lm, nf = R.logmel_batch(r16, 80, True, lengths=r16.lengths)
        q = R.qwen_post_process_batch(rb, 24000)
        sp = R.resample_any_batch(rb, 26400, 24000)
        torch.cuda.synchronize()
        rec = j.records_host()
        ya = torch.cat([j.audio.clip(i, int(rec["out_len"][i])) for i in range(len(first) - 1)])
        r16a = torch.cat([r16.clip(i, -(-2 * int(L) // 3)) for i, L in enumerate(LENS)])
        qa = torch.cat([q.clip(i) for i in range(len(LENS))])
        spa = torch.cat([sp.clip(i) for i in range(len(LENS))])
        res.append((rec.copy(), ya, r16a, lm.clone(), qa, spa))
    assert res[0][0].tobytes() == res[1][0].tobytes()
    for a, b in zip(res[0][1:], res[1][1:]):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
        assert bool(torch.isfinite(a).all())
