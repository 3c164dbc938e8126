"""GPU parity: librho_b200 (through the C ABI / ctypes shim) against the CPU oracle.

Bar (BASELINE.json north_star): trim bounds, join lengths / segment boundaries and accept/reject
decisions exact; waveforms, RMS, ratio, cosine and log-mel within 1e-4 (abs-or-rel, tests/util.TOL).
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import dsp as odsp
from tests.util import TOL, assert_close, tone_clip

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R():
    import rho_tts_b200 as r
    return r


def _records_match(a, b):
    """Records of two runs of the same items: integers, flags and the cosine bit for bit; the decay sums are fp32 partial
    sums whose grouping follows the tile schedule (which depends on the batch), so RMS / ratio agree to rounding only."""
    for f in ("start", "end", "out_len", "flags", "ok", "n_segments", "cosine", "dc"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("first_rms", "last_rms", "decay_ratio"):
        assert np.allclose(a[f], b[f], rtol=1e-6, atol=0), f


def _rb(R, clips, dev):
    return R.RaggedBatch.from_list([torch.from_numpy(np.ascontiguousarray(c)) for c in clips], dev)


# ----------------------------------------------------------------------------- trim
@pytest.mark.parametrize("sr", [24000, 16000, 22050, 44100])
def test_trim_scan_exact(R, cuda_device, sr):
    rng = np.random.default_rng(7 + sr)
    c = oracle.derive_constants(sr=sr)
    lens = [1, 5, c.hop - 1, c.hop, c.hop + 1, c.window - 1, c.window, c.window + 1, 1000, 4097, sr, sr + 1,
            3 * sr + 77, 10 * sr]
    clips, flags = [], []
    for i, L in enumerate(lens):
        for tf in (3, 1, 2, 0):
            kind = (i + tf) % 5
            if kind == 0:
                x = rng.normal(0, 1e-4, L).astype(np.float32)                 # all silent
            elif kind == 1:
                x = rng.normal(0, 0.2, L).astype(np.float32)                  # loud everywhere
            else:
                x = tone_clip(rng, L, lead=min(L // 4, int(rng.integers(0, sr // 2))),
                              trail=min(L // 5, int(rng.integers(0, sr // 2))), sr=sr)
            clips.append(x); flags.append(tf)
    rb = _rb(R, clips, cuda_device)
    p = R.make_params(sr=sr)
    info = R.trim_scan_batch(rb, p, torch.tensor(flags, dtype=torch.uint8, device=cuda_device))
    info = info.cpu().numpy().view(R.SEG_DTYPE).reshape(-1)
    for i, (x, tf) in enumerate(zip(clips, flags)):
        tr = odsp.trim_bounds(x, c, bool(tf & 1), bool(tf & 2))
        assert (int(info["start"][i]), int(info["end"][i])) == (tr.start, tr.end), (i, len(x), tf, info[i], tr)
        assert bool(info["flags"][i] & 1) == tr.all_silent, (i, len(x), tf)
        seg = x[tr.start:tr.end]
        if seg.size:
            assert abs(float(info["dc"][i]) - float(seg.mean(dtype=np.float64))) <= 1e-6


def test_trim_disabled_and_empty(R, cuda_device):
    rng = np.random.default_rng(3)
    clips = [tone_clip(rng, 5000, 1000, 1000), np.zeros(0, np.float32), tone_clip(rng, 300)]
    rb = _rb(R, clips, cuda_device)
    p = R.make_params(trim_silence=False)
    info = R.trim_scan_batch(rb, p).cpu().numpy().view(R.SEG_DTYPE).reshape(-1)
    assert [(int(a), int(b)) for a, b in zip(info["start"], info["end"])] == [(0, 5000), (0, 0), (0, 300)]
    assert all(info["flags"] & 8)


# ----------------------------------------------------------------------------- post-process
def test_post_process_batch_vs_oracle(R, cuda_device):
    from rho_tts_b200 import synth
    x = synth.make_clip_block(48, 240000, 1234 + 2)                # CPU generator: bit-identical inputs
    rb = R.RaggedBatch.from_dense(x.to(cuda_device))
    p = R.make_params()
    out = R.post_process_batch(rb, p, want_seg_info=True)
    rec = out.records_host()
    c = oracle.derive_constants()
    n_reject = 0
    for i in range(x.shape[0]):
        o = oracle.post_process_clip(x[i].numpy(), c)
        assert (rec["start"][i], rec["end"][i], rec["out_len"][i]) == (o["start"], o["end"], o["out_len"]), i
        y = out.audio.clip(i, o["out_len"]).cpu().numpy()
        assert_close(y, o["audio"], what=f"clip {i} audio")
        assert_close(rec["first_rms"][i], o["first_rms"], what="first_rms")
        assert_close(rec["last_rms"][i], o["last_rms"], what="last_rms")
        assert abs(rec["decay_ratio"][i] - o["decay_ratio"]) <= TOL * max(1.0, abs(o["decay_ratio"]))
        assert bool(rec["ok"][i]) == o["ok"], (i, rec["decay_ratio"][i], o["decay_ratio"])
        n_reject += not o["ok"]
    assert 0 < n_reject < x.shape[0]        # the workload exercises both decisions


def test_post_process_edge_lengths(R, cuda_device):
    rng = np.random.default_rng(11)
    lens = [0, 1, 2, 3, 7, 119, 240, 241, 959, 960, 961, 2000, 24000, 24001, 100003]
    clips = [tone_clip(rng, L, lead=min(L // 4, 700), trail=min(L // 5, 900)) for L in lens]
    clips += [rng.normal(0, 1e-4, L).astype(np.float32) for L in (1, 100, 240, 5000)]      # all-silent
    clips += [np.full(5000, 0.25, np.float32), np.zeros(4000, np.float32)]
    rb = _rb(R, clips, cuda_device)
    out = R.post_process_batch(rb, R.make_params())
    rec = out.records_host()
    c = oracle.derive_constants()
    for i, x in enumerate(clips):
        o = oracle.post_process_clip(x, c)
        assert (rec["start"][i], rec["end"][i], rec["out_len"][i]) == (o["start"], o["end"], o["out_len"]), (i, len(x))
        assert bool(rec["flags"][i] & 4) == o["all_silent"]
        assert_close(out.audio.clip(i, o["out_len"]).cpu().numpy(), o["audio"], what=f"edge clip {i}")
        if o["first_rms"] > 1e-6:       # away from the 1e-8 early-out, decisions must agree
            assert bool(rec["ok"][i]) == o["ok"]
            assert abs(rec["decay_ratio"][i] - o["decay_ratio"]) <= TOL * max(1.0, abs(o["decay_ratio"]))


# ----------------------------------------------------------------------------- join
def _join_case(rng, sr=24000):
    n = int(rng.integers(1, 7))
    segs = []
    for _ in range(n):
        L = int(rng.choice([0, 5, 200, 600, 1199, 1200, 1201, 1500, 2411, 5000, 30000, 48017]))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            segs.append(rng.normal(0, 1e-4, L).astype(np.float32))
        else:
            segs.append(tone_clip(rng, L, lead=min(L // 4, int(rng.integers(0, 3000))),
                                  trail=min(L // 5, int(rng.integers(0, 3000)))))
    return segs


@pytest.mark.parametrize("pause_sec", [0.1, 0.0])
def test_join_batch_vs_oracle(R, cuda_device, pause_sec):
    rng = np.random.default_rng(2024)
    items = [_join_case(rng) for _ in range(160)]
    segs = [s for it in items for s in it]
    first = np.concatenate([[0], np.cumsum([len(it) for it in items])]).astype(np.int32)
    rb = _rb(R, segs, cuda_device)
    p = R.make_params(inter_sentence_pause_sec=pause_sec)
    out = R.join_batch(rb, first, p)
    rec, seg = out.records_host(), out.seg_info_host()
    c = oracle.derive_constants(pause_sec=pause_sec)
    n_fb = 0
    for i, it in enumerate(items):
        o = oracle.smooth_segment_join(it, c)
        assert int(rec["out_len"][i]) == o.audio.size, (i, [len(s) for s in it], rec[i], o.fallback)
        assert bool(rec["flags"][i] & 2) == o.fallback, i
        assert bool(rec["flags"][i] & 4) == o.two_d, i
        for k, tr in enumerate(o.plan.trims):
            s = first[i] + k
            assert (int(seg["start"][s]), int(seg["end"][s])) == (tr.start, tr.end), (i, k)
        assert_close(out.audio.clip(i, o.audio.size).cpu().numpy(), o.audio, what=f"item {i}")
        ratio, ok, fr, lr = oracle.sound_decay(o.audio, 0.3)
        if fr > 1e-6:
            assert bool(rec["ok"][i]) == ok and abs(rec["decay_ratio"][i] - ratio) <= TOL * max(1, abs(ratio))
        n_fb += o.fallback
    assert n_fb > 10


@pytest.mark.parametrize("xfade_sec", [0.0, 0.0004])
def test_join_with_crossfade_disabled(R, cuda_device, xfade_sec):
    """crossfade_samples == 0 drops segment 0 like the reference's [..., :-0] slice (base_tts.py:485); <= 10 samples:
    no crossfade.  Golden vectors from the reference (tests/golden/make_golden_cf0.py) and the oracle on random items."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
    GCF = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cf0_v1.npz"))
    v = [float(x) for x in GCF["xfade_secs"]].index(xfade_sec)
    rng = np.random.default_rng(77)
    items = [[G[f"clip{j}"] for j in GCF[f"v{v}_item{k}_idx"]] for k in range(int(GCF["n_items"]))]
    n_golden = len(items)
    items += [_join_case(rng) for _ in range(60)]
    segs = [s for it in items for s in it]
    first = np.concatenate([[0], np.cumsum([len(it) for it in items])]).astype(np.int32)
    p = R.make_params(crossfade_duration_sec=xfade_sec)
    out = R.join_batch(_rb(R, segs, cuda_device), first, p)
    rec = out.records_host()
    c = oracle.derive_constants(xfade_sec=xfade_sec)
    for i, it in enumerate(items):
        o = oracle.smooth_segment_join(it, c)
        assert int(rec["out_len"][i]) == o.audio.size, (i, [len(s) for s in it], rec[i], o.fallback)
        assert bool(rec["flags"][i] & 2) == o.fallback and bool(rec["flags"][i] & 4) == o.two_d, i
        got = out.audio.clip(i, o.audio.size).cpu().numpy()
        assert_close(got, o.audio, what=f"item {i}")
        if i < n_golden:
            assert_close(got, GCF[f"v{v}_item{i}"], what=f"golden item {i}")


# ----------------------------------------------------------------------------- resample
def test_resample_vs_oracle(R, cuda_device):
    rng = np.random.default_rng(5)
    lens = [0, 1, 2, 3, 4, 5, 7, 22, 23, 24, 1000, 1535, 1536, 1537, 6143, 6144, 6145, 24001, 240000, 719999]
    clips = [rng.normal(0, 0.3, L).astype(np.float32) for L in lens]
    rb = _rb(R, clips, cuda_device)
    out = R.resample_batch(rb)
    got_len = out.lengths.cpu().numpy()
    for i, x in enumerate(clips):
        ref = oracle.resample(x) if x.size else np.zeros(0, np.float32)
        assert int(got_len[i]) == ref.size == -(-2 * x.size // 3), (i, x.size)
        assert_close(out.clip(i, ref.size).cpu().numpy(), ref, what=f"resample L={x.size}")


def test_resample_linearity_full_size(R, cuda_device):
    from rho_tts_b200 import synth
    x = synth.make_clip_block(64, 240000, 99, device=cuda_device)
    a = R.resample_batch(R.RaggedBatch.from_dense(x)).data
    b = R.resample_batch(R.RaggedBatch.from_dense((2.0 * x).contiguous())).data
    assert torch.equal(2.0 * a, b)          # scaling by a power of two commutes exactly with an fp32 FIR


# ----------------------------------------------------------------------------- log-mel
@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("pad", [True, False])
def test_logmel_vs_oracle(R, cuda_device, n_mels, pad):
    from rho_tts_b200 import synth
    lens = [160000, 159999, 16000, 4000, 480000 if pad else 200000, 201 if not pad else 50, 100001]
    if pad:
        lens += [500000, 479999, 479800]     # truncation and the right-edge reflection
    clips = [oracle.resample(synth.make_clip_block(1, (3 * L + 1) // 2, 40 + i)[0].numpy())[:L] for i, L in enumerate(lens)]
    rb = _rb(R, clips, cuda_device)
    mel, n_frames = R.logmel_batch(rb, n_mels=n_mels, pad_to_30s=pad)
    mel = mel.cpu().numpy(); n_frames = n_frames.cpu().numpy()
    worst = 0.0
    for i, w in enumerate(clips):
        ref = oracle.log_mel(w, n_mels, pad)
        assert int(n_frames[i]) == ref.shape[1], (i, len(w))
        worst = max(worst, assert_close(mel[i][:, :ref.shape[1]], ref, what=f"logmel L16={len(w)}"))
    print(f"log-mel {n_mels} pad={pad}: worst abs-or-rel error vs oracle {worst:.2e}")


def test_logmel_silence_and_constant(R, cuda_device):
    clips = [np.zeros(16000, np.float32), np.full(32000, 0.5, np.float32), np.zeros(300, np.float32)]
    rb = _rb(R, clips, cuda_device)
    mel, _ = R.logmel_batch(rb, 80, True)
    mel = mel.cpu().numpy()
    for i, w in enumerate(clips):
        assert_close(mel[i], oracle.log_mel(w, 80, True), what=f"degenerate {i}")


# ----------------------------------------------------------------------------- cosine
def test_cosine_vs_oracle(R, cuda_device):
    from rho_tts_b200 import synth
    emb, ref = synth.make_embeddings(300, 256, 4321)
    got = R.cosine_batch(emb.to(cuda_device), ref.to(cuda_device)).cpu().numpy()
    want = np.array([oracle.cosine_similarity(ref.numpy(), e) for e in emb.numpy()], dtype=np.float32)
    assert_close(got, want, what="cosine")
    g = torch.Generator().manual_seed(1)
    emb2 = torch.randn(17, 100, generator=g); ref2 = torch.randn(100, generator=g)     # odd dim, signed
    got = R.cosine_batch(emb2.to(cuda_device), ref2.to(cuda_device)).cpu().numpy()
    want = np.array([oracle.cosine_similarity(ref2.numpy(), e) for e in emb2.numpy()], dtype=np.float32)
    assert_close(got, want, what="cosine odd dim")


# ----------------------------------------------------------------------------- whole front end
def _oracle_pipeline(x, c, n_mels, emb, ref):
    o = oracle.post_process_clip(x, c)
    w16 = oracle.resample(o["audio"])
    return o, oracle.log_mel(w16, n_mels, True), oracle.cosine_similarity(ref, emb)


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("n_mels", [80, 128])
def test_validate_pipeline_vs_oracle(R, cuda_device, n_mels, fuse):
    from rho_tts_b200 import synth
    x = synth.make_clip_block(12, 240000, 1234 + 1)
    emb, ref = synth.make_embeddings(12)
    rb = R.RaggedBatch.from_dense(x.to(cuda_device))
    out = R.validate_batch(rb, R.make_params(), emb.to(cuda_device), ref.to(cuda_device), n_mels=n_mels, fuse=fuse)
    rec = out.records_host(); mel = out.mel.cpu().numpy()
    c = oracle.derive_constants()
    for i in range(x.shape[0]):
        o, m, cs = _oracle_pipeline(x[i].numpy(), c, n_mels, emb[i].numpy(), ref.numpy())
        assert (rec["start"][i], rec["end"][i], rec["out_len"][i], bool(rec["ok"][i])) == \
            (o["start"], o["end"], o["out_len"], o["ok"])
        assert_close(out.audio.clip(i, o["out_len"]).cpu().numpy(), o["audio"], what="audio")
        assert_close(mel[i], m, what=f"mel clip {i}")
        assert_close(rec["cosine"][i], cs, what="cosine")


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("pad", [True, False])
def test_validate_ragged_lengths_vs_oracle(R, cuda_device, pad, fuse):
    """Ragged one-clip items from 4 ms to beyond Whisper's 30 s window (truncated features, full audio)."""
    from rho_tts_b200 import synth
    rng = np.random.default_rng(77)
    lens = [100, 959, 5000, 7680, 24000, 100001, 184320, 480000]
    lens += [730000] if pad else [300]
    clips = [synth.make_clip_block(1, L, 500 + i)[0].numpy() for i, L in enumerate(lens)]
    clips.append(rng.normal(0, 1e-4, 30000).astype(np.float32))                    # all silent -> 240 samples
    rb = _rb(R, clips, cuda_device)
    out = R.validate_batch(rb, R.make_params(), None, None, n_mels=80, pad_to_30s=pad, fuse=fuse)
    rec = out.records_host(); mel = out.mel.cpu().numpy()
    c = oracle.derive_constants()
    for i, x in enumerate(clips):
        o = oracle.post_process_clip(x, c)
        assert (rec["start"][i], rec["end"][i], rec["out_len"][i]) == (o["start"], o["end"], o["out_len"]), (i, len(x))
        assert_close(out.audio.clip(i, o["out_len"]).cpu().numpy(), o["audio"], what=f"audio {i}")
        if o["first_rms"] > 1e-6:
            assert bool(rec["ok"][i]) == o["ok"]
            assert abs(rec["decay_ratio"][i] - o["decay_ratio"]) <= TOL * max(1.0, abs(o["decay_ratio"]))
        w16 = oracle.resample(o["audio"])
        if not pad and w16.size <= 200:
            continue                                  # torch.stft cannot reflect-pad such a clip: no features
        m = oracle.log_mel(w16, 80, pad)
        assert_close(mel[i][:, :m.shape[1]], m, what=f"mel {i} (len {len(x)})")


def test_fused_and_unfused_paths_agree(R, cuda_device):
    from rho_tts_b200 import synth
    x = synth.make_clip_block(40, 120000, 31, device=cuda_device)
    emb, ref = synth.make_embeddings(40, device=cuda_device)
    rb = R.RaggedBatch.from_dense(x)
    a = R.validate_batch(rb, R.make_params(), emb, ref, fuse=True)
    ra, ma, ya = a.records_host(), a.mel.clone(), a.audio.data.clone()
    b = R.validate_batch(rb, R.make_params(), emb, ref, fuse=False)
    rb_ = b.records_host()
    for f in ("start", "end", "out_len", "ok", "flags", "cosine"):
        assert np.array_equal(ra[f], rb_[f]), f
    assert np.allclose(ra["decay_ratio"], rb_["decay_ratio"], rtol=1e-6)
    for i in range(40):
        L = int(ra["out_len"][i])
        assert torch.equal(a.audio.clip(i, L), b.audio.clip(i, L)) or \
            torch.equal(ya[a.audio.h_offsets[i]:a.audio.h_offsets[i] + L], b.audio.clip(i, L))
    # two fp32 evaluation orders of the same FIR / FFT: they differ by rounding only, amplified on the
    # -70 dB bins exactly like the reference's own fp32 noise (tests/diagnostics/mel_error.py: each path is
    # 4e-5 .. 7e-5 from the fp64 value, as is the numpy oracle)
    d = float((ma - b.mel).abs().max())
    assert d < 1e-4, d


def test_validate_host_matches_device_path(R, cuda_device):
    """The host-buffer entry point against the device-resident path, byte for byte: only the frames that can see signal
    cross PCIe, the constant tail of every row is written on the host (include/rho_b200.h)."""
    from rho_tts_b200 import synth
    n = 150                                   # > 2 chunks of 64, last one ragged
    x = synth.make_clip_block(n, 48000, 77).pin_memory()
    emb, ref = synth.make_embeddings(n)
    p = R.make_params()
    mel_h = torch.full((n, 80, 3000), float("nan"), dtype=torch.float32).pin_memory()
    y, mel_h, rec_h = R.validate_host(x, p, emb.pin_memory(), ref.pin_memory(), 80, mel=mel_h)
    out = R.validate_batch(R.RaggedBatch.from_dense(x.to(cuda_device)), p, emb.to(cuda_device), ref.to(cuda_device))
    rec_d = out.records_host(); rec_h = rec_h.numpy().view(R.REC_DTYPE).reshape(-1)
    _records_match(rec_d, rec_h)
    assert torch.equal(out.mel.cpu(), mel_h)
    for i in range(n):
        L = int(rec_d["out_len"][i])
        assert torch.equal(out.audio.clip(i, L).cpu(), y[i, :L])
    # the per-clip constant on the device is the tail of the full rows
    assert torch.equal(out.pad_value.cpu(), mel_h[:, 0, 2999])


def test_validate_host_features_stay_in_hbm(R, cuda_device):
    """`mel` of the host entry points may be DEVICE memory: audio and records come back to the host, the features are
    written in place in HBM for a consumer on the device (SURVEY 8f NEXT-2) -- the same bytes as the all-host call."""
    from rho_tts_b200 import synth
    n = 150
    x = synth.make_clip_block(n, 48000, 78).pin_memory()
    emb, ref = synth.make_embeddings(n)
    emb, ref = emb.pin_memory(), ref.pin_memory()
    p = R.make_params()
    y_h, mel_h, rec_h = R.validate_host(x, p, emb, ref, 80, mel=torch.empty((n, 80, 3000)).pin_memory())
    mel_d = torch.full((n, 80, 3000), float("nan"), dtype=torch.float32, device=cuda_device)
    y_d, mel_d2, rec_d = R.validate_host(x, p, emb, ref, 80, mel=mel_d)
    assert mel_d2.is_cuda and mel_d2.data_ptr() == mel_d.data_ptr()
    assert torch.equal(mel_d.cpu(), mel_h) and torch.equal(rec_d, rec_h)
    rec = rec_h.numpy().view(R.REC_DTYPE).reshape(-1)
    for i in range(n):
        L = int(rec["out_len"][i])
        assert torch.equal(y_d[i, :L], y_h[i, :L])
    # ragged entry point, compact rows + pad values, joined items; features on the device
    rng = np.random.default_rng(17)
    seg_lens = rng.integers(2400, 200000, size=60).astype(np.int32)
    first = np.asarray([0, 1, 4, 6, 10, 11, 15, 20, 26, 30, 33, 40, 41, 47, 52, 60], np.int32)
    padded = (seg_lens.astype(np.int64) + 31) // 32 * 32
    seg_off = np.concatenate([[0], np.cumsum(padded)])[:-1].astype(np.int64)
    flat = torch.zeros(int(seg_off[-1] + padded[-1]), dtype=torch.float32)
    for k, L in enumerate(seg_lens):
        flat[seg_off[k]:seg_off[k] + L] = synth.make_clip_block(1, int(L), 7000 + k)[0]
    flat = flat.pin_memory()
    a = R.validate_host_ragged(flat, seg_off, seg_lens, first, p, compact=True)
    T = a.mel.shape[2]
    mel_dev = torch.full((len(first) - 1, 80, T), float("nan"), dtype=torch.float32, device=cuda_device)
    b = R.validate_host_ragged(flat, seg_off, seg_lens, first, p, compact=True, mel=mel_dev)
    assert b.mel.data_ptr() == mel_dev.data_ptr()
    assert torch.equal(mel_dev.cpu(), a.mel) and torch.equal(a.pad_value, b.pad_value)
    assert np.array_equal(a.records.view(np.uint8), b.records.view(np.uint8))


@pytest.mark.parametrize("n_mels", [80, 128])
def test_compact_feature_rows(R, cuda_device, n_mels):
    """RHO_V_COMPACT_PAD: rows of rho_b200_compact_frames(L) frames + pad_value[i] == the head and the constant tail
    of the complete [n_mels][3000] rows, bit for bit (feature_extraction_whisper.py:296-303)."""
    from rho_tts_b200 import synth
    lens = [240000, 200001, 100, 24000, 730000, 7680]
    clips = [synth.make_clip_block(1, L, 900 + i)[0].numpy() for i, L in enumerate(lens)]
    for group in (clips[:4], clips):                     # 10 s rows (1004 frames) and rows that hit the 3000-frame window
        rb = _rb(R, group, cuda_device)
        full = R.validate_batch(rb, R.make_params(), n_mels=n_mels)
        mel_full, rec_full = full.mel.clone(), full.records.clone()
        comp = R.validate_batch(rb, R.make_params(), n_mels=n_mels, compact=True)
        T = comp.mel.shape[2]
        assert T == int(R._lib.load().rho_b200_compact_frames(max(len(c) for c in group), 3000))
        assert torch.equal(comp.records, rec_full)
        assert torch.equal(comp.mel, mel_full[:, :, :T])
        if T < 3000:
            tail = mel_full[:, :, T:]
            assert torch.equal(tail, comp.pad_value[:, None, None].expand_as(tail))


@pytest.mark.parametrize("pad", [True, False])
def test_validate_joined_items_vs_oracle(R, cuda_device, pad):
    """Items of several segments through the whole front end (base_tts.py:912-926 then the features): join -> y, then the
    fused kernel reads the finished y (resample + log-mel).  Against the oracle chain and against the kernel-per-stage path."""
    from rho_tts_b200 import synth
    rng = np.random.default_rng(4242)
    items = []
    for k in range(14):
        n = int(rng.integers(1, 5))
        segs = []
        for _ in range(n):
            L = int(rng.choice([300, 5000, 24000, 48017, 100001, 250000]))
            if rng.integers(0, 7) == 0:
                segs.append(rng.normal(0, 1e-4, L).astype(np.float32))           # all-silent segment -> fallback items
            else:
                segs.append(synth.make_clip_block(1, L, 3000 + 17 * k + len(segs))[0].numpy())
        items.append(segs)
    if pad:
        items.append([synth.make_clip_block(1, 400000, 3900 + j)[0].numpy() for j in range(2)])   # > 30 s: truncated features
    segs = [s for it in items for s in it]
    first = np.concatenate([[0], np.cumsum([len(it) for it in items])]).astype(np.int32)
    emb, ref = synth.make_embeddings(len(items))
    rb = _rb(R, segs, cuda_device)
    p = R.make_params()
    out = R.validate_batch(rb, p, emb.to(cuda_device), ref.to(cuda_device), n_mels=80, pad_to_30s=pad, item_first_seg=first)
    rec = out.records_host(); mel = out.mel.cpu().numpy()
    audio = [out.audio.clip(i, int(rec["out_len"][i])).cpu().numpy() for i in range(len(items))]
    c = oracle.derive_constants()
    for i, it in enumerate(items):
        o = oracle.smooth_segment_join(it, c)
        assert int(rec["out_len"][i]) == o.audio.size and bool(rec["flags"][i] & 2) == o.fallback, i
        assert_close(audio[i], o.audio, what=f"item {i} audio")
        ratio, ok, fr, _ = oracle.sound_decay(o.audio, 0.3)
        if fr > 1e-6:
            assert bool(rec["ok"][i]) == ok
        assert_close(rec["cosine"][i], oracle.cosine_similarity(ref.numpy(), emb[i].numpy()), what="cosine")
        w16 = oracle.resample(o.audio)
        if not pad and w16.size <= 200:
            continue
        m = oracle.log_mel(w16, 80, pad)
        assert_close(mel[i][:, :m.shape[1]], m, what=f"mel item {i}")
    stage = R.validate_batch(rb, p, emb.to(cuda_device), ref.to(cuda_device), n_mels=80, pad_to_30s=pad,
                             item_first_seg=first, fuse=False)
    rs = stage.records_host()
    for f in ("start", "end", "out_len", "ok", "flags", "cosine", "n_segments"):
        assert np.array_equal(rec[f], rs[f]), f
    for i in range(len(items)):                      # the frames each item has (rows are not written past them when unpadded)
        T_i = 3000 if pad else ((2 * int(rec["out_len"][i]) + 2) // 3) // 160
        if T_i > 0 and (pad or (2 * int(rec["out_len"][i]) + 2) // 3 > 200):
            assert float((out.mel[i, :, :T_i] - stage.mel[i, :, :T_i]).abs().max()) < 1e-4, i
    # the single-kernel join (default) against "k_gather writes y, the feature kernel reads it back": same samples, bit
    # for bit; the features differ by the rounding of the DC folded into the FIR of interior batches
    gf = R.validate_batch(rb, p, emb.to(cuda_device), ref.to(cuda_device), n_mels=80, pad_to_30s=pad,
                          item_first_seg=first, gather_first=True)
    rg = gf.records_host()
    for f in ("start", "end", "out_len", "ok", "flags", "cosine", "n_segments"):
        assert np.array_equal(rec[f], rg[f]), f
    assert np.allclose(rec["first_rms"], rg["first_rms"], rtol=2e-6, atol=1e-9)
    assert np.allclose(rec["last_rms"], rg["last_rms"], rtol=2e-6, atol=1e-9)
    for i in range(len(items)):
        L = int(rec["out_len"][i])
        assert torch.equal(out.audio.clip(i, L), gf.audio.clip(i, L)), i
        T_i = 3000 if pad else ((2 * L + 2) // 3) // 160
        if T_i > 0 and (pad or (2 * L + 2) // 3 > 200):
            assert float((out.mel[i, :, :T_i] - gf.mel[i, :, :T_i]).abs().max()) < 6e-5, i


def test_validate_joined_many_short_segments(R, cuda_device):
    """The single-kernel join on items of up to 14 segments (more than a half caches), segments shorter than a batch's
    window (several joints per window), odd lengths (segments at every alignment inside their item), no pauses, a
    crossfade longer than some segments: against k_gather's output bit for bit and against the oracle."""
    from rho_tts_b200 import synth
    rng = np.random.default_rng(777)
    for kw in (dict(), dict(inter_sentence_pause_sec=0.0, crossfade_duration_sec=0.11), dict(trim_silence=False)):
        items = []
        for k in range(10):
            n = int(rng.integers(2, 15))
            items.append([synth.make_clip_block(1, int(rng.integers(700, 30000)) if rng.integers(0, 3) else
                                                int(rng.integers(30000, 200001)), 9000 + 31 * k + j)[0].numpy()
                          for j in range(n)])
        segs = [s for it in items for s in it]
        first = np.concatenate([[0], np.cumsum([len(it) for it in items])]).astype(np.int32)
        rb = _rb(R, segs, cuda_device)
        p = R.make_params(**kw)
        a = R.validate_batch(rb, p, n_mels=80, pad_to_30s=True, item_first_seg=first)
        b = R.validate_batch(rb, p, n_mels=80, pad_to_30s=True, item_first_seg=first, gather_first=True)
        ra, rb_ = a.records_host(), b.records_host()
        for f in ("start", "end", "out_len", "ok", "flags", "n_segments"):
            assert np.array_equal(ra[f], rb_[f]), f
        c = oracle.derive_constants(xfade_sec=kw.get("crossfade_duration_sec", 0.05),
                                    pause_sec=kw.get("inter_sentence_pause_sec", 0.1))
        for i, it in enumerate(items):
            L = int(ra["out_len"][i])
            assert torch.equal(a.audio.clip(i, L), b.audio.clip(i, L)), (kw, i)
            o = oracle.smooth_segment_join(it, c, kw.get("trim_silence", True))
            assert L == o.audio.size
            assert_close(a.audio.clip(i, L).cpu().numpy(), o.audio, what=f"item {i} audio")
            assert float((a.mel[i] - b.mel[i]).abs().max()) < 6e-5, (kw, i)


def test_validate_joined_degenerate_items(R, cuda_device):
    """The single-kernel join on degenerate items: empty segments, items made of empty segments only, one-sample and
    all-silent segments, an item whose segments are all shorter than the crossfade -- against k_gather's output."""
    from rho_tts_b200 import synth
    rng = np.random.default_rng(31)
    tone = lambda n, k: synth.make_clip_block(1, n, 8100 + k)[0].numpy()          # noqa: E731
    z = np.zeros(0, np.float32)
    items = [[z, z], [z, tone(30000, 0), z], [tone(1, 1), tone(50000, 2)], [tone(700, 3), tone(900, 4), tone(800, 5)],
             [rng.normal(0, 1e-5, 20000).astype(np.float32), tone(40000, 6)], [tone(90000, 7)], [z],
             [tone(1199, 8), tone(1201, 9), tone(11, 10), tone(60000, 11)]]
    segs = [s for it in items for s in it]
    first = np.concatenate([[0], np.cumsum([len(it) for it in items])]).astype(np.int32)
    rb = _rb(R, segs, cuda_device)
    p = R.make_params()
    for pad in (True, False):
        a = R.validate_batch(rb, p, n_mels=80, pad_to_30s=pad, item_first_seg=first)
        b = R.validate_batch(rb, p, n_mels=80, pad_to_30s=pad, item_first_seg=first, gather_first=True)
        ra, rb_ = a.records_host(), b.records_host()
        for f in ("start", "end", "out_len", "ok", "flags", "n_segments"):
            assert np.array_equal(ra[f], rb_[f]), (pad, f)
        c = oracle.derive_constants()
        for i, it in enumerate(items):
            L = int(ra["out_len"][i])
            assert torch.equal(a.audio.clip(i, L), b.audio.clip(i, L)), (pad, i)
            o = oracle.smooth_segment_join(it, c)
            assert L == (0 if o.audio is None else o.audio.size), (pad, i)
            T_i = 3000 if pad else ((2 * L + 2) // 3) // 160
            if T_i > 0 and (pad or (2 * L + 2) // 3 > 200):
                assert torch.isfinite(a.mel[i, :, :T_i]).all()
                assert float((a.mel[i, :, :T_i] - b.mel[i, :, :T_i]).abs().max()) < 6e-5, (pad, i)


def test_validate_host_ragged_joins(R, cuda_device):
    """rho_b200_validate_host_ragged: ragged segments and items in host memory (BASELINE config C3's shape, small), several
    chunks, with and without features, aligned and unaligned offsets -- against the device-resident calls, byte for byte."""
    from rho_tts_b200 import synth
    rng = np.random.default_rng(99)
    seg_lens = rng.integers(2400, 260000, size=220).astype(np.int32)     # 29 M samples: two chunks
    first = [0]
    while first[-1] < len(seg_lens):
        first.append(min(len(seg_lens), first[-1] + int(rng.integers(1, 5))))
    first = np.asarray(first, np.int32)
    n_items = len(first) - 1
    p = R.make_params()
    emb, ref = synth.make_embeddings(n_items)
    for align in (32, 1):
        padded = (seg_lens.astype(np.int64) + align - 1) // align * align + (0 if align > 1 else 3)
        seg_off = np.concatenate([[0], np.cumsum(padded)])[:-1].astype(np.int64)
        x = torch.zeros(int(seg_off[-1] + padded[-1]), dtype=torch.float32)
        for s, L in enumerate(seg_lens):
            x[seg_off[s]:seg_off[s] + L] = synth.make_clip_block(1, int(L), 5000 + s)[0]
        x = x.pin_memory()
        rb = R.RaggedBatch.from_list([x[seg_off[s]:seg_off[s] + int(L)] for s, L in enumerate(seg_lens)], cuda_device)
        # (a) join + decay only (mel is NULL): what config C3 asks for
        h = R.validate_host_ragged(x, seg_off, seg_lens, first, p, features=False)
        d = R.join_batch(rb, first, p, want_seg_info=False)
        rd = d.records_host()
        _records_match(rd, h.records)
        for i in range(n_items):
            L = int(rd["out_len"][i])
            assert torch.equal(d.audio.clip(i, L).cpu(), h.audio[h.y_offsets[i]:h.y_offsets[i] + L]), i
        # (b) with features, complete rows and compact rows
        dv = R.validate_batch(rb, p, emb.to(cuda_device), ref.to(cuda_device), item_first_seg=first)
        rdv = dv.records_host()
        hv = R.validate_host_ragged(x, seg_off, seg_lens, first, p, emb.pin_memory(), ref.pin_memory())
        _records_match(rdv, hv.records)
        assert torch.equal(dv.mel.cpu(), hv.mel) and torch.equal(dv.pad_value.cpu(), hv.pad_value)
        hc = R.validate_host_ragged(x, seg_off, seg_lens, first, p, emb.pin_memory(), ref.pin_memory(), compact=True)
        T = hc.mel.shape[2]
        assert torch.equal(hc.mel, hv.mel[:, :, :T]) and torch.equal(hc.pad_value, hv.pad_value)
        for i in range(0, n_items, 3):
            L = int(rdv["out_len"][i])
            assert torch.equal(dv.audio.clip(i, L).cpu(), hv.audio[hv.y_offsets[i]:hv.y_offsets[i] + L]), i


def test_host_fill_threads_adapt_to_a_slow_host(cuda_device):
    """A call whose host fill ends well behind its last copy gives the handle's next calls two more fill threads (a host
    with slow memory; simulated with RHO_HOST_DEBUG_FILL_DELAY_US, read once per process: subprocess).  The outputs do
    not depend on the thread count."""
    import subprocess
    import sys
    code = """
import os, sys, torch
sys.path.insert(0, os.getcwd())
import rho_tts_b200 as R
from rho_tts_b200 import synth, _lib
h = _lib.Handle.get(0)
x = synth.make_clip_block(200, 48000, 5).pin_memory()
p = R.make_params()
seen = [int(h.lib.rho_b200_host_fill_threads(h.ptr))]
outs = []
for _ in range(4):
    y, mel, rec = R.validate_host(x, p, None, None, 80, mel=torch.full((200, 80, 3000), float("nan")).pin_memory())
    outs.append((y.clone(), mel.clone(), rec.clone()))
    seen.append(int(h.lib.rho_b200_host_fill_threads(h.ptr)))
assert all(torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2]) for o in outs[1:])
assert not torch.isnan(outs[0][1]).any()
print("THREADS", seen)
"""
    env = dict(os.environ, RHO_HOST_DEBUG_FILL_DELAY_US="4000")
    env.pop("RHO_HOST_FILL_THREADS", None)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    seen = eval(r.stdout.strip().splitlines()[-1].split("THREADS", 1)[1])
    assert seen[0] == 3 and seen[1] == 5 and seen[-1] == 9 and seen == sorted(seen), seen
    # a host that keeps up stays at the default; a pinned count is never changed
    env2 = {k: v for k, v in os.environ.items() if k not in ("RHO_HOST_DEBUG_FILL_DELAY_US",)}
    env2["RHO_HOST_FILL_THREADS"] = "2"
    r2 = subprocess.run([sys.executable, "-c", code], env=dict(env2, RHO_HOST_DEBUG_FILL_DELAY_US="4000"), capture_output=True,
                        text=True, cwd=ROOT, timeout=300)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert eval(r2.stdout.strip().splitlines()[-1].split("THREADS", 1)[1]) == [2] * 5


def test_validate_host_odd_clip_length_and_unpadded_errors(R, cuda_device):
    """clip_len % 4 != 0: the host rows are not 16-byte aligned, the entry point falls back to one copy per clip."""
    from rho_tts_b200 import synth
    n, L = 9, 30001
    x = synth.make_clip_block(n, L, 13).pin_memory()
    p = R.make_params()
    y, _, rec_h = R.validate_host(x, p, None, None, 80)
    out = R.post_process_batch(R.RaggedBatch.from_dense(x.to(cuda_device)), p)
    rec_d = out.records_host(); rec_h = rec_h.numpy().view(R.REC_DTYPE).reshape(-1)
    for f in ("start", "end", "out_len", "ok"):
        assert np.array_equal(rec_d[f], rec_h[f]), f
    for i in range(n):
        k = int(rec_d["out_len"][i])
        assert torch.equal(out.audio.clip(i, k).cpu(), y[i, :k])


def test_host_calls_from_concurrent_threads(R, cuda_device):
    """One shared handle, four threads in the host entry point at once (a provider instance is shared by concurrent
    sessions, ui/state.py:85-87 of the reference): every caller gets its own streams / arena, results are unaffected."""
    import threading
    from rho_tts_b200 import synth
    p = R.make_params()
    xs = [synth.make_clip_block(40, 36000, 600 + t).pin_memory() for t in range(4)]
    want = []
    for x in xs:
        y, mel, rec = R.validate_host(x, p, None, None, 80, mel=torch.empty((40, 80, 3000)).pin_memory())
        want.append((y.clone(), mel.clone(), rec.clone()))
    got = [None] * 4
    errs = []

    def work(t):
        try:
            for _ in range(3):
                y, mel, rec = R.validate_host(xs[t], p, None, None, 80, mel=torch.empty((40, 80, 3000)).pin_memory())
            got[t] = (y, mel, rec)
        except Exception as e:          # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(4):
        rw = want[t][2].numpy().view(R.REC_DTYPE).reshape(-1); rg = got[t][2].numpy().view(R.REC_DTYPE).reshape(-1)
        _records_match(rw, rg)
        assert torch.equal(want[t][1], got[t][1])
        for i in range(40):
            k = int(rw["out_len"][i])
            assert torch.equal(want[t][0][i, :k], got[t][0][i, :k])


# ----------------------------------------------------------------------------- size-independent properties
def test_full_size_properties(R, cuda_device):
    """BASELINE config C2 size: 1000 x 10 s.  Checks that do not need the oracle at full size."""
    from rho_tts_b200 import synth
    x = synth.make_clip_block(1000, 240000, 0xB200, device=cuda_device)
    emb, ref = synth.make_embeddings(1000, device=cuda_device)
    rb = R.RaggedBatch.from_dense(x)
    p = R.make_params()
    out = R.validate_batch(rb, p, emb, ref)
    rec = out.records_host()
    assert np.all(rec["out_len"] == rec["end"] - rec["start"]) and np.all(rec["start"] % 120 == 0)
    assert np.all((rec["end"] % 120 == 0) | (rec["end"] == 240000))
    assert 0.05 < 1.0 - rec["ok"].mean() < 0.5
    assert np.all(np.abs(rec["cosine"]) <= 1.0 + 1e-6)
    mel = out.mel
    assert torch.isfinite(mel).all()
    # frames that only see zero padding are one constant per clip: (max(-10, m-8)+4)/4
    tail = mel[:, :, 1010:]
    assert torch.equal(tail.amax(dim=(1, 2)), tail.amin(dim=(1, 2)))
    # ... and it is the constant of the FINAL clip maximum (the half that finishes a clip last must have seen the maxima
    # of all the other tiles): with m the clip maximum, the features peak at (m+4)/4 and the padding sits at (m-8+4)/4
    assert float((mel.amax(dim=(1, 2)) - 2.0 - tail[:, 0, 0]).abs().max()) <= 1e-6
    # the Whisper clamp: per clip, max - min <= 8/4
    assert float((mel.amax(dim=(1, 2)) - mel.amin(dim=(1, 2))).max()) <= 2.0 + 1e-6
    # idempotence of the trim: the processed clip starts and ends loud, so a second scan trims at most
    # the faded edges; DC of the output is ~0
    means = torch.stack([out.audio.clip(i, int(rec["out_len"][i])).double().mean() for i in range(0, 1000, 50)])
    assert float(means.abs().max()) < 5e-4
    # spot-check 4 clips of the big batch against the oracle
    c = oracle.derive_constants()
    for i in (0, 333, 999):
        o, m, cs = _oracle_pipeline(x[i].cpu().numpy(), c, 80, emb[i].cpu().numpy(), ref.cpu().numpy())
        assert (rec["start"][i], rec["end"][i], bool(rec["ok"][i])) == (o["start"], o["end"], o["ok"])
        assert_close(mel[i].cpu().numpy(), m, what=f"mel clip {i}")


# ----------------------------------------------------------------------------- mel projection on the tensor cores
@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("n_frames", [1, 47, 48, 49, 1000, 148 * 32 * 5 + 5])
def test_mel_project_tensor_core_vs_fp64(R, cuda_device, n_mels, n_frames):
    """rho_b200_mel_project (tcgen05, 3xTF32) against the float64 product with the oracle's filterbank
    (transformers audio_utils.py:453-544): fp32-class accuracy over a power range of 100 dB per frame."""
    g = torch.Generator().manual_seed(77 + n_frames)
    # power spectra with a wide dynamic range, like |STFT|^2 of speech; the 7 pad columns hold garbage
    p = torch.rand(n_frames, 208, generator=g) * torch.pow(10.0, torch.rand(n_frames, 208, generator=g) * 10.0 - 6.0)
    p[:, 201:] = float("nan")
    got = R.mel_project(p.to(cuda_device), n_mels).cpu().numpy().astype(np.float64)
    bank = oracle.slaney_mel_filterbank(n_mels).astype(np.float32).astype(np.float64)       # (201, n_mels)
    want = bank.T @ p[:, :201].numpy().astype(np.float64).T
    assert got.shape == want.shape == (n_mels, n_frames)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    assert float(err.max()) < 2e-6, float(err.max())


def test_mel_project_matches_product_filterbank_path(R, cuda_device):
    """The tensor-core projection of the power spectrum of real frames equals what the product log-mel
    kernel (sparse FFMA form) writes, through the same log/clamp/scale."""
    from rho_tts_b200 import synth
    x = synth.make_clip_block(2, 48000, 5)
    w16 = [oracle.resample(x[i].numpy()) for i in range(2)]
    for w in w16:
        power = oracle.stft_power(w).T.copy()                                                # (T, 201) fp32
        T = power.shape[0]
        pw = np.zeros((T, 204), np.float32); pw[:, :201] = power
        mel = R.mel_project(torch.from_numpy(pw).to(cuda_device), 80).cpu().numpy()
        ls = np.log10(np.maximum(mel, 1e-10)); ls = np.maximum(ls, ls.max() - 8.0)
        got = (ls + 4.0) / 4.0
        want = oracle.log_mel(w, 80, pad_to_30s=False)
        assert_close(got, want, what="log-mel through the tensor-core projection")


def test_mel_project_batched_layout(R, cuda_device):
    """frames of consecutive clips -> [clip][n_mels][T_out] (the Whisper feature layout), ragged last clip."""
    g = torch.Generator().manual_seed(5)
    n_frames, T = 5 * 100 - 37, 100
    p = torch.rand(n_frames, 204, generator=g)
    out = torch.full((5, 80, 128), -1.0, device=cuda_device)
    R.mel_project(p.to(cuda_device), 80, T, out)
    bank = oracle.slaney_mel_filterbank(80).astype(np.float32).astype(np.float64)
    want = bank.T @ p[:, :201].numpy().astype(np.float64).T                                   # (80, n_frames)
    got = out.cpu().numpy()
    for i in range(5):
        nt = min(T, n_frames - i * T)
        np.testing.assert_allclose(got[i, :, :nt], want[:, i * T:i * T + nt], rtol=2e-6)
        assert np.all(got[i, :, nt:] == -1.0)                                                 # nothing else is touched


def test_full_size_ragged_joins_properties(R, cuda_device):
    """BASELINE config C3 size: 4000 ragged clips of 1..30 s (6 GB) joined into ~1000 items of 2..6 segments.
    Size-independent checks, plus three items against the oracle."""
    from rho_tts_b200 import synth
    n = 4000
    lens = synth.make_ragged_lengths(n, 1234 + 3)
    rb = R.RaggedBatch.empty_like_lengths(lens, cuda_device)
    order = np.argsort(lens)
    for g0 in range(0, n, 50):                                   # groups of similar length share one generator call
        idx = order[g0:g0 + 50]
        blk = synth.make_clip_block(len(idx), int(lens[idx].max()), 7919 * 1237 + g0, device=cuda_device)
        for j, i in enumerate(idx):
            rb.clip(int(i)).copy_(blk[j, :int(lens[i])])
    first = synth.make_item_partition(n, 1234 + 3)
    p = R.make_params()
    out = R.join_batch(rb, first, p, want_seg_info=True)
    rec, seg = out.records_host(), out.seg_info_host()
    n_items = len(first) - 1
    assert rec.shape[0] == n_items and np.array_equal(rec["n_segments"], np.diff(first))
    # trim bounds: hop-aligned starts, ends hop-aligned or at the clip end, inside the clip
    assert np.all(seg["start"] % 120 == 0) and np.all((seg["end"] % 120 == 0) | (seg["end"] == lens))
    assert np.all((0 <= seg["start"]) & (seg["start"] <= seg["end"]) & (seg["end"] <= lens))
    # length bookkeeping of base_tts.py:481-523 for regular items (every segment longer than the crossfade): the kept
    # samples, minus one crossfade overlap per joint, plus one pause after every middle segment
    kept = (seg["end"] - seg["start"]).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(kept)])
    k = np.diff(first).astype(np.int64)
    regular = np.array([kept[first[i]:first[i + 1]].min() > 2400 for i in range(n_items)]) & ((rec["flags"] & 0x3) == 0)
    want_len = (csum[first[1:]] - csum[first[:-1]]) - (k - 1) * 1200 + np.maximum(0, k - 2) * 2400
    assert regular.mean() > 0.9
    assert np.array_equal(rec["out_len"][regular], want_len[regular])
    # the decay decision equals the one recomputed from the joined audio itself
    for i in range(0, n_items, 97):
        y = out.audio.clip(i, int(rec["out_len"][i])).double()
        third = y.numel() // 3
        if third < 1:
            continue
        fr, lr = float(y[:third].pow(2).mean().sqrt()), float(y[-third:].pow(2).mean().sqrt())
        if fr >= 1e-8:
            assert abs(lr / fr - rec["decay_ratio"][i]) <= 1e-4 * max(1.0, lr / fr)
            assert bool(rec["ok"][i]) == (lr / fr >= 0.3) or abs(lr / fr - 0.3) < 1e-4
    # determinism: the same call again gives the same bytes
    out2 = R.join_batch(rb, first, p, want_seg_info=False)
    assert torch.equal(out.records, out2.records)
    # three items against the oracle
    c = oracle.derive_constants()
    for i in (0, n_items // 2, n_items - 1):
        segs = [rb.clip(s).cpu().numpy() for s in range(first[i], first[i + 1])]
        o = oracle.smooth_segment_join(segs, c)
        L = int(rec["out_len"][i])
        assert L == o.audio.size
        assert_close(out.audio.clip(i, L).cpu().numpy(), o.audio, what=f"item {i}")


def test_full_size_c4_shard_128_bins(R, cuda_device):
    """One GPU's shard of BASELINE config C4: 8000 x 10 s clips, 128-bin log-mel (30 s pad), cosine: 28 GB resident."""
    from rho_tts_b200 import synth
    n = 8000
    x = synth.make_clip_block(n, 240000, 0xC4, device=cuda_device)
    emb, ref = synth.make_embeddings(n, device=cuda_device)
    out = R.validate_batch(R.RaggedBatch.from_dense(x), R.make_params(), emb, ref, n_mels=128)
    rec = out.records_host()
    mel = out.mel
    assert mel.shape == (n, 128, 3000)
    assert np.all(rec["out_len"] == rec["end"] - rec["start"]) and np.all(np.abs(rec["cosine"]) <= 1.0 + 1e-6)
    mx = mel.amax(dim=(1, 2))
    tail = mel[:, :, 1010:]
    assert torch.equal(tail.amax(dim=(1, 2)), tail.amin(dim=(1, 2)))
    assert float((mx - 2.0 - tail[:, 0, 0]).abs().max()) <= 1e-6          # padding constant of the final clip maximum
    assert float((mx - mel.amin(dim=(1, 2))).max()) <= 2.0 + 1e-6         # the Whisper clamp
    c = oracle.derive_constants()
    for i in (17, n - 1):
        o, m, cs = _oracle_pipeline(x[i].cpu().numpy(), c, 128, emb[i].cpu().numpy(), ref.cpu().numpy())
        assert (rec["start"][i], rec["end"][i], bool(rec["ok"][i])) == (o["start"], o["end"], o["ok"])
        assert_close(mel[i].cpu().numpy(), m, what=f"mel clip {i}")
        assert abs(rec["cosine"][i] - cs) <= 1e-5


# ----------------------------------------------------------------------------- windowed DFT on the tensor cores
def _frames64(w16, pad):
    """The frames torch.stft(center=True, reflect) cuts out of the (padded) 16 kHz clip, float64."""
    w = np.asarray(w16, np.float64)
    if pad:
        buf = np.zeros(480000); buf[:min(w.size, 480000)] = w[:480000]; w = buf
    p = np.concatenate([w[200:0:-1], w, w[-2:-202:-1]])
    return np.lib.stride_tricks.sliding_window_view(p, 400)[::160][:w.size // 160]


@pytest.mark.parametrize("pad", [True, False])
def test_stft_power_tensor_core_vs_fp64(R, cuda_device, pad):
    """rho_b200_stft_power_tc (the windowed DFT factored 400 = 25 x 16 into two tcgen05 GEMMs, 3xTF32) against the
    float64 DFT of the same windowed frames.  The error is relative to the frame's strongest bin (as for any
    fixed-precision transform); two 3xTF32 stages carry ~22 bits each, measured 1.3e-6 of the peak amplitude -- about five
    times the error of the fp32 FFT of the product path."""
    from rho_tts_b200 import synth
    lens = [160000, 16000, 4001, 100001, 481000 if pad else 1000, 201 if not pad else 37]
    clips = [oracle.resample(synth.make_clip_block(1, (3 * L + 1) // 2, 70 + i)[0].numpy())[:L] for i, L in enumerate(lens)]
    rb = _rb(R, clips, cuda_device)
    power, base = R.stft_power_tc(rb, pad_to_30s=pad)
    power = power.cpu().numpy()
    hann = oracle.hann_periodic().astype(np.float64)
    worst_rel_peak, worst_amp = 0.0, 0.0
    for i, w in enumerate(clips):
        T = int(base[i + 1] - base[i])
        fr = _frames64(w, pad)[:T]
        assert fr.shape[0] == T, (i, len(w), T, fr.shape)
        if T == 0:
            continue
        X = np.fft.rfft(fr * hann[None, :], axis=1)
        want = np.abs(X) ** 2                                                        # (T, 201)
        got = power[base[i]:base[i + 1], :201].astype(np.float64)
        peak = np.maximum(want.max(axis=1, keepdims=True), 1e-30)
        worst_rel_peak = max(worst_rel_peak, float((np.abs(got - want) / peak).max()))
        worst_amp = max(worst_amp, float((np.abs(np.sqrt(got) - np.abs(X)) / np.sqrt(peak)).max()))
    print(f"stft_power_tc pad={pad}: worst |P - P64| / max_k P64 = {worst_rel_peak:.2e}, amplitude error / peak amplitude {worst_amp:.2e}")
    assert worst_rel_peak < 4e-6 and worst_amp < 2e-6


def test_log_mel_on_tensor_cores_only(R, cuda_device):
    """The whole Whisper front end of a clip on the tensor cores: rho_b200_stft_power_tc -> rho_b200_mel_project ->
    log10 / clamp / scale, against the numpy oracle and the product (FFT) path.  On bins 70..80 dB below a frame's
    peak the 3xTF32 DFT error (test above) shows up as ~2e-4 after the log: this path does NOT meet the 1e-4 contract,
    which is one of the two reasons it stays an isolated measurement (the other is speed: profiles/)."""
    from rho_tts_b200 import synth
    x = synth.make_clip_block(3, 240000, 11)
    clips16 = [oracle.resample(x[i].numpy()) for i in range(3)]
    rb = _rb(R, clips16, cuda_device)
    power, base = R.stft_power_tc(rb, pad_to_30s=False)
    prod, _ = R.logmel_batch(rb, 80, False)
    for i, w in enumerate(clips16):
        mel = R.mel_project(power[base[i]:base[i + 1]].contiguous(), 80).cpu().numpy()
        ls = np.log10(np.maximum(mel, 1e-10)); ls = np.maximum(ls, ls.max() - 8.0)
        got = (ls + 4.0) / 4.0
        want = oracle.log_mel(w, 80, pad_to_30s=False)
        T = want.shape[1]
        e_or = float(np.abs(got[:, :T] - want).max())
        e_pr = float(np.abs(got[:, :T] - prod[i, :, :T].cpu().numpy()).max())
        print(f"tensor-core-only log-mel, clip {i}: {e_or:.2e} from the numpy oracle, {e_pr:.2e} from the product path")
        assert e_or < 5e-4 and e_pr < 5e-4


# ----------------------------------------------------------------------------- error behaviour of the new entry points
def test_new_entry_points_fail_loudly(R, cuda_device):
    """Bad layouts / sizes are refused with a message (RuntimeError through the shim, never a silent fallback)."""
    import ctypes
    from rho_tts_b200 import synth, _lib
    x = synth.make_clip_block(4, 48000, 9)
    rb = R.RaggedBatch.from_dense(x.to(cuda_device))
    p = R.make_params()
    # compact rows shorter than the frames that can see signal
    plan = R.ValidatePlan(rb, np.arange(5, dtype=np.int32), p, 80, True, compact=True)
    plan.T_alloc = 100
    with pytest.raises(RuntimeError, match="compact rows need"):
        plan.run(rb, None, None)
    # host entry point: overlapping segments, outputs without room, mel rows too short
    xh = x.reshape(-1).pin_memory()
    off = np.array([0, 40000, 96000, 144000], np.int64)                     # segment 1 starts inside segment 0
    ln = np.full(4, 48000, np.int32)
    with pytest.raises(RuntimeError, match="non-overlapping"):
        R.validate_host_ragged(xh, off, ln, np.arange(5, dtype=np.int32), p, features=False)
    with pytest.raises(RuntimeError, match="item_first_seg"):
        R.validate_host_ragged(xh, np.arange(4, dtype=np.int64) * 48000, ln, np.array([0, 2, 3], np.int32), p, features=False)
    h = _lib.Handle.get(0)
    rec = torch.zeros((4, 48), dtype=torch.uint8).pin_memory()
    y = torch.empty(4 * 48000).pin_memory()
    mel = torch.empty((4, 80, 64)).pin_memory()
    seg_off = (np.arange(4, dtype=np.int64) * 48000)
    first = np.arange(5, dtype=np.int32)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)     # noqa: E731
    rc = h.lib.rho_b200_validate_host_ragged(h.ptr, ctypes.c_void_p(xh.data_ptr()), vp(seg_off), vp(ln), 4, vp(first), 4,
                                             ctypes.byref(p), ctypes.c_void_p(y.data_ptr()), vp(seg_off), 80, 3000,
                                             ctypes.c_void_p(mel.data_ptr()), 64, None, None, None, 0,
                                             ctypes.c_void_p(rec.data_ptr()))
    assert rc == -1 and "mel_stride_frames" in _lib.last_error()
    # the tensor-core DFT refuses a misaligned tile list and short rows
    pw = torch.empty((8, 208), device=cuda_device)
    tl = torch.zeros((2, 4), dtype=torch.int32, device=cuda_device)
    assert h.lib.rho_b200_stft_power_tc(h.ptr, ctypes.c_void_p(rb.data.data_ptr()), ctypes.c_void_p(rb.offsets.data_ptr()),
                                        ctypes.c_void_p(rb.lengths.data_ptr()), 0, ctypes.c_void_p(tl.data_ptr()), 1,
                                        ctypes.c_void_p(pw.data_ptr()), 100, None) == -1
    # the exchange cannot be waited on / read before it exists
    assert h.lib.rho_b200_exchange_wait(h.ptr, 1, None) == -1 and "not connected" in _lib.last_error()
