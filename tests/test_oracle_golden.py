"""CPU: the oracle against the committed golden vectors (outputs of the reference itself, of
torchaudio.functional.resample and of transformers.WhisperFeatureExtractor; tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle
from tests.util import TOL, assert_close

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
N_CLIPS, N_ITEMS = int(G["n_clips"]), int(G["n_items"])
C = oracle.derive_constants()


@pytest.mark.parametrize("i", range(N_CLIPS))
def test_trim_bounds_exact(i):
    x = G[f"clip{i}"]
    for tf in range(4):
        tr = oracle.trim_bounds(x, C, bool(tf & 1), bool(tf & 2))
        start, end, dim = (int(v) for v in G[f"trim{i}"][tf])
        assert (tr.start, tr.end) == (start, end), (i, tf)
        assert (2 if tr.all_silent else 1) == dim


@pytest.mark.parametrize("i", range(N_CLIPS))
def test_post_process_and_decay(i):
    o = oracle.post_process_clip(G[f"clip{i}"], C)
    want = G[f"post{i}"]
    assert o["out_len"] == want.size and (2 if o["all_silent"] else 1) == int(G[f"post_dim{i}"])
    assert_close(o["audio"], want, tol=1e-6, what=f"post {i}")
    ratio, ok = G[f"decay{i}"]
    assert o["ok"] == bool(ok) and abs(o["decay_ratio"] - ratio) <= 1e-5 * max(1.0, abs(ratio))


@pytest.mark.parametrize("k", range(N_ITEMS))
def test_join_items(k):
    segs = [G[f"clip{j}"] for j in G[f"item{k}_idx"]]
    o = oracle.smooth_segment_join(segs, C)
    want = G[f"item{k}"]
    assert o.audio.size == want.size, (k, o.fallback)
    assert (2 if o.two_d else 1) == int(G[f"item_dim{k}"])
    assert_close(o.audio, want, tol=1e-6, what=f"item {k}")
    ratio, ok = G[f"item_decay{k}"]
    r, okk, fr, _ = oracle.sound_decay(o.audio, 0.3)
    assert okk == bool(ok) and abs(r - ratio) <= 1e-5 * max(1.0, abs(ratio))


def test_golden_covers_fallback_and_two_d():
    dims = [int(G[f"item_dim{k}"]) for k in range(N_ITEMS)]
    fallbacks = [oracle.smooth_segment_join([G[f"clip{j}"] for j in G[f"item{k}_idx"]], C).fallback for k in range(N_ITEMS)]
    assert 2 in dims and any(fallbacks) and not all(fallbacks)
    assert not bool(G["decay8"][1])                 # the decaying clip is rejected
    assert bool(G["decay0"][1])


@pytest.mark.parametrize("i", range(4))
def test_resample(i):
    want = G[f"rs{i}"]
    got = oracle.resample(G[f"post{i}"])
    assert got.size == want.size == -(-2 * G[f"post{i}"].size // 3)
    assert_close(got, want, tol=1e-6, what=f"resample {i}")


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("i", [0, 2])
def test_log_mel(n_mels, i):
    w = G[f"rs{i}"]
    pad = oracle.log_mel(w, n_mels, True)
    head = G[f"mel{n_mels}_{i}_pad_head"]
    assert pad.shape == (n_mels, 3000)
    # 1e-4 is the contract; the library's own fp32 STFT sits up to ~9e-5 from the fp64 value on -70 dB
    # bins (DESIGN.md "log-mel accuracy"), the oracle's pocketfft is ~1e-5 from it.
    assert_close(pad[:, :head.shape[1]], head, tol=TOL, what="log-mel padded head")
    assert_close(pad[:, head.shape[1]:], np.full((n_mels, 3000 - head.shape[1]), G[f"mel{n_mels}_{i}_pad_fill"]),
                 tol=1e-6, what="log-mel padding fill")
    nop = oracle.log_mel(w, n_mels, False)
    assert_close(nop, G[f"mel{n_mels}_{i}_nopad"], tol=TOL, what="log-mel unpadded")


def test_cosine():
    e = G["emb"]
    got = np.asarray([oracle.cosine_similarity(e[0], g) for g in e[1:]], dtype=np.float32)
    assert_close(got, G["cos"], tol=1e-6, what="cosine")


# ----------------------------------------------------------------------------- crossfade disabled (cf == 0, cf <= 10)
GCF = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cf0_v1.npz"))


@pytest.mark.parametrize("v", range(2))
def test_join_with_crossfade_disabled(v):
    """crossfade_duration_sec = 0: the reference's [..., :-0] slice drops segment 0 (base_tts.py:485); 0.0004 s gives
    9 samples <= 10, so no crossfade but segment 0 loses its last 9 samples.  Vectors from the reference itself."""
    c = oracle.derive_constants(xfade_sec=float(GCF["xfade_secs"][v]))
    assert c.cf == (0, 9)[v]
    for k in range(int(GCF["n_items"])):
        segs = [G[f"clip{j}"] for j in GCF[f"v{v}_item{k}_idx"]]
        o = oracle.smooth_segment_join(segs, c)
        want = GCF[f"v{v}_item{k}"]
        assert o.audio.size == want.size, (v, k, o.fallback)
        assert (2 if o.two_d else 1) == int(GCF[f"v{v}_item_dim{k}"])
        assert_close(o.audio, want, tol=1e-6, what=f"cf0 v{v} item {k}")
