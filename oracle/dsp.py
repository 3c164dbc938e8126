"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the in-tree DSP stages.

Follows, stage by stage, the reference at
  /root/reference/src/rho_tts/base_tts.py
    _validate_sound_decay          :297-323
    _compute_speaker_similarity    :341-344   (cosine part only)
    _trim_silence                  :348-392
    _remove_dc_offset              :394-399
    _apply_fades                   :401-433
    _smooth_segment_join           :435-536
Mono float32 only (providers emit mono: qwen.py:265, chatterbox.py:167).

Numerical contract that the CUDA path is checked against:
  * trim start/end, join lengths/piece boundaries, decay accept/reject: exact.
  * waveforms, RMS, ratio, cosine: |a-b| <= 1e-4 * max(1, |b|).

Nothing here is imported by the product package.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# constants derived at call time (base_tts.py:366-367, :420, :455, :519)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Consts:
    sr: int
    window: int      # int(sr * 0.01)
    hop: int         # window // 2   (avg_pool1d stride)
    pad: int         # window // 2   (avg_pool1d padding)
    thr: np.float32  # 10 ** (dB / 20), compared in fp32
    fade: int        # int(sr * fade_duration_sec)
    cf: int          # int(sr * crossfade_duration_sec)
    pause: int       # int(sr * inter_sentence_pause_sec)
    pause_on: bool   # inter_sentence_pause_sec > 0


def derive_constants(sr: int = 24000, silence_db: float = -50.0, fade_sec: float = 0.02,
                     xfade_sec: float = 0.05, pause_sec: float = 0.1) -> Consts:
    window = int(sr * 0.01)
    return Consts(
        sr=int(sr), window=window, hop=window // 2, pad=window // 2,
        thr=F32(10 ** (silence_db / 20)),
        fade=int(sr * fade_sec), cf=int(sr * xfade_sec),
        pause=int(sr * pause_sec), pause_on=pause_sec > 0,
    )


# --------------------------------------------------------------------------
# silence trim  (base_tts.py:348-392)
# --------------------------------------------------------------------------
def frame_energy(x: np.ndarray, c: Consts) -> np.ndarray:
    """RMS per 10 ms window / 5 ms hop, as avg_pool1d(x**2, k=window, s=hop, p=pad)
    then sqrt (base_tts.py:369-375).

    torch's CPU avg_pool kernel adds the `window` squared samples of one output
    frame one after another in fp32, zero padding included, and divides by
    `window` (count_include_pad).  np.add.accumulate along a row is the same
    left-to-right fp32 chain, so the values are bit-identical
    (tests/test_oracle_vs_reference.py checks this against torch itself).
    """
    x = np.ascontiguousarray(x, dtype=F32).reshape(-1)
    L = x.shape[0]
    n_frames = (L + 2 * c.pad - c.window) // c.hop + 1
    if n_frames <= 0:
        return np.zeros(0, dtype=F32)
    sq = np.zeros(L + 2 * c.pad, dtype=F32)
    np.multiply(x, x, out=sq[c.pad:c.pad + L])
    frames = np.lib.stride_tricks.sliding_window_view(sq, c.window)[::c.hop][:n_frames]
    run = np.add.accumulate(frames, axis=1, dtype=F32)[:, -1]
    mean_sq = run / F32(c.window)
    return np.sqrt(mean_sq, dtype=F32)


@dataclass
class TrimResult:
    start: int
    end: int
    all_silent: bool   # reference returns a 2-D (1, min(window, L)) view in this case
    untouched: bool    # trimming disabled or empty input: input object returned
    first_frame: int = -1
    last_frame: int = -1

    @property
    def length(self) -> int:
        return self.end - self.start


def trim_bounds(x: np.ndarray, c: Consts, from_start: bool = True, from_end: bool = True,
                enabled: bool = True) -> TrimResult:
    L = int(np.asarray(x).size)
    if not enabled or L == 0:
        return TrimResult(0, L, False, True)
    e = frame_energy(x, c)
    loud = np.flatnonzero(e > c.thr)
    if loud.size == 0:
        return TrimResult(0, min(c.window, L), True, False)
    first, last = int(loud[0]), int(loud[-1])
    start = (first * c.window // 2) if from_start else 0
    end = ((last + 2) * c.window // 2) if from_end else L
    start = max(0, min(start, L))
    end = max(start, min(end, L))
    return TrimResult(start, end, False, False, first, last)


def trim_silence(x: np.ndarray, c: Consts, from_start: bool = True, from_end: bool = True,
                 enabled: bool = True) -> Tuple[np.ndarray, TrimResult]:
    x = np.asarray(x, dtype=F32).reshape(-1)
    r = trim_bounds(x, c, from_start, from_end, enabled)
    return x[r.start:r.end], r


# --------------------------------------------------------------------------
# DC removal (base_tts.py:394-399) and fades (:401-433)
# --------------------------------------------------------------------------
def remove_dc_offset(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=F32)
    if x.size == 0:
        return x
    return (x - x.mean(dtype=F32)).astype(F32)


def _linspace32(a: float, b: float, n: int) -> np.ndarray:
    """fp32 linspace the way torch builds it: step in fp32, first half counted up
    from `a`, second half counted down from `b`."""
    if n == 0:
        return np.zeros(0, dtype=F32)
    if n == 1:
        return np.array([a], dtype=F32)
    a32, b32 = F32(a), F32(b)
    step = F32((b32 - a32) / F32(n - 1))
    i = np.arange(n)
    up = (a32 + step * i.astype(F32)).astype(F32)
    down = (b32 - step * (n - 1 - i).astype(F32)).astype(F32)
    return np.where(i < n // 2, up, down).astype(F32)


def fade_curves(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """(fade_in, fade_out) raised-cosine curves of length n (base_tts.py:426, :430)."""
    ramp = _linspace32(0.0, np.pi, n)
    cs = np.cos(ramp, dtype=F32)
    return (F32(0.5) * (F32(1) - cs)).astype(F32), (F32(0.5) * (F32(1) + cs)).astype(F32)


def apply_fades(x: np.ndarray, c: Consts, fade_in: bool = True, fade_out: bool = True) -> np.ndarray:
    """Returns a faded copy (the reference mutates in place; aliasing is the host
    shim's business, the values are what is checked here)."""
    y = np.array(x, dtype=F32, copy=True).reshape(-1)
    n = c.fade
    if y.size == 0 or y.size < 2 * n or n == 0:
        return y
    fi, fo = fade_curves(n)
    if fade_in:
        y[:n] = y[:n] * fi
    if fade_out:
        y[-n:] = y[-n:] * fo
    return y


# --------------------------------------------------------------------------
# crossfade join (base_tts.py:435-536)
# --------------------------------------------------------------------------
def crossfade_curves(ov: int) -> Tuple[np.ndarray, np.ndarray]:
    """(fade_out for the previous tail, fade_in for the current head), :497-498."""
    fo = np.cos(_linspace32(0.0, np.pi / 2, ov), dtype=F32)
    fi = np.cos(_linspace32(np.pi / 2, 0.0, ov), dtype=F32)
    return fo, fi


# piece kinds of a join plan
P_COPY = 0    # y[dst:dst+n] = seg[src:src+n]
P_XFADE = 1   # y[dst+j] = prev[src_prev+j]*fo[j] + cur[src+j]*fi[j]
P_ZERO = 2    # pause


@dataclass
class Piece:
    kind: int
    dst: int
    n: int
    seg: int = -1        # current segment index within the item
    src: int = 0         # offset inside the *processed* (trimmed) segment
    prev_src: int = 0    # P_XFADE: offset inside the previous processed segment
    two_d: bool = False  # dimensionality the reference's tensor would have


@dataclass
class JoinPlan:
    n_seg: int
    fallback: bool                 # reference's except-branch: cat of the ORIGINAL segments
    out_len: int
    two_d: bool                    # result is (1, out_len) instead of (out_len,)
    pieces: List[Piece] = field(default_factory=list)
    trims: List[TrimResult] = field(default_factory=list)


def join_plan(seg_lens: Sequence[int], trims: Sequence[TrimResult], c: Consts) -> JoinPlan:
    """Piece list for one multi-segment item (N >= 2), given per-segment trim results.

    Mirrors the emit order of base_tts.py:481-523 and models tensor
    dimensionality so the torch.cat failure (-> fallback, :530-533) is reproduced:
    an all-silent segment is 2-D, everything else 1-D, the pause is always 1-D,
    a crossfade piece is 2-D when either side is 2-D.
    """
    n = len(seg_lens)
    assert n >= 2 and len(trims) == n
    Ls = [t.length for t in trims]
    dims2 = [t.all_silent for t in trims]
    pieces: List[Piece] = []
    pos = 0

    def emit(kind, cnt, **kw):
        nonlocal pos
        pieces.append(Piece(kind=kind, dst=pos, n=cnt, **kw))
        pos += cnt

    cf = c.cf
    # segment 0 (:484-488)
    if Ls[0] > cf:
        # current_segment[..., :-crossfade_samples]: with crossfade_samples == 0 this is [..., :0], an EMPTY slice --
        # the reference drops segment 0 when the crossfade is disabled (kept bug for bug)
        emit(P_COPY, Ls[0] - cf if cf > 0 else 0, seg=0, src=0, two_d=dims2[0])
    else:
        emit(P_COPY, Ls[0], seg=0, src=0, two_d=dims2[0])
    for i in range(1, n):
        ov = min(cf, Ls[i - 1], Ls[i])
        if ov > 10:
            # prev_tail.dim()==2 adds a leading axis to the curves (:500-502); either
            # way the product broadcasts and the piece is 2-D if any side is 2-D.
            emit(P_XFADE, ov, seg=i, src=0, prev_src=Ls[i - 1] - ov,
                 two_d=dims2[i - 1] or dims2[i])
            if i < n - 1 and Ls[i] > ov + cf:
                rem = Ls[i] - ov - cf
            else:
                rem = Ls[i] - ov
            if rem > 0:
                emit(P_COPY, rem, seg=i, src=ov, two_d=dims2[i])
            if c.pause_on and i < n - 1:
                emit(P_ZERO, c.pause, two_d=False)
        else:
            emit(P_COPY, Ls[i], seg=i, src=0, two_d=dims2[i])
    # torch.cat accepts the list only if every (non-legacy-empty) piece has one rank.
    # A 1-D tensor of 0 elements is skipped by cat's legacy rule; 2-D ones never have 0 here.
    ranks = {p.two_d for p in pieces if not (p.n == 0 and not p.two_d)}
    if len(ranks) > 1:
        return JoinPlan(n, True, int(sum(seg_lens)), False, [], list(trims))
    two_d = bool(ranks.pop()) if ranks else False
    return JoinPlan(n, False, pos, two_d, pieces, list(trims))


@dataclass
class JoinResult:
    audio: Optional[np.ndarray]
    two_d: bool
    fallback: bool
    plan: Optional[JoinPlan]
    dc: List[float] = field(default_factory=list)


def smooth_segment_join(segments: Sequence[np.ndarray], c: Consts, trim_enabled: bool = True) -> JoinResult:
    n = len(segments)
    if n == 0:
        return JoinResult(None, False, False, None)
    segs = [np.asarray(s, dtype=F32).reshape(-1) for s in segments]
    if n == 1:
        y, tr = trim_silence(segs[0], c, True, True, trim_enabled)
        dc = float(y.mean(dtype=F32)) if y.size else 0.0
        y = remove_dc_offset(y)
        y = apply_fades(y, c, True, True)
        plan = JoinPlan(1, False, int(y.size), tr.all_silent, [], [tr])
        return JoinResult(y, tr.all_silent, False, plan, [dc])

    trims: List[TrimResult] = []
    proc: List[np.ndarray] = []
    dcs: List[float] = []
    for i, s in enumerate(segs):
        if i == 0:
            t, tr = trim_silence(s, c, False, True, trim_enabled)
        elif i == n - 1:
            t, tr = trim_silence(s, c, True, False, trim_enabled)
        else:
            t, tr = trim_silence(s, c, True, True, trim_enabled)
        trims.append(tr)
        dcs.append(float(t.mean(dtype=F32)) if t.size else 0.0)
        proc.append(remove_dc_offset(t))

    plan = join_plan([s.size for s in segs], trims, c)
    if plan.fallback:
        y = np.concatenate(segs).astype(F32)
        return JoinResult(apply_fades(y, c, True, True), False, True, plan, dcs)

    y = np.zeros(plan.out_len, dtype=F32)
    for p in plan.pieces:
        if p.n == 0:
            continue
        if p.kind == P_COPY:
            y[p.dst:p.dst + p.n] = proc[p.seg][p.src:p.src + p.n]
        elif p.kind == P_XFADE:
            fo, fi = crossfade_curves(p.n)
            prev = proc[p.seg - 1][p.prev_src:p.prev_src + p.n]
            cur = proc[p.seg][p.src:p.src + p.n]
            y[p.dst:p.dst + p.n] = prev * fo + cur * fi
    # _apply_fades on a (1, L) tensor squeezes, fades, and views back (:416-433)
    y = apply_fades(y, c, True, True)
    return JoinResult(y, plan.two_d, False, plan, dcs)


# --------------------------------------------------------------------------
# sound decay (base_tts.py:297-323) and cosine (:341-344)
# --------------------------------------------------------------------------
def sound_decay(x: np.ndarray, threshold: float = 0.3) -> Tuple[float, bool, float, float]:
    """(ratio, ok, first_rms, last_rms).  RMS in fp32, ratio and compare in double."""
    flat = np.asarray(x, dtype=F32).reshape(-1)
    n = flat.size
    if n == 0:
        return 1.0, True, 0.0, 0.0
    third = n // 3
    if third < 1:
        return 1.0, True, 0.0, 0.0
    first = float(np.sqrt(np.mean(flat[:third] ** 2, dtype=F32), dtype=F32))
    last = float(np.sqrt(np.mean(flat[n - third:] ** 2, dtype=F32), dtype=F32))
    if first < 1e-8:
        return 1.0, True, first, last
    ratio = last / first
    return ratio, bool(ratio >= threshold), first, last


def cosine_similarity(ref: np.ndarray, gen: np.ndarray) -> np.float32:
    ref = np.asarray(ref, dtype=F32)
    gen = np.asarray(gen, dtype=F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.dot(ref, gen) / (np.linalg.norm(ref) * np.linalg.norm(gen))


# --------------------------------------------------------------------------
# the per-clip pipeline as _run_pipeline calls it (base_tts.py:912-926)
# --------------------------------------------------------------------------
def post_process_clip(x: np.ndarray, c: Consts, decay_threshold: float = 0.3,
                      trim_enabled: bool = True) -> dict:
    """_smooth_segment_join([x]) -> _post_process_audio (base: identity) -> _validate_sound_decay."""
    jr = smooth_segment_join([x], c, trim_enabled)
    tr = jr.plan.trims[0]
    ratio, ok, first, last = sound_decay(jr.audio, decay_threshold)
    return dict(audio=jr.audio, start=tr.start, end=tr.end, out_len=int(jr.audio.size),
                all_silent=tr.all_silent, dc=jr.dc[0], first_rms=first, last_rms=last,
                decay_ratio=ratio, ok=ok)
