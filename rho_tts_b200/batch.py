"""Batched host API over librho_b200: what the benchmarks and the BaseTTS shim drive.

Every function takes device-resident torch tensors (PyTorch is the allocator and stream
provider, nothing more), enqueues the library's kernels on the current CUDA stream and returns
device tensors; nothing here synchronises or computes on the CPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import RhoParams, Handle
from .ragged import RaggedBatch, ALIGN

REC_DTYPE = np.dtype([
    ("start", "<i4"), ("end", "<i4"), ("out_len", "<i4"), ("flags", "<u4"),
    ("dc", "<f4"), ("first_rms", "<f4"), ("last_rms", "<f4"), ("cosine", "<f4"),
    ("decay_ratio", "<f8"), ("ok", "<i4"), ("n_segments", "<i4"),
])
SEG_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("dc", "<f4"), ("flags", "<u4")])
assert REC_DTYPE.itemsize == 48 and SEG_DTYPE.itemsize == 16


def make_params(sr: int = 24000, trim_silence: bool = True, silence_threshold_db: float = -50.0,
                fade_duration_sec: float = 0.02, crossfade_duration_sec: float = 0.05,
                inter_sentence_pause_sec: float = 0.1, sound_decay_threshold: float = 0.3) -> RhoParams:
    return RhoParams(int(sr), 1 if trim_silence else 0, float(silence_threshold_db), float(fade_duration_sec),
                     float(crossfade_duration_sec), float(inter_sentence_pause_sec), float(sound_decay_threshold))


def params_from_tts(tts) -> RhoParams:
    """Read the BaseTTS attributes at call time, as the reference methods do."""
    return make_params(
        sr=tts.sample_rate,
        trim_silence=getattr(tts, "trim_silence", True),
        silence_threshold_db=getattr(tts, "silence_threshold_db", -50.0),
        fade_duration_sec=getattr(tts, "fade_duration_sec", 0.02),
        crossfade_duration_sec=getattr(tts, "crossfade_duration_sec", 0.05),
        inter_sentence_pause_sec=getattr(tts, "inter_sentence_pause_sec", 0.1),
        sound_decay_threshold=getattr(tts, "sound_decay_threshold", 0.3),
    )


def _dev_index(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("rho_tts_b200 runs on CUDA tensors only (no CPU fallback)")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(dev: int) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _workspace(h: Handle, n_seg: int, n_items: int, max_seg_len: int, device) -> torch.Tensor:
    nbytes = int(h.lib.rho_b200_workspace_bytes(int(n_seg), int(n_items), int(max_seg_len)))
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


@dataclass
class JoinOutput:
    audio: RaggedBatch          # item i at audio.h_offsets[i]; valid length = records["out_len"][i] (device)
    records: torch.Tensor       # uint8 [n_items, 48] on the device (rho_record)
    seg_info: Optional[torch.Tensor]   # uint8 [n_segments, 16] on the device (rho_seg_info)

    def records_host(self) -> np.ndarray:
        return self.records.cpu().numpy().view(REC_DTYPE).reshape(-1)

    def seg_info_host(self) -> np.ndarray:
        return self.seg_info.cpu().numpy().view(SEG_DTYPE).reshape(-1)


def trim_scan_batch(rb: RaggedBatch, p: RhoParams, trim_flags: Optional[torch.Tensor] = None) -> torch.Tensor:
    """BaseTTS._trim_silence bounds for every clip (base_tts.py:348-392). Returns uint8 [n,16] rho_seg_info."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    info = torch.empty((rb.n, 16), dtype=torch.uint8, device=rb.device)
    ws = _workspace(h, rb.n, rb.n, rb.max_len, rb.device)
    _lib.check(h.lib.rho_b200_trim_scan(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(rb.lengths), _ptr(trim_flags),
                                        rb.n, rb.max_len, ctypes.byref(p), _ptr(info), _ptr(ws), ws.numel(),
                                        _stream(dev)), "trim_scan")
    return info


def join_batch(rb: RaggedBatch, item_first_seg: Sequence[int], p: RhoParams, want_seg_info: bool = True) -> JoinOutput:
    """BaseTTS._smooth_segment_join + _validate_sound_decay for many items at once
    (base_tts.py:435-536, 297-323).  item i = segments [item_first_seg[i], item_first_seg[i+1])."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    first = np.asarray(item_first_seg, dtype=np.int32)
    n_items = len(first) - 1
    assert n_items >= 0 and first[0] == 0 and first[-1] == rb.n
    pause = int(p.sr * p.pause_sec) if p.pause_sec > 0 else 0
    seg_tot = np.concatenate([[0], np.cumsum(rb.h_lengths.astype(np.int64))])
    cap = (seg_tot[first[1:]] - seg_tot[first[:-1]]) + np.maximum(0, np.diff(first) - 2) * pause
    out = RaggedBatch.empty_like_lengths(np.maximum(cap, 0).astype(np.int32), rb.device) if n_items else \
        RaggedBatch.empty_like_lengths(np.zeros(0, np.int32), rb.device)
    rec = torch.empty((n_items, 48), dtype=torch.uint8, device=rb.device)
    seg = torch.empty((rb.n, 16), dtype=torch.uint8, device=rb.device) if want_seg_info else None
    d_first = torch.from_numpy(first).to(rb.device)
    ws = _workspace(h, rb.n, n_items, rb.max_len, rb.device)
    _lib.check(h.lib.rho_b200_join(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(rb.lengths), rb.n, rb.max_len,
                                   _ptr(d_first), n_items, int(cap.max()) if n_items else 0, ctypes.byref(p),
                                   _ptr(out.data), _ptr(out.offsets), _ptr(rec), _ptr(seg), _ptr(ws), ws.numel(),
                                   _stream(dev)), "join")
    return JoinOutput(out, rec, seg)


def post_process_batch(rb: RaggedBatch, p: RhoParams, want_seg_info: bool = False) -> JoinOutput:
    """Every clip is its own item: trim both ends -> DC -> fades -> decay record
    (the N == 1 branch of _smooth_segment_join, base_tts.py:447-452)."""
    return join_batch(rb, np.arange(rb.n + 1, dtype=np.int32), p, want_seg_info)


def resample_batch(rb: RaggedBatch, lengths: Optional[torch.Tensor] = None, len_stride: int = 4):
    """torchaudio.functional.resample(x, 24000, 16000) per clip.  `lengths` (device int32, optional,
    strided by len_stride bytes) overrides rb.lengths, e.g. the out_len column of the records."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    cap = (2 * rb.h_lengths.astype(np.int64) + 2) // 3
    out = RaggedBatch.empty_like_lengths(cap.astype(np.int32), rb.device)
    lens = rb.lengths if lengths is None else lengths
    _lib.check(h.lib.rho_b200_resample3to2(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(lens), int(len_stride), rb.n,
                                           rb.max_len, _ptr(out.data), _ptr(out.offsets), _ptr(out.lengths),
                                           _stream(dev)), "resample3to2")
    return out          # out.lengths (device) now holds ceil(2*len/3)


def resample_any_batch(rb: RaggedBatch, orig_freq: int, new_freq: int) -> RaggedBatch:
    """torchaudio.functional.resample(x, orig_freq, new_freq) per clip, any ratio -- the speed control of
    BaseTTS._apply_speed_pitch (base_tts.py:631-637)."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    if int(orig_freq) == int(new_freq):
        return rb
    cap = np.asarray([int(h.lib.rho_b200_resample_out_len(int(L), int(orig_freq), int(new_freq)))
                      for L in rb.h_lengths], dtype=np.int32)
    out = RaggedBatch.empty_like_lengths(cap, rb.device)
    _lib.check(h.lib.rho_b200_resample(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(rb.lengths), 4, rb.n, rb.max_len,
                                       int(orig_freq), int(new_freq), _ptr(out.data), _ptr(out.offsets),
                                       _ptr(out.lengths), _stream(dev)), "resample")
    return out


TORCH_ARANGE_LANES = 8      # lanes of the vectorised torch.arange CPU kernel in the torch 2.11 wheels (oracle/pitch.py)


def pitch_shift_batch(rb: RaggedBatch, sample_rate: int, n_steps: float,
                      arange_lanes: int = TORCH_ARANGE_LANES) -> RaggedBatch:
    """torchaudio.functional.pitch_shift(x, sample_rate, n_steps) per clip (same lengths out) -- the pitch control of
    BaseTTS._apply_speed_pitch (base_tts.py:639-648).  Clips of <= 256 samples are refused like torch.stft's reflect
    padding refuses them."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    if float(n_steps) == 0.0 or rb.n == 0:
        return rb
    out = RaggedBatch.empty_like_lengths(np.asarray(rb.h_lengths, dtype=np.int32), rb.device)
    need = int(h.lib.rho_b200_pitch_workspace_bytes(rb.n, rb.max_len, float(n_steps)))
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=rb.device)
    _lib.check(h.lib.rho_b200_pitch_shift(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(rb.lengths), 4, rb.n,
                                          int(min(rb.h_lengths)), rb.max_len, int(sample_rate), float(n_steps),
                                          int(arange_lanes), _ptr(out.data), _ptr(out.offsets), _ptr(ws), ws.numel(),
                                          _stream(dev)), "pitch_shift")
    return out


def pcm16_batch(rb: RaggedBatch, lengths: Optional[torch.Tensor] = None, len_stride: int = 4) -> torch.Tensor:
    """int16 samples of every clip at the clip's offsets (one int16 tensor as long as rb.data): the payload
    BaseTTS._save_wav's in-tree writer puts into the WAV file (base_tts.py:661-667)."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    out = torch.zeros(rb.data.numel(), dtype=torch.int16, device=rb.device)
    if rb.n == 0:
        return out
    lens = rb.lengths if lengths is None else lengths
    _lib.check(h.lib.rho_b200_pcm16(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(lens), int(len_stride), rb.n, rb.max_len,
                                    _ptr(out), _ptr(rb.offsets), _stream(dev)), "pcm16")
    return out


def write_wav(path: str, audio: torch.Tensor, sample_rate: int, device=None) -> None:
    """The file BaseTTS._save_wav's in-tree writer produces (mono 16-bit PCM, base_tts.py:661-667), with the
    float -> int16 conversion on the B200 and half the bytes crossing PCIe."""
    import wave
    flat = audio.detach().reshape(-1)
    dev = flat.device if flat.is_cuda else torch.device("cuda", 0 if device is None else int(device))
    rb = RaggedBatch.from_list([flat.to(dev, torch.float32)], dev)
    pcm = pcm16_batch(rb)[int(rb.h_offsets[0]):int(rb.h_offsets[0]) + flat.numel()].cpu().numpy()
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(int(sample_rate))
        wf.writeframes(pcm.tobytes())


def mfcc_stats_batch(rb16: RaggedBatch, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n, 26] fp32: mean and std over the frames of librosa.feature.mfcc(y, sr=16000, n_mfcc=13) per 16 kHz clip -- the
    MFCC part of the drift classifier's feature vector (validation/classifier/trainer.py:50-52)."""
    dev = _dev_index(rb16.data)
    h = Handle.get(dev)
    n = rb16.n
    out = torch.empty((n, 26), dtype=torch.float32, device=rb16.device)
    if n == 0:
        return out
    if lengths is None and int(min(rb16.h_lengths)) < 1:
        raise RuntimeError("rho_tts_b200.mfcc_stats_batch: empty clip")
    need = int(h.lib.rho_b200_mfcc_workspace_bytes(n, rb16.max_len))
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=rb16.device)
    lens = rb16.lengths if lengths is None else lengths
    _lib.check(h.lib.rho_b200_mfcc_stats(h.ptr, _ptr(rb16.data), _ptr(rb16.offsets), _ptr(lens), 4, n, rb16.max_len,
                                         _ptr(out), _ptr(ws), ws.numel(), _stream(dev)), "mfcc_stats")
    return out


def logmel_batch(rb16: RaggedBatch, n_mels: int = 80, pad_to_30s: bool = True,
                 lengths: Optional[torch.Tensor] = None):
    """WhisperFeatureExtractor features.  Returns (mel [n, n_mels, T], n_frames int32 [n]) on the device;
    T = 3000 when padded, else max_len16 // 160 (clip i valid for n_frames[i])."""
    dev = _dev_index(rb16.data)
    h = Handle.get(dev)
    n = rb16.n
    T = 3000 if pad_to_30s else max(rb16.max_len // 160, 1)
    T_alloc = (T + 3) // 4 * 4            # rows 16-byte aligned: the normaliser's 128-bit path
    mel_buf = torch.empty((n, n_mels, T_alloc), dtype=torch.float32, device=rb16.device)
    mel = mel_buf[:, :, :T]
    n_frames = torch.zeros(n, dtype=torch.int32, device=rb16.device)
    ws = torch.empty(max(256, 4 * n + 256), dtype=torch.uint8, device=rb16.device)
    lens = rb16.lengths if lengths is None else lengths
    _lib.check(h.lib.rho_b200_logmel(h.ptr, _ptr(rb16.data), _ptr(rb16.offsets), _ptr(lens), n, rb16.max_len,
                                     int(n_mels), 3000 if pad_to_30s else 0, _ptr(mel_buf), T_alloc, _ptr(n_frames),
                                     _ptr(ws), ws.numel(), _stream(dev)), "logmel")
    return mel, n_frames


def mel_project(power: torch.Tensor, n_mels: int = 80, frames_per_item: int = 0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`mel_filters.T @ magnitudes` (feature_extraction_whisper.py:159) as a tensor-core GEMM (tcgen05, 3xTF32).
    power: [n_frames, ld >= 201] fp32 on the device, ld % 4 == 0 (columns >= 201 are ignored).
    frames_per_item = 0: returns mel energies [n_mels, n_frames].
    frames_per_item = T: the frames are n_frames / T consecutive clips; returns [n_items, n_mels, T]
    (or fills `out` [n_items, n_mels, T_out >= T], e.g. the 3000-frame Whisper layout)."""
    dev = _dev_index(power)
    h = Handle.get(dev)
    if power.dim() != 2 or power.dtype != torch.float32 or not power.is_contiguous():
        raise RuntimeError("mel_project: power must be a contiguous fp32 [n_frames, ld] tensor")
    n_frames, ld = power.shape
    if frames_per_item <= 0:
        ld_out = (n_frames + 3) // 4 * 4
        mel = torch.empty((n_mels, ld_out), dtype=torch.float32, device=power.device)
        _lib.check(h.lib.rho_b200_mel_project(h.ptr, _ptr(power), n_frames, ld, int(n_mels), _ptr(mel), ld_out, 0, 0,
                                              _stream(dev)), "mel_project")
        return mel[:, :n_frames]
    n_items = (n_frames + frames_per_item - 1) // frames_per_item
    if out is None:
        out = torch.empty((n_items, n_mels, frames_per_item), dtype=torch.float32, device=power.device)
    if out.dim() != 3 or out.shape[0] < n_items or out.shape[1] != n_mels or not out.is_contiguous():
        raise RuntimeError("mel_project: out must be a contiguous [n_items, n_mels, T_out] tensor")
    _lib.check(h.lib.rho_b200_mel_project(h.ptr, _ptr(power), n_frames, ld, int(n_mels), _ptr(out), out.shape[2],
                                          frames_per_item, out.shape[1] * out.shape[2], _stream(dev)), "mel_project")
    return out


def stft_power_tc(rb16: RaggedBatch, pad_to_30s: bool = True):
    """|STFT|^2 of the Whisper front end (feature_extraction_whisper.py:149-158) on the tensor cores: the windowed DFT
    factored 400 = 25 x 16 into two tcgen05 GEMMs (rho_b200_stft_power_tc).  Returns (power [n_frames_total, 208] fp32
    on the device -- columns 0..200 are the bins, the layout rho_b200_mel_project reads --, frame_base int64 [n + 1]:
    clip i owns rows frame_base[i] .. frame_base[i + 1]).  Padded: the frames of the 30 s window that see signal."""
    dev = _dev_index(rb16.data)
    h = Handle.get(dev)
    n16 = rb16.h_lengths.astype(np.int64)
    if pad_to_30s:
        nv = np.minimum(n16, 480000)
        T = np.minimum(3000, np.maximum(2, (nv + 200 + 159) // 160))
    else:
        T = np.where(n16 > 200, n16 // 160, 0)
    base = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    tiles = []
    for c in range(rb16.n):
        t0 = np.arange(0, int(T[c]), 8, dtype=np.int64)
        if t0.size:
            tiles.append(np.stack([np.full_like(t0, c), t0, np.minimum(8, int(T[c]) - t0), base[c] + t0], axis=1))
    power = torch.zeros((max(int(base[-1]), 1), 208), dtype=torch.float32, device=rb16.device)
    if tiles:
        tl = torch.from_numpy(np.concatenate(tiles).astype(np.int32)).to(rb16.device)
        _lib.check(h.lib.rho_b200_stft_power_tc(h.ptr, _ptr(rb16.data), _ptr(rb16.offsets), _ptr(rb16.lengths),
                                                3000 if pad_to_30s else 0, _ptr(tl), int(tl.shape[0]), _ptr(power), 208,
                                                _stream(dev)), "stft_power_tc")
    return power[:int(base[-1])], base


def qwen_post_process_batch(rb: RaggedBatch, sr: int = 24000, lengths: Optional[torch.Tensor] = None,
                            len_stride: int = 4, in_place: bool = False) -> RaggedBatch:
    """QwenTTS._post_process_audio (providers/qwen.py:268-378) for every clip of the batch: windowed decay
    correction, -23 dBFS, tanh soft clip.  `lengths` (device int32, strided by len_stride bytes) overrides
    rb.lengths, e.g. the out_len column of the records of a join.  Returns a batch with the same layout."""
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    out = rb if in_place else RaggedBatch(torch.empty_like(rb.data), rb.offsets, rb.lengths, rb.h_offsets, rb.h_lengths)
    nbytes = int(h.lib.rho_b200_qwen_workspace_bytes(rb.n, rb.max_len, int(sr)))
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=rb.device)
    lens = rb.lengths if lengths is None else lengths
    _lib.check(h.lib.rho_b200_qwen_postprocess(h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(lens), int(len_stride),
                                               rb.n, rb.max_len, int(sr), _ptr(out.data), _ptr(out.offsets),
                                               _ptr(ws), ws.numel(), _stream(dev)), "qwen_postprocess")
    return out


def qwen_pipeline_batch(rb: RaggedBatch, item_first_seg: Sequence[int], p: RhoParams, qwen3_sr: int = 24000) -> JoinOutput:
    """What BaseTTS._run_pipeline does to the segments of every item of a Qwen provider (base_tts.py:911-926):
    _smooth_segment_join -> QwenTTS._post_process_audio -> _validate_sound_decay ON THE HOOKED AUDIO.
    Returns the join's output with the audio replaced by the hook's and the decay fields of the records recomputed."""
    out = join_batch(rb, item_first_seg, p, want_seg_info=False)
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    n = out.audio.n
    if n == 0:
        return out
    lens = out.records[:, 8:12].contiguous().view(torch.int32).reshape(-1)        # out_len of every item (device)
    qwen_post_process_batch(out.audio, qwen3_sr, lengths=lens, in_place=True)
    ws = torch.empty(max(16 * n, 256), dtype=torch.uint8, device=rb.device)
    _lib.check(h.lib.rho_b200_sound_decay_batch(h.ptr, _ptr(out.audio.data), _ptr(out.audio.offsets), _ptr(lens), 4, n,
                                                out.audio.max_len, ctypes.byref(p), _ptr(out.records), _ptr(ws),
                                                ws.numel(), _stream(dev)), "sound_decay_batch")
    return out


def qwen_validate_batch(rb: RaggedBatch, item_first_seg: Sequence[int], p: RhoParams,
                        emb: Optional[torch.Tensor] = None, ref: Optional[torch.Tensor] = None, n_mels: int = 80,
                        pad_to_30s: bool = True, qwen3_sr: int = 24000) -> "ValidateOutput":
    """validate_batch for a Qwen provider: join -> loudness hook -> decay check, then the features of the HOOKED audio
    (resample 24k->16k -> log-mel) and the cosine, in the order _run_pipeline applies them (base_tts.py:911-926)."""
    out = qwen_pipeline_batch(rb, item_first_seg, p, qwen3_sr)
    dev = _dev_index(rb.data)
    h = Handle.get(dev)
    n = out.audio.n
    lens = out.records[:, 8:12].contiguous().view(torch.int32).reshape(-1)        # out_len (device)
    rb16 = resample_batch(out.audio, lengths=lens)
    mel, _ = logmel_batch(rb16, n_mels, pad_to_30s, lengths=rb16.lengths)
    if emb is not None and n:
        emb = emb.contiguous().float()
        ref = ref.contiguous().float().to(emb.device)
        cos_ptr = ctypes.c_void_p(out.records.data_ptr() + 28)                   # the cosine column of the 48-byte records
        _lib.check(h.lib.rho_b200_cosine(h.ptr, _ptr(emb), _ptr(ref), n, emb.shape[1], cos_ptr, 48, _stream(dev)),
                   "cosine")
    return ValidateOutput(out.audio, mel, out.records)


def cosine_batch(emb: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """dot(ref, e) / (|ref| |e|) per row of emb (base_tts.py:341-344)."""
    dev = _dev_index(emb)
    h = Handle.get(dev)
    emb = emb.contiguous().float()
    ref = ref.contiguous().float().to(emb.device)
    out = torch.empty(emb.shape[0], dtype=torch.float32, device=emb.device)
    _lib.check(h.lib.rho_b200_cosine(h.ptr, _ptr(emb), _ptr(ref), emb.shape[0], emb.shape[1], _ptr(out), 4,
                                     _stream(dev)), "cosine")
    return out


@dataclass
class ValidateOutput:
    audio: RaggedBatch
    mel: torch.Tensor           # [n, n_mels, T]: T = 3000, the unpadded frame count, or the compact row length
    records: torch.Tensor
    pad_value: Optional[torch.Tensor] = None    # [n] fp32 (30 s padding): the value of every frame past the signal

    def records_host(self) -> np.ndarray:
        return self.records.cpu().numpy().view(REC_DTYPE).reshape(-1)


class ValidatePlan:
    """Pre-allocated buffers for repeated rho_b200_validate calls on batches of one shape
    (the benchmark's steady state: allocation is not part of the hot path)."""

    def __init__(self, rb: RaggedBatch, item_first_seg: Sequence[int], p: RhoParams, n_mels: int = 80,
                 pad_to_30s: bool = True, fuse: bool = True, compact: bool = False, gather_first: bool = False):
        self.dev = _dev_index(rb.data)
        self.h = Handle.get(self.dev)
        self.p = p
        self.n_mels = int(n_mels)
        self.pad_frames = 3000 if pad_to_30s else 0
        first = np.asarray(item_first_seg, dtype=np.int32)
        self.n_items = len(first) - 1
        self.n_seg = rb.n
        self.max_seg_len = rb.max_len
        one_seg = self.n_items == self.n_seg and bool(np.all(np.diff(first) == 1))
        # compact (30 s padding only): rows hold just the frames that can see signal; every later frame of item i is
        # pad_value[i] (RHO_V_COMPACT_PAD) -- a third of the bytes for 10 s clips
        self.compact = bool(compact and pad_to_30s)
        self.flags = (_lib.V_ONE_SEGMENT_ITEMS if one_seg else 0) | (0 if fuse else _lib.V_NO_FUSION) | \
            (_lib.V_COMPACT_PAD if self.compact else 0) | (_lib.V_GATHER_FIRST if gather_first else 0)
        pause = int(p.sr * p.pause_sec) if p.pause_sec > 0 else 0
        seg_tot = np.concatenate([[0], np.cumsum(rb.h_lengths.astype(np.int64))])
        cap = (seg_tot[first[1:]] - seg_tot[first[:-1]]) + np.maximum(0, np.diff(first) - 2) * pause
        self.max_item_len = int(cap.max()) if self.n_items else 0
        self.out = RaggedBatch.empty_like_lengths(cap.astype(np.int32), rb.device)
        # the 16 kHz intermediate exists only on the kernel-per-stage path (fuse=False)
        self.scratch16 = torch.empty_like(self.out.data) if not fuse else None
        self.T = 3000 if pad_to_30s else max(((2 * self.max_item_len + 2) // 3) // 160, 1)
        if self.compact:
            self.T = int(self.h.lib.rho_b200_compact_frames(self.max_item_len, 3000))
        self.pad_value = torch.empty(self.n_items, dtype=torch.float32, device=rb.device) if pad_to_30s else None
        self.T_alloc = (self.T + 3) // 4 * 4      # rows 16-byte aligned: the normaliser's 128-bit path
        self.mel_buf = torch.empty((self.n_items, self.n_mels, self.T_alloc), dtype=torch.float32, device=rb.device)
        self.mel = self.mel_buf[:, :, :self.T]
        self.rec = torch.empty((self.n_items, 48), dtype=torch.uint8, device=rb.device)
        self.d_first = torch.from_numpy(first).to(rb.device)
        self.ws = _workspace(self.h, self.n_seg, self.n_items, self.max_seg_len, rb.device)

    def run(self, rb: RaggedBatch, emb: Optional[torch.Tensor], ref: Optional[torch.Tensor]) -> ValidateOutput:
        h = self.h
        _lib.check(h.lib.rho_b200_validate(
            h.ptr, _ptr(rb.data), _ptr(rb.offsets), _ptr(rb.lengths), self.n_seg, self.max_seg_len,
            _ptr(self.d_first), self.n_items, self.max_item_len, ctypes.byref(self.p),
            _ptr(self.out.data), _ptr(self.out.offsets), self.n_mels, self.pad_frames, _ptr(self.mel_buf), self.T_alloc,
            _ptr(self.pad_value), _ptr(emb), _ptr(ref), int(emb.shape[1]) if emb is not None else 0, _ptr(self.rec),
            _ptr(self.scratch16), self.flags, _ptr(self.ws), self.ws.numel(), _stream(self.dev)), "validate")
        return ValidateOutput(self.out, self.mel, self.rec, self.pad_value)


def validate_batch(rb: RaggedBatch, p: RhoParams, emb: Optional[torch.Tensor] = None,
                   ref: Optional[torch.Tensor] = None, n_mels: int = 80, pad_to_30s: bool = True,
                   item_first_seg: Optional[Sequence[int]] = None, fuse: bool = True,
                   compact: bool = False, gather_first: bool = False) -> ValidateOutput:
    """post-process/join -> resample 24k->16k -> log-mel -> cosine, all on the device.  gather_first (joined items
    only): k_gather writes the joined audio and the feature kernel reads it back, instead of one kernel doing both."""
    first = np.arange(rb.n + 1, dtype=np.int32) if item_first_seg is None else item_first_seg
    return ValidatePlan(rb, first, p, n_mels, pad_to_30s, fuse, compact, gather_first).run(rb, emb, ref)


def validate_host(x: torch.Tensor, p: RhoParams, emb: Optional[torch.Tensor], ref: Optional[torch.Tensor],
                  n_mels: int = 80, y: Optional[torch.Tensor] = None, mel: Optional[torch.Tensor] = None,
                  rec: Optional[torch.Tensor] = None, device: int = 0):
    """HOST buffers in, HOST buffers out (rho_b200_validate_host): x is a pinned CPU tensor (n, L);
    the library stages chunks through HBM with copy/compute overlap.  Returns (y, mel, rec) CPU tensors.  `mel` may be
    a CUDA tensor [n, n_mels, 3000] on `device`: the features are then written there and stay in HBM (only audio and
    records come back), for a consumer on the device such as the Whisper encoder."""
    assert not x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.is_contiguous()
    h = Handle.get(device)
    n, L = x.shape
    if y is None:
        y = torch.empty_like(x).pin_memory()
    if rec is None:
        rec = torch.empty((n, 48), dtype=torch.uint8).pin_memory()
    dim = int(emb.shape[1]) if emb is not None else 0
    _lib.check(h.lib.rho_b200_validate_host(h.ptr, _ptr(x), n, L, ctypes.byref(p), _ptr(y), int(n_mels), 3000,
                                            _ptr(mel), _ptr(emb), _ptr(ref), dim, _ptr(rec)), "validate_host")
    return y, mel, rec


@dataclass
class HostRaggedOutput:
    audio: torch.Tensor         # flat CPU fp32; item i at y_offsets[i], valid length records["out_len"][i]
    y_offsets: np.ndarray       # int64 [n_items]
    records: np.ndarray         # REC_DTYPE [n_items]
    mel: Optional[torch.Tensor]         # [n_items, n_mels, T] CPU (None: no features were asked for)
    pad_value: Optional[torch.Tensor]   # [n_items] CPU fp32: the value of every frame past the signal (30 s padding)


def host_item_layout(seg_lengths: np.ndarray, item_first_seg: np.ndarray, p: RhoParams, align: int = ALIGN):
    """Output offsets (aligned) and capacities for the items of a ragged host batch, as rho_b200_join wants them:
    sum of the item's segment lengths + max(0, n - 2) pauses."""
    first = np.asarray(item_first_seg, dtype=np.int64)
    pause = int(p.sr * p.pause_sec) if p.pause_sec > 0 else 0
    tot = np.concatenate([[0], np.cumsum(np.asarray(seg_lengths, dtype=np.int64))])
    cap = (tot[first[1:]] - tot[first[:-1]]) + np.maximum(0, np.diff(first) - 2) * pause
    padded = (cap + align - 1) // align * align
    off = np.concatenate([[0], np.cumsum(padded)])[:-1].astype(np.int64)
    return off, cap.astype(np.int64), int(padded.sum())


def validate_host_ragged(x: torch.Tensor, seg_offsets: np.ndarray, seg_lengths: np.ndarray,
                         item_first_seg: Sequence[int], p: RhoParams, emb: Optional[torch.Tensor] = None,
                         ref: Optional[torch.Tensor] = None, n_mels: int = 80, features: bool = True,
                         pad_to_30s: bool = True, compact: bool = False, y: Optional[torch.Tensor] = None,
                         mel: Optional[torch.Tensor] = None, device: int = 0) -> HostRaggedOutput:
    """rho_b200_validate_host_ragged: ragged segments / items in HOST memory (x flat CPU fp32, pinned for full speed)
    -> _smooth_segment_join + _validate_sound_decay per item (base_tts.py:435-536, 297-323) and, with features, the
    16 kHz resample + Whisper log-mel + cosine.  Everything comes back in host memory -- except the features when `mel`
    is given as a CUDA tensor on `device`: they are written there and stay in HBM."""
    assert not x.is_cuda and x.dim() == 1 and x.dtype == torch.float32 and x.is_contiguous()
    h = Handle.get(device)
    seg_off = np.ascontiguousarray(seg_offsets, dtype=np.int64)
    seg_len = np.ascontiguousarray(seg_lengths, dtype=np.int32)
    first = np.ascontiguousarray(item_first_seg, dtype=np.int32)
    n_items, n_seg = len(first) - 1, len(seg_len)
    y_off, cap, y_total = host_item_layout(seg_len, first, p)
    if y is None:
        y = torch.empty(max(y_total, 1), dtype=torch.float32).pin_memory()
    assert y.numel() >= y_total
    # pinned: a device -> pageable-host copy would block the enqueuing thread and serialise the chunk pipeline
    rec_t = torch.zeros((max(n_items, 1), 48), dtype=torch.uint8).pin_memory()
    rec = rec_t.numpy().view(REC_DTYPE).reshape(-1)[:n_items]
    pad_frames = 3000 if pad_to_30s else 0
    pad_value = None
    T = 0
    if features:
        max_cap = int(cap.max()) if n_items else 0
        T_need = int(h.lib.rho_b200_compact_frames(max_cap, pad_frames))
        T = 3000 if (pad_to_30s and not compact) else max((T_need + 3) // 4 * 4, 4)
        if mel is None:
            mel = torch.empty((n_items, n_mels, T), dtype=torch.float32).pin_memory()
        assert mel.shape == (n_items, n_mels, T) and mel.is_contiguous()
        if pad_to_30s:
            pad_value = torch.empty(n_items, dtype=torch.float32)
    else:
        mel = None
    dim = int(emb.shape[1]) if emb is not None else 0
    _lib.check(h.lib.rho_b200_validate_host_ragged(
        h.ptr, _ptr(x), ctypes.c_void_p(seg_off.ctypes.data), ctypes.c_void_p(seg_len.ctypes.data), n_seg,
        ctypes.c_void_p(first.ctypes.data), n_items, ctypes.byref(p), _ptr(y), ctypes.c_void_p(y_off.ctypes.data),
        int(n_mels), pad_frames, _ptr(mel), int(T), _ptr(pad_value), _ptr(emb), _ptr(ref), dim,
        _ptr(rec_t)), "validate_host_ragged")
    return HostRaggedOutput(y, y_off, rec, mel, pad_value)
