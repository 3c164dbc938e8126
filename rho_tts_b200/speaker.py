"""Speaker-similarity front end on the device (SURVEY.md 8f NEXT-3): the host mirror of what resemblyzer does between a
waveform and its LSTM, batched -- the reference's `_compute_speaker_similarity` (base_tts.py:326-347) runs
`preprocess_wav(wav, source_sr=self.sample_rate)` and `voice_encoder.embed_utterance(wav)` per clip on the CPU.

Names and argument meaning follow resemblyzer (audio.py / voice_encoder.py); every function here takes a ragged batch in
HBM and enqueues kernels of librho_b200 (csrc/spk.cu).  Not replaced: librosa's resampler (use `resample_batch` for
24 -> 16 kHz, a different anti-aliasing filter than soxr), webrtcvad's `trim_long_silences` (pass `vad=` to run one on the
host between normalisation and the spectrogram), and the LSTM (`encoder`: any callable [P, 160, 40] -> [P, D] on the
device, e.g. resemblyzer's own VoiceEncoder module moved to CUDA).
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import Handle
from .batch import _dev_index, _ptr, _stream, cosine_batch, resample_batch
from .ragged import RaggedBatch

SAMPLING_RATE = 16000
MEL_N_CHANNELS = 40
PARTIALS_N_FRAMES = 160
SAMPLES_PER_FRAME = 160
AUDIO_NORM_TARGET_DBFS = -30.0


def frame_step_of(rate: float = 1.3) -> int:
    """voice_encoder.py compute_partial_slices: int(np.round((sampling_rate / rate) / samples_per_frame))."""
    step = int(np.round((SAMPLING_RATE / rate) / SAMPLES_PER_FRAME))
    assert 0 < step, "The rate is too high"
    assert step <= PARTIALS_N_FRAMES, "The rate is too low, it should be %f at least" % (SAMPLING_RATE / (SAMPLES_PER_FRAME * PARTIALS_N_FRAMES))
    return step


def partial_count(n_samples: int, rate: float = 1.3, min_coverage: float = 0.75) -> Tuple[int, int]:
    """(number of partial utterances, sample the last one ends at) of a clip -- rho_b200_spk_slices, host only."""
    assert 0 < min_coverage <= 1
    padded = ctypes.c_int64(0)
    cnt = _lib.load().rho_b200_spk_slices(int(n_samples), frame_step_of(rate), float(min_coverage), ctypes.byref(padded))
    if cnt < 0:
        _lib.check(cnt, "spk_slices")
    return int(cnt), int(padded.value)


def partial_counts(lengths: np.ndarray, rate: float = 1.3, min_coverage: float = 0.75) -> Tuple[np.ndarray, np.ndarray]:
    """partial_count for an array of clip lengths (the same integer arithmetic as rho_b200_spk_slices, vectorised)."""
    assert 0 < min_coverage <= 1
    n = np.asarray(lengths, dtype=np.int64)
    step = frame_step_of(rate)
    n_frames = (n + SAMPLES_PER_FRAME) // SAMPLES_PER_FRAME                     # ceil((n + 1) / 160)
    steps = np.maximum(1, n_frames - PARTIALS_N_FRAMES + step + 1)
    count = (steps + step - 1) // step
    last = (count - 1) * step * SAMPLES_PER_FRAME
    coverage = (n - last).astype(np.float64) / float(PARTIALS_N_FRAMES * SAMPLES_PER_FRAME)
    count = count - ((coverage < min_coverage) & (count > 1))
    return count, ((count - 1) * step + PARTIALS_N_FRAMES) * SAMPLES_PER_FRAME


def compute_partial_slices(n_samples: int, rate: float = 1.3, min_coverage: float = 0.75):
    """Same return value as VoiceEncoder.compute_partial_slices: (wav_slices, mel_slices)."""
    cnt, _ = partial_count(n_samples, rate, min_coverage)
    step = frame_step_of(rate)
    mel_slices = [slice(step * j, step * j + PARTIALS_N_FRAMES) for j in range(cnt)]
    wav_slices = [slice(s.start * SAMPLES_PER_FRAME, s.stop * SAMPLES_PER_FRAME) for s in mel_slices]
    return wav_slices, mel_slices


def volume_gains(rb16: RaggedBatch, target_dBFS: float = AUDIO_NORM_TARGET_DBFS, increase_only: bool = False,
                 decrease_only: bool = False, out: Optional[RaggedBatch] = None) -> torch.Tensor:
    """The per-clip factor of audio.py normalize_volume, [n] fp32 on the device; `out` (may be rb16 itself) receives the
    scaled clips."""
    if increase_only and decrease_only:
        raise ValueError("Both increase only and decrease only are set")
    dev = _dev_index(rb16.data)
    h = Handle.get(dev)
    n = rb16.n
    gain = torch.empty((n,), dtype=torch.float32, device=rb16.device)
    if n == 0:
        return gain
    ws = torch.empty((n,), dtype=torch.float64, device=rb16.device)
    mode = 1 if increase_only else 2 if decrease_only else 0
    _lib.check(h.lib.rho_b200_normalize_volume(h.ptr, _ptr(rb16.data), _ptr(rb16.offsets), _ptr(rb16.lengths), 4, n,
                                               rb16.max_len, float(target_dBFS), mode,
                                               _ptr(out.data) if out is not None else None,
                                               _ptr(out.offsets) if out is not None else None, _ptr(gain), _ptr(ws),
                                               ws.numel() * 8, _stream(dev)), "normalize_volume")
    return gain


def normalize_volume(rb16: RaggedBatch, target_dBFS: float = AUDIO_NORM_TARGET_DBFS, increase_only: bool = False,
                     decrease_only: bool = False) -> RaggedBatch:
    """audio.py normalize_volume for every clip of the batch (a new batch; the input is not modified)."""
    out = RaggedBatch.empty_like_lengths(rb16.h_lengths, rb16.device)
    volume_gains(rb16, target_dBFS, increase_only, decrease_only, out=out)
    return out


def _mel_call(rb16: RaggedBatch, rate: float, min_coverage: float, pad: bool, gain: Optional[torch.Tensor],
              want_mel: bool, want_partials: bool):
    dev = _dev_index(rb16.data)
    h = Handle.get(dev)
    n = rb16.n
    step = frame_step_of(rate)
    lens = rb16.h_lengths.astype(np.int64)
    counts = np.zeros(n, dtype=np.int64)
    padded = lens.copy()
    if (pad or want_partials) and n:
        counts, ends = partial_counts(lens, rate, min_coverage)
        padded = np.maximum(lens, ends)
    frames = 1 + padded // SAMPLES_PER_FRAME
    frame_off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    part_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    mel = torch.empty((int(frame_off[-1]), MEL_N_CHANNELS), dtype=torch.float32, device=rb16.device) if want_mel else None
    partials = (torch.empty((int(part_off[-1]), PARTIALS_N_FRAMES, MEL_N_CHANNELS), dtype=torch.float32, device=rb16.device)
                if want_partials else None)
    d_foff = torch.from_numpy(frame_off).to(rb16.device) if want_mel else None
    d_poff = torch.from_numpy(part_off).to(rb16.device) if want_partials else None
    if n:
        _lib.check(h.lib.rho_b200_spk_mel(h.ptr, _ptr(rb16.data), _ptr(rb16.offsets), _ptr(rb16.lengths), 4, n,
                                          rb16.max_len, step, float(min_coverage), 1 if pad else 0, _ptr(gain), _ptr(mel),
                                          _ptr(d_foff), _ptr(partials), _ptr(d_poff), _stream(dev)), "spk_mel")
    return mel, frame_off, partials, part_off, d_poff


def wav_to_mel_spectrogram(rb16: RaggedBatch, gain: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, np.ndarray]:
    """audio.py wav_to_mel_spectrogram for every clip: ([sum T_i, 40] fp32 on the device, frame offsets [n + 1] on the
    host); clip i's [T_i, 40] spectrogram is rows frame_off[i] : frame_off[i + 1], T_i = 1 + len_i // 160."""
    mel, frame_off, _, _, _ = _mel_call(rb16, 1.3, 0.75, False, gain, True, False)
    return mel, frame_off


def partial_mels(rb16: RaggedBatch, rate: float = 1.3, min_coverage: float = 0.75,
                 gain: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, np.ndarray]:
    """The encoder input of embed_utterance for every clip: ([P, 160, 40] on the device, partial offsets [n + 1] on the
    host)."""
    _, _, partials, part_off, _ = _mel_call(rb16, rate, min_coverage, True, gain, False, True)
    return partials, part_off


def pool_partials(partial_embeds: torch.Tensor, part_off) -> torch.Tensor:
    """embed_utterance's tail: per clip, mean of its partial embeddings divided by its L2 norm -> [n, D]."""
    assert partial_embeds.dim() == 2 and partial_embeds.dtype == torch.float32 and partial_embeds.is_contiguous()
    dev = _dev_index(partial_embeds)
    h = Handle.get(dev)
    d_poff = part_off if isinstance(part_off, torch.Tensor) else torch.from_numpy(np.asarray(part_off, dtype=np.int32)).to(partial_embeds.device)
    n = int(d_poff.numel()) - 1
    out = torch.empty((n, partial_embeds.shape[1]), dtype=torch.float32, device=partial_embeds.device)
    if n > 0:
        _lib.check(h.lib.rho_b200_spk_pool(h.ptr, _ptr(partial_embeds), _ptr(d_poff), n, int(partial_embeds.shape[1]),
                                           _ptr(out), _stream(dev)), "spk_pool")
    return out


def preprocess_wav(rb: RaggedBatch, source_sr: Optional[int] = None,
                   vad: Optional[Callable[[List[np.ndarray]], List[np.ndarray]]] = None) -> RaggedBatch:
    """audio.py preprocess_wav for a batch: resample to 16 kHz (24 kHz input only -- the library's 3:2 polyphase filter,
    not librosa's soxr), raise quiet clips to -30 dBFS, and -- only if the caller brings one -- a host VAD over the clips
    (resemblyzer uses webrtcvad, which is not part of this library)."""
    if source_sr is not None and int(source_sr) != SAMPLING_RATE:
        if int(source_sr) != 24000:
            raise RuntimeError("rho_tts_b200.speaker.preprocess_wav: source_sr must be 16000 or 24000")
        rb = resample_batch(rb)
    out = normalize_volume(rb, AUDIO_NORM_TARGET_DBFS, increase_only=True)
    if vad is not None:
        clips = vad([out.clip(i).cpu().numpy() for i in range(out.n)])
        out = RaggedBatch.from_list([torch.from_numpy(np.asarray(c, dtype=np.float32)) for c in clips], rb.device)
    return out


def embed_utterances(rb16: RaggedBatch, encoder: Callable[[torch.Tensor], torch.Tensor], rate: float = 1.3,
                     min_coverage: float = 0.75, gain: Optional[torch.Tensor] = None) -> torch.Tensor:
    """VoiceEncoder.embed_utterance for every clip of a preprocessed 16 kHz batch -> [n, D] L2-normalised embeddings on
    the device.  `encoder` maps the [P, 160, 40] partial spectrograms to [P, D]."""
    partials, part_off = partial_mels(rb16, rate, min_coverage, gain)
    with torch.no_grad():
        pe = encoder(partials)
    return pool_partials(pe.to(torch.float32).contiguous(), part_off)


def speaker_similarity(rb: RaggedBatch, reference_embedding: torch.Tensor, encoder: Callable[[torch.Tensor], torch.Tensor],
                       sample_rate: int = 24000) -> torch.Tensor:
    """BaseTTS._compute_speaker_similarity (base_tts.py:326-347) for a batch of generated clips -> [n] cosines on the
    device.  Volume normalisation is folded into the spectrogram kernel (gain per clip), no scaled copy is written."""
    rb16 = rb if int(sample_rate) == SAMPLING_RATE else resample_batch(rb) if int(sample_rate) == 24000 else None
    if rb16 is None:
        raise RuntimeError("rho_tts_b200.speaker.speaker_similarity: sample_rate must be 16000 or 24000")
    gain = volume_gains(rb16, AUDIO_NORM_TARGET_DBFS, increase_only=True)
    emb = embed_utterances(rb16, encoder, gain=gain)
    return cosine_batch(emb, reference_embedding.to(device=emb.device, dtype=torch.float32).reshape(-1))
