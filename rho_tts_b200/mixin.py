"""Drop-in host mirror of the reference's audio-processing methods.

`B200AudioMixin` overrides the BaseTTS methods on the hot path with the same names, arguments,
return types, aliasing and error behaviour (src/rho_tts/base_tts.py:297-536), and routes the
arithmetic to librho_b200 (hand-written sm_100a kernels).  Put it before a provider in the MRO:

    class QwenB200(B200AudioMixin, QwenTTS): ...
    TTSFactory.register_provider("qwen_b200", QwenB200)          # factory.py:110-122

Attributes (silence_threshold_db, fade_duration_sec, ...) are read at call time, exactly like the
reference (base_tts.py:366-367, 420, 455, 518).  Failures raise RuntimeError, never ValueError
(_run_pipeline treats ValueError as a configuration error, base_tts.py:786-787).  There is no
CPU fallback: without librho_b200.so or without a B200 the methods raise.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .batch import (REC_DTYPE, SEG_DTYPE, _ptr, _stream, cosine_batch, join_batch, params_from_tts,
                    trim_scan_batch)
from .ragged import RaggedBatch


class B200AudioMixin:
    #: CUDA device index used for the DSP (tensors living elsewhere are staged through it)
    b200_device: int = 0
    #: True: what resemblyzer does on the CPU around its LSTM (volume normalisation, 40-band spectrogram, partial
    #: utterances, pooling; rho_tts_b200.speaker) runs on the B200 as well.  Opt-in: the resampler is this library's 3:2
    #: polyphase filter instead of librosa's soxr and webrtcvad's silence trimming is skipped, so the similarity is close
    #: to, not identical with, the reference's.
    b200_speaker_front_end: bool = False

    # ------------------------------------------------------------------ helpers
    def _b200_dev(self) -> torch.device:
        if not torch.cuda.is_available():
            raise RuntimeError("rho_tts_b200: no CUDA device visible; the B200 audio path has no CPU fallback")
        return torch.device("cuda", self.b200_device)

    def _b200_mono(self, audio: torch.Tensor, what: str) -> torch.Tensor:
        if audio.dim() == 2 and audio.shape[0] != 1:
            raise RuntimeError(f"rho_tts_b200.{what}: multi-channel audio {tuple(audio.shape)} is not supported "
                               "(providers emit mono: qwen.py:265, chatterbox.py:167)")
        if audio.dim() > 2:
            raise RuntimeError(f"rho_tts_b200.{what}: expected (samples,) or (1, samples), got {tuple(audio.shape)}")
        return audio.reshape(-1)

    def _b200_stage(self, flat: torch.Tensor) -> torch.Tensor:
        """fp32 contiguous copy on the B200 (a new tensor unless it already is one)."""
        return flat.detach().to(device=self._b200_dev(), dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ base_tts.py:348-392
    def _trim_silence(self, audio: torch.Tensor, from_start: bool = True, from_end: bool = True) -> torch.Tensor:
        if not self.trim_silence or audio.numel() == 0:
            return audio
        flat = self._b200_mono(audio, "_trim_silence")
        dev = self._b200_dev()
        rb = RaggedBatch.from_list([flat], dev)
        flags = torch.tensor([(1 if from_start else 0) | (2 if from_end else 0)], dtype=torch.uint8, device=dev)
        info = trim_scan_batch(rb, params_from_tts(self), flags).cpu().numpy().view(SEG_DTYPE)[0]
        start, end = int(info["start"].item()), int(info["end"].item())
        a2 = audio.unsqueeze(0) if audio.dim() == 1 else audio
        if int(info["flags"].item()) & _lib.F_ALL_SILENT:
            return a2[:, start:end]                      # reference returns the 2-D view here (:380)
        return a2[:, start:end].squeeze(0)               # a view of the caller's tensor (:392)

    # ------------------------------------------------------------------ base_tts.py:394-399
    def _remove_dc_offset(self, audio: torch.Tensor) -> torch.Tensor:
        if audio.numel() == 0:
            return audio
        dev = self._b200_dev()
        work = audio.detach().to(device=dev, dtype=torch.float32).contiguous()
        if work.data_ptr() == audio.data_ptr():
            work = work.clone()                          # the reference returns a NEW tensor
        h = _lib.Handle.get(dev.index)
        ws = torch.empty(256, dtype=torch.uint8, device=dev)
        _lib.check(h.lib.rho_b200_remove_dc(h.ptr, _ptr(work), work.numel(), None, _ptr(ws), 256, _stream(dev.index)),
                   "remove_dc")
        return work.to(device=audio.device, dtype=audio.dtype)

    # ------------------------------------------------------------------ base_tts.py:401-433
    def _apply_fades(self, audio: torch.Tensor, fade_in: bool = True, fade_out: bool = True) -> torch.Tensor:
        if audio.numel() == 0:
            return audio
        original_shape = audio.shape
        flat = self._b200_mono(audio, "_apply_fades")
        fade_samples = int(self.sample_rate * self.fade_duration_sec)
        if flat.shape[-1] < fade_samples * 2 or fade_samples == 0 or not (fade_in or fade_out):
            return audio.view(original_shape)
        dev = self._b200_dev()
        in_place = flat.is_cuda and flat.device == dev and flat.dtype == torch.float32 and flat.is_contiguous() \
            and flat.data_ptr() == audio.data_ptr()
        work = flat if in_place else self._b200_stage(flat)
        if not in_place and work.data_ptr() == flat.data_ptr():
            work = work.clone()
        h = _lib.Handle.get(dev.index)
        p = params_from_tts(self)
        _lib.check(h.lib.rho_b200_apply_fades(h.ptr, _ptr(work), work.numel(), int(fade_in), int(fade_out),
                                              ctypes.byref(p), _stream(dev.index)), "apply_fades")
        if not in_place:
            with torch.no_grad():
                flat.copy_(work.to(device=flat.device, dtype=flat.dtype))   # the reference mutates its argument
        return audio.view(original_shape)

    # ------------------------------------------------------------------ base_tts.py:435-536
    def _smooth_segment_join(self, audio_segments: List[torch.Tensor]) -> Optional[torch.Tensor]:
        if len(audio_segments) == 0:
            return None
        dev = self._b200_dev()
        flats = [self._b200_mono(s, "_smooth_segment_join") for s in audio_segments]
        n_seg = len(flats)
        orig2d = audio_segments[0].dim() == 2
        target = torch.device(self.device) if n_seg > 1 else audio_segments[0].device
        if n_seg > 1 and orig2d and not self.trim_silence:
            # Trimming disabled hands the (1, L) tensors through unchanged, so every piece is 2-D
            # while the pause is 1-D: torch.cat throws and the reference falls back to the plain
            # concatenation of its inputs (:530-533).  Lengths are host-known here (no trim).
            cf = int(self.sample_rate * self.crossfade_duration_sec)
            pz = int(self.sample_rate * self.inter_sentence_pause_sec) if self.inter_sentence_pause_sec > 0 else 0
            L = [int(f.numel()) for f in flats]
            if pz > 0 and any(min(cf, L[i - 1], L[i]) > 10 for i in range(1, n_seg - 1)):
                y = torch.cat([self._b200_stage(f) for f in flats]).unsqueeze(0)
                return self._apply_fades(y, fade_in=True, fade_out=True).to(target)
        rb = RaggedBatch.from_list(flats, dev)
        out = join_batch(rb, [0, n_seg], params_from_tts(self), want_seg_info=False)
        rec = out.records_host()[0]
        y = out.audio.clip(0, int(rec["out_len"])).clone()
        flags = int(rec["flags"])
        two_d = bool(flags & _lib.F_TWO_D)
        if flags & _lib.F_FALLBACK:
            two_d = orig2d        # torch.cat of the caller's own tensors (:532)
        elif orig2d and (not self.trim_silence or (n_seg == 1 and flags & _lib.F_UNTOUCHED)):
            two_d = True          # untrimmed (1, L) inputs keep their rank; trimmed ones come back 1-D (:392)
        y = y.to(target)
        return y.unsqueeze(0) if two_d else y

    # ------------------------------------------------------------------ base_tts.py:297-323
    def _validate_sound_decay(self, audio: torch.Tensor) -> tuple:
        if audio.numel() == 0:
            return 1.0, True
        dev = self._b200_dev()
        work = self._b200_stage(audio.flatten())
        h = _lib.Handle.get(dev.index)
        p = params_from_tts(self)
        rec = torch.zeros(48, dtype=torch.uint8, device=dev)
        ws = torch.empty(256, dtype=torch.uint8, device=dev)
        _lib.check(h.lib.rho_b200_sound_decay(h.ptr, _ptr(work), work.numel(), ctypes.byref(p), _ptr(rec), _ptr(ws),
                                              256, _stream(dev.index)), "sound_decay")
        r = rec.cpu().numpy().view(REC_DTYPE)[0]
        return float(r["decay_ratio"]), bool(r["ok"])

    # ------------------------------------------------------------------ base_tts.py:325-346
    def _compute_speaker_similarity(self, wav_tensor: torch.Tensor) -> float:
        # The speaker encoder is third-party (resemblyzer) and out of scope; only the
        # cosine moves to the GPU.
        if self.b200_speaker_front_end:
            from .speaker import speaker_similarity
            enc = self.voice_encoder                         # resemblyzer's VoiceEncoder: forward([P, 160, 40]) -> [P, 256]
            dev = self._b200_dev()

            def encode(mels: torch.Tensor) -> torch.Tensor:
                return enc(mels.to(getattr(enc, "device", mels.device))).to(mels.device)
            rb = RaggedBatch.from_list([self._b200_stage(self._b200_mono(wav_tensor, "_compute_speaker_similarity"))], dev)
            ref = torch.as_tensor(np.asarray(self.reference_embedding, dtype=np.float32))
            return np.float32(speaker_similarity(rb, ref, encode, sample_rate=int(self.sample_rate)).item())
        from resemblyzer import preprocess_wav
        wav_np = wav_tensor.cpu().numpy().flatten()
        generated = self.voice_encoder.embed_utterance(preprocess_wav(wav_np, source_sr=self.sample_rate))
        return self._b200_cosine(self.reference_embedding, generated)

    # ------------------------------------------------------------------ base_tts.py:618-650
    def _apply_speed_pitch(self, audio: torch.Tensor, speed: float, pitch_semitones: float) -> torch.Tensor:
        """Speed = torchaudio.functional.resample(audio, int(sr * speed), sr) (any ratio) and pitch =
        torchaudio.functional.pitch_shift(audio, sr, pitch_semitones) (stft -> phase vocoder -> istft -> resample),
        both on the B200, in the reference's order."""
        if speed != 1.0:
            from .batch import resample_any_batch
            orig = int(self.sample_rate * speed)
            two_d = audio.dim() == 2
            flat = self._b200_mono(audio, "_apply_speed_pitch")
            if orig != self.sample_rate and flat.numel() > 0:
                dev = self._b200_dev()
                out = resample_any_batch(RaggedBatch.from_list([flat], dev), orig, self.sample_rate)
                y = out.clip(0).clone().to(device=audio.device, dtype=audio.dtype)
                audio = y.unsqueeze(0) if two_d and audio.shape[0] != 1 else y      # (1, L) comes back squeezed (:636-637)
        if pitch_semitones != 0.0:
            from .batch import pitch_shift_batch
            shape = audio.shape                                                     # :640-648: the shape is kept
            flat = self._b200_mono(audio, "_apply_speed_pitch")
            dev = self._b200_dev()
            out = pitch_shift_batch(RaggedBatch.from_list([flat], dev), int(self.sample_rate), float(pitch_semitones))
            audio = out.clip(0).clone().to(device=audio.device, dtype=audio.dtype).reshape(shape)
        return audio

    def _b200_cosine(self, reference_embedding, generated_embedding) -> np.float32:
        dev = self._b200_dev()
        ref = torch.as_tensor(np.asarray(reference_embedding, dtype=np.float32), device=dev)
        gen = torch.as_tensor(np.asarray(generated_embedding, dtype=np.float32), device=dev).reshape(1, -1)
        return np.float32(cosine_batch(gen, ref).item())


class B200QwenAudioMixin(B200AudioMixin):
    """B200AudioMixin plus the Qwen provider's loudness hook.  BaseTTS._post_process_audio is the identity
    (base_tts.py:601-614), so only providers that override it -- QwenTTS (providers/qwen.py:268-378) -- get
    this one:  class QwenB200(B200QwenAudioMixin, QwenTTS)."""

    def _post_process_audio(self, audio: torch.Tensor) -> torch.Tensor:
        from .batch import qwen_post_process_batch
        original_shape = audio.shape
        flat = audio.squeeze() if audio.dim() > 1 else audio            # qwen.py:284-286
        if flat.dim() != 1:
            raise RuntimeError(f"rho_tts_b200._post_process_audio: expected one non-trivial axis, got {tuple(original_shape)}")
        if flat.numel() == 0:
            return audio
        dev = self._b200_dev()
        sr = int(getattr(self, "qwen3_sr", None) or 24000)              # :294, read at call time
        rb = RaggedBatch.from_list([flat], dev)
        out = qwen_post_process_batch(rb, sr, in_place=True)
        return out.clip(0).to(device=audio.device, dtype=audio.dtype).reshape(original_shape)


def make_b200_provider(provider_class, name: Optional[str] = None):
    """class <Provider>B200(B200AudioMixin, <Provider>) -- the template of examples/custom_provider.py:22-51.
    Providers that define their own _post_process_audio named QwenTTS get the loudness hook as well."""
    base = B200QwenAudioMixin if provider_class.__name__ == "QwenTTS" else B200AudioMixin
    return type(name or f"{provider_class.__name__}B200", (base, provider_class), {})


def register_b200_providers(factory=None) -> list:
    """Register `<name>_b200` twins of every provider the reference's TTSFactory can import
    (factory.py:51-73, :110-122).  Needs the reference package `rho_tts` to be importable."""
    if factory is None:
        from rho_tts import TTSFactory as factory   # noqa: N813
    factory._register_default_providers()
    added = []
    for name, cls in list(factory._providers.items()):
        if name.endswith("_b200") or issubclass(cls, B200AudioMixin):
            continue
        factory.register_provider(f"{name}_b200", make_b200_provider(cls))
        added.append(f"{name}_b200")
    return added
