"""Tensor-taking sibling of the reference's STT validation entry point (SURVEY.md 8f NEXT-2).

The reference validates every generated segment through a temporary WAV file
(base_tts.py:821-827: `audio.cpu()` -> `_save_wav` -> path; stt_validator.py:116-148: decode -> resample to 16 kHz ->
WhisperFeatureExtractor on the host -> features to the model) and returns `(is_valid, similarity, transcribed)`
(stt_validator.py:235-259).  `validate_audio_text_match_tensor` keeps that return contract and the text metric, but
takes the audio TENSOR: the 24 kHz -> 16 kHz resample and the Whisper log-mel run on the B200 (librho_b200), and the
[1, n_mels, 3000] feature tensor is handed to the Whisper model in place -- no D2H copy, no disk, no decode.

The STT model itself (weights, decoding loop, tokenizer) and the text normalisation / similarity metric are the
reference's: out of scope here, passed in or imported from the reference package.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from .batch import logmel_batch, resample_any_batch, resample_batch
from .ragged import RaggedBatch

WHISPER_SR = 16000


def whisper_features(audio: torch.Tensor, sample_rate: int, n_mels: int = 80, device: int = 0) -> torch.Tensor:
    """[1, n_mels, 3000] fp32 Whisper input features of one clip, computed on the B200 and left there:
    torchaudio.functional.resample(audio, sample_rate, 16000) (functional.py:1305-1432; the fixed 3:2 kernel for
    24 kHz) -> WhisperFeatureExtractor (feature_extraction_whisper.py:135-164, 296-303: 30 s pad / truncate)."""
    if not torch.cuda.is_available():
        raise RuntimeError("rho_tts_b200.validation: no CUDA device visible; the B200 path has no CPU fallback")
    if audio.dim() == 2 and audio.shape[0] == 1:
        audio = audio[0]
    if audio.dim() != 1:
        raise RuntimeError(f"rho_tts_b200.validation: expected mono audio (samples,) or (1, samples), got {tuple(audio.shape)}")
    if audio.numel() == 0:
        raise RuntimeError("rho_tts_b200.validation: empty audio")
    dev = torch.device("cuda", int(device))
    rb = RaggedBatch.from_list([audio.detach().to(device=dev, dtype=torch.float32)], dev)
    if int(sample_rate) == 24000:
        rb16 = resample_batch(rb)
    elif int(sample_rate) == WHISPER_SR:
        rb16 = rb
    else:
        rb16 = resample_any_batch(rb, int(sample_rate), WHISPER_SR)
    mel, _ = logmel_batch(rb16, n_mels=n_mels, pad_to_30s=True, lengths=rb16.lengths)
    return mel.contiguous()


def _reference_similarity() -> Callable[[str, str], float]:
    try:
        from rho_tts.validation.stt.stt_validator import calculate_text_similarity
        return calculate_text_similarity
    except Exception as e:      # noqa: BLE001
        raise RuntimeError("rho_tts_b200.validation: the reference's text metric "
                           "(rho_tts.validation.stt.stt_validator.calculate_text_similarity) is not importable; "
                           f"pass similarity_fn explicitly ({e!r})")


def transcribe_tensor(audio: torch.Tensor, sample_rate: int, model, tokenizer, *, device: int = 0,
                      generate_kwargs: Optional[dict] = None) -> Optional[str]:
    """Transcription of one clip from its tensor.  `model` is a transformers Whisper model (or anything with
    `.generate(input_features=...)` and, optionally, `.config.num_mel_bins` / `.dtype`) already on the B200;
    `tokenizer` has `batch_decode`.  Returns None when transcription fails, like stt_validator.py:116-148."""
    # the B200 part is outside the try: a missing library / GPU is a loud RuntimeError, not a skipped validation
    n_mels = int(getattr(getattr(model, "config", None), "num_mel_bins", 80) or 80)
    feats = whisper_features(audio, sample_rate, n_mels=n_mels, device=device)
    try:
        dtype = getattr(model, "dtype", torch.float32)
        with torch.no_grad():
            ids = model.generate(input_features=feats.to(dtype), **(generate_kwargs or {}))
        text = tokenizer.batch_decode(ids, skip_special_tokens=True)[0]
        return text.strip()
    except Exception:           # noqa: BLE001  (the reference logs and returns None: validation is then skipped)
        return None


def validate_audio_text_match_tensor(audio: torch.Tensor, sample_rate: int, expected_text: str, threshold: float = 0.85, *,
                                     model=None, tokenizer=None, device: int = 0,
                                     similarity_fn: Optional[Callable[[str, str], float]] = None,
                                     generate_kwargs: Optional[dict] = None) -> Tuple[bool, float, Optional[str]]:
    """validate_audio_text_match(audio_path, expected_text, threshold) (stt_validator.py:235-259) on a tensor:
    `(is_valid, similarity, transcribed)`; a failed transcription gives `(True, 0.0, None)` exactly like the reference
    (validation skipped, not failed)."""
    if model is None or tokenizer is None:
        raise RuntimeError("rho_tts_b200.validation: pass the STT model and its tokenizer (the Whisper weights are the "
                           "reference's concern, stt_validator.py:43-113)")
    transcribed = transcribe_tensor(audio, sample_rate, model, tokenizer, device=device, generate_kwargs=generate_kwargs)
    if transcribed is None:
        return True, 0.0, None
    sim = float((similarity_fn or _reference_similarity())(expected_text, transcribed))
    return sim >= threshold, sim, transcribed
