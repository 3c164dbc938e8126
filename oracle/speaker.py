"""Oracle (TEST INFRASTRUCTURE): resemblyzer's CPU front end and embedding pooling, numpy.

The reference computes speaker similarity with resemblyzer (base_tts.py:326-347 `_compute_speaker_similarity`:
`preprocess_wav(wav, source_sr)` -> `voice_encoder.embed_utterance(wav)` -> cosine; providers/qwen.py:199-216;
validation/classifier/trainer.py:41-47).  resemblyzer (pyproject.toml: `resemblyzer>=0.1.4`) is a third-party dependency
that is NOT under /root/reference and NOT in this image, and neither is librosa: PARITY UNPINNED against resemblyzer
itself.  This file restates the published algorithm of resemblyzer 0.1.4 (audio.py, hparams.py, voice_encoder.py):

    hparams          sampling_rate 16000, mel_window_length 25 ms, mel_window_step 10 ms, mel_n_channels 40,
                     partials_n_frames 160, audio_norm_target_dBFS -30
    normalize_volume rms = sqrt(mean((wav * 32767)^2)); dBFS = 20 log10(rms / 32767); change = target - dBFS;
                     unchanged if the mode forbids the sign of the change, else wav * 10^(change / 20)   (float32 wav)
    compute_partial_slices(n_samples, rate=1.3, min_coverage=0.75)
    wav_to_mel_spectrogram   librosa.feature.melspectrogram(y, sr=16000, n_fft=400, hop_length=160, n_mels=40).T
                     (librosa >= 0.10: periodic hann, center=True with ZERO padding, power 2, slaney mel / slaney norm)
    embed_utterance  zero-pad to the end of the last partial, mel, stack the partials' rows, encoder, mean, / L2 norm

and is pinned on the piece of that chain that IS in this image: transformers.audio_utils.spectrogram / mel_filter_bank (a
port of librosa's stft / filters.mel that transformers tests against librosa), tests/golden/make_golden_speaker.py.
Not restated: librosa.resample (soxr; the 16 kHz signal is an input here), trim_long_silences (webrtcvad, C code), the
LSTM (weights not in this image).
"""
from __future__ import annotations

import numpy as np

from .mfcc import mel_filterbank

F32 = np.float32
SR = 16000
N_FFT = 400            # int(16000 * 25 / 1000)
HOP = 160              # int(16000 * 10 / 1000)
N_MELS = 40
PARTIALS_N_FRAMES = 160
TARGET_DBFS = -30
INT16_MAX = (2 ** 15) - 1


def normalize_volume(wav: np.ndarray, target_dBFS: float = TARGET_DBFS, increase_only: bool = False,
                     decrease_only: bool = False) -> np.ndarray:
    """audio.py normalize_volume, numpy's own float32 arithmetic (wav is float32, python scalars are weak)."""
    if increase_only and decrease_only:
        raise ValueError("Both increase only and decrease only are set")
    wav = np.asarray(wav, dtype=F32)
    with np.errstate(all="ignore"):
        rms = np.sqrt(np.mean((wav * INT16_MAX) ** 2))
        wave_dBFS = 20 * np.log10(rms / INT16_MAX)
        dBFS_change = target_dBFS - wave_dBFS
        if dBFS_change < 0 and increase_only or dBFS_change > 0 and decrease_only:
            return wav
        return wav * (10 ** (dBFS_change / 20))


def volume_gain(wav: np.ndarray, target_dBFS: float = TARGET_DBFS, increase_only: bool = False,
                decrease_only: bool = False) -> float:
    """The factor normalize_volume multiplies by (1.0 where it returns the waveform itself)."""
    wav = np.asarray(wav, dtype=F32)
    if wav.size == 0:
        return 1.0
    with np.errstate(all="ignore"):
        rms = np.sqrt(np.mean((wav * INT16_MAX) ** 2))
        wave_dBFS = 20 * np.log10(rms / INT16_MAX)
        dBFS_change = target_dBFS - wave_dBFS
        if dBFS_change < 0 and increase_only or dBFS_change > 0 and decrease_only:
            return 1.0
        return float(F32(10 ** (dBFS_change / 20)))


def frame_step_of(rate: float = 1.3) -> int:
    return int(np.round((SR / rate) / HOP))


def compute_partial_slices(n_samples: int, rate: float = 1.3, min_coverage: float = 0.75):
    """voice_encoder.py compute_partial_slices -> (wav_slices, mel_slices)."""
    assert 0 < min_coverage <= 1
    samples_per_frame = HOP
    n_frames = int(np.ceil((n_samples + 1) / samples_per_frame))
    frame_step = frame_step_of(rate)
    assert 0 < frame_step, "The rate is too high"
    assert frame_step <= PARTIALS_N_FRAMES, "The rate is too low"
    wav_slices, mel_slices = [], []
    steps = max(1, n_frames - PARTIALS_N_FRAMES + frame_step + 1)
    for i in range(0, steps, frame_step):
        mel_range = np.array([i, i + PARTIALS_N_FRAMES])
        wav_range = mel_range * samples_per_frame
        mel_slices.append(slice(*mel_range))
        wav_slices.append(slice(*wav_range))
    last_wav_range = wav_slices[-1]
    coverage = (n_samples - last_wav_range.start) / (last_wav_range.stop - last_wav_range.start)
    if coverage < min_coverage and len(mel_slices) > 1:
        mel_slices = mel_slices[:-1]
        wav_slices = wav_slices[:-1]
    return wav_slices, mel_slices


def hann400() -> np.ndarray:
    k = np.arange(N_FFT, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / N_FFT)).astype(F32)


def wav_to_mel_spectrogram(wav: np.ndarray) -> np.ndarray:
    """audio.py wav_to_mel_spectrogram -> [T, 40] float32, T = 1 + len(wav) // 160."""
    wav = np.asarray(wav, dtype=F32).reshape(-1)
    yp = np.pad(wav, (N_FFT // 2, N_FFT // 2))
    T = 1 + wav.size // HOP
    fr = np.lib.stride_tricks.sliding_window_view(yp, N_FFT)[::HOP][:T]
    spec = np.fft.rfft((fr * hann400()[None, :]).astype(F32), axis=1)
    power = (spec.real.astype(F32) ** 2 + spec.imag.astype(F32) ** 2).astype(F32)          # [T, 201]
    return (power @ mel_filterbank(N_MELS, N_FFT, SR).T).astype(F32)


def partial_mels(wav: np.ndarray, rate: float = 1.3, min_coverage: float = 0.75) -> np.ndarray:
    """The encoder's input batch of embed_utterance: [n_partials, 160, 40]."""
    wav = np.asarray(wav, dtype=F32).reshape(-1)
    wav_slices, mel_slices = compute_partial_slices(len(wav), rate, min_coverage)
    max_wave_length = wav_slices[-1].stop
    if max_wave_length >= len(wav):
        wav = np.pad(wav, (0, max_wave_length - len(wav)), "constant")
    mel = wav_to_mel_spectrogram(wav)
    return np.array([mel[s] for s in mel_slices])


def pool_partials(partial_embeds: np.ndarray) -> np.ndarray:
    """embed_utterance's tail: mean over the partials, divided by the L2 norm."""
    raw = np.mean(np.asarray(partial_embeds, dtype=F32), axis=0)
    return raw / np.linalg.norm(raw, 2)


def embed_utterance(wav: np.ndarray, encoder, rate: float = 1.3, min_coverage: float = 0.75) -> np.ndarray:
    """voice_encoder.py embed_utterance with `encoder`: [P, 160, 40] float32 -> [P, D] (the LSTM stands outside)."""
    return pool_partials(encoder(partial_mels(wav, rate, min_coverage)))


def speaker_similarity(reference_embedding: np.ndarray, generated_embedding: np.ndarray) -> float:
    """base_tts.py:341-345."""
    dot = np.dot(reference_embedding, generated_embedding)
    return dot / (np.linalg.norm(reference_embedding) * np.linalg.norm(generated_embedding))
