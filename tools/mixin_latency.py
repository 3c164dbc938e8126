"""Latency of the drop-in methods (B200AudioMixin) on ONE 10 s clip, the way a provider calls them: tensor in, tensor /
scalars out, synchronous.  CPU tensors are staged to the GPU and back by the mixin; CUDA tensors stay.
    python tools/mixin_latency.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import rho_tts_b200 as R  # noqa: E402
from rho_tts_b200 import synth  # noqa: E402


class T(R.B200QwenAudioMixin):
    device = "cuda"
    sample_rate = 24000
    silence_threshold_db = -50.0
    fade_duration_sec = 0.02
    crossfade_duration_sec = 0.05
    inter_sentence_pause_sec = 0.1
    trim_silence = True
    sound_decay_threshold = 0.3
    qwen3_sr = 24000


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


tts = T()
clips = synth.make_clips([240000, 120000, 168000], 3)
for where in ("cpu", "cuda"):
    x = clips[0].to(where)
    segs = [c.to(where) for c in clips]
    rows = [
        ("_trim_silence", lambda: tts._trim_silence(x)),
        ("_remove_dc_offset", lambda: tts._remove_dc_offset(x)),
        ("_apply_fades", lambda: tts._apply_fades(x.clone())),
        ("_smooth_segment_join (3 segments, 22 s)", lambda: tts._smooth_segment_join(segs)),
        ("_validate_sound_decay", lambda: tts._validate_sound_decay(x)),
        ("_post_process_audio (Qwen hook)", lambda: tts._post_process_audio(x.clone())),
        ("_apply_speed_pitch(speed=1.1)", lambda: tts._apply_speed_pitch(x, 1.1, 0.0)),
        ("_apply_speed_pitch(pitch=+2)", lambda: tts._apply_speed_pitch(x, 1.0, 2.0)),
    ]
    for name, fn in rows:
        print(f"{where:5s} {name:42s} {timeit(fn):8.3f} ms", flush=True)
