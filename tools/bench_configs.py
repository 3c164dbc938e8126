"""The BASELINE.json configurations that are not the default bench line, at full size on one GPU:

  c2u : C2 with unpadded features (T = len16 // 160)
  c3  : 4000 ragged clips (1..30 s) grouped into items of 2..6 segments -> trim + crossfade joins (join_batch)
  c3v : the same items through the whole front end (join -> resample -> log-mel 80 -> cosine), unfused path
  c4  : one GPU's shard of C4: 8000 x 10 s clips, 128-bin log-mel (30 s pad), records
  qwen: the Qwen loudness hook (providers/qwen.py:268-378) on 1000 x 10 s post-processed clips
  pitch / mfcc / spk : the NEXT rows (pitch shift, MFCC statistics, resemblyzer's front end) on 1000 x 10 s clips

One JSON line per configuration: device-timed step (CUDA events, inputs resident, larger than L2), audio-s/s and
the per-kernel times with algorithmic bytes.  Results are copied to profiles/.
    python tools/bench_configs.py [c2u c3 c3v c4] [--steps K]
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rho_tts_b200 as R
from rho_tts_b200 import synth

SR = 24000
dev = torch.device("cuda", 0)
args = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()] or ["c2u", "c3", "c3v", "c4", "qwen"]
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
peak = 6543.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:   # noqa: BLE001
    pass


def ragged_c3(n=4000, seed=1234 + 3):
    """C3: lengths U[1, 30] s; every clip is the same generator as C2 (blocks of equal length share a launch)."""
    lens = synth.make_ragged_lengths(n, seed)
    rb = R.RaggedBatch.empty_like_lengths(lens, dev)
    order = np.argsort(lens)
    # generate in groups of similar length (one generator call per group, cut to each clip's length)
    for g0 in range(0, n, 50):
        idx = order[g0:g0 + 50]
        L = int(lens[idx].max())
        blk = synth.make_clip_block(len(idx), L, seed * 7919 + g0, device=dev)
        for j, i in enumerate(idx):
            rb.clip(int(i)).copy_(blk[j, :int(lens[i])])
    return rb, synth.make_item_partition(n, seed)


def timed(fn, steps):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    h = R._lib.Handle.get(0)
    h.profile_begin()
    for _ in range(min(steps, 5)):
        fn()
    prof = {k: v[0] / v[1] for k, v in h.profile_end().items()}
    return out, ms, prof


def line(name, workload, audio_s, ms, prof, alg):
    ks = {}
    for k, t in prof.items():
        ks[k] = {"ms": round(t, 4)}
        if k in alg:
            ks[k]["alg_GB"] = round(alg[k] / 1e9, 3)
            ks[k]["GBps"] = round(alg[k] / t / 1e6, 1)
            ks[k]["frac_hbm"] = round(alg[k] / t / 1e6 / peak, 3)
    print(json.dumps({"config": name, "workload": workload, "audio_s_per_step": audio_s, "ms_per_step": round(ms, 4),
                      "audio_s_per_s": round(audio_s / ms * 1e3, 1), "hbm_peak_gbs": peak, "kernels": ks}), flush=True)


p = R.make_params()
for cfg in args:
    torch.cuda.empty_cache()
    if cfg in ("c2u", "c4"):
        n, nm, pad = (1000, 80, False) if cfg == "c2u" else (8000, 128, True)
        x = synth.make_clip_block(n, 240000, 0xB200, device=dev) if n <= 1000 else \
            torch.cat([synth.make_clip_block(1000, 240000, 0xB200 + i, device=dev) for i in range(n // 1000)])
        emb, ref = synth.make_embeddings(n, device=dev)
        rb = R.RaggedBatch.from_dense(x)
        plan = R.ValidatePlan(rb, np.arange(n + 1, dtype=np.int32), p, nm, pad)
        out, ms, prof = timed(lambda: plan.run(rb, emb, ref), steps)
        rec = out.records_host()
        s_in, s_out = 4.0 * n * 240000, 4.0 * float(rec["out_len"].astype(np.int64).sum())
        len16 = (2 * rec["out_len"].astype(np.int64) + 2) // 3
        t_real = np.minimum(3000, np.maximum(2, (len16 + 359) // 160)) if pad else len16 // 160
        mel_real = 4.0 * nm * float(t_real.sum())
        mel_full = 4.0 * nm * (3000.0 * n if pad else float(t_real.sum()))
        fills = bool(R._lib.load().rho_b200_build_flags() & 1) and pad
        fill_b = 4.0 * nm * float((3000 - np.minimum(3000, (t_real + 3) // 4 * 4)).sum()) if pad else 0.0
        alg = {"k_scan": s_in, "k_fused_features": 2 * s_out + mel_real + (fill_b if fills else 0.0),
               "k_logmel_norm": mel_real if fills else mel_full}   # reads; writes only what the clamp changes
        line(cfg, f"{n} x 10 s clips, post-process + log-mel {nm} ({'30 s pad' if pad else 'unpadded'}) + cosine",
             n * 10.0, ms, prof, alg)
        del x, rb, plan, out
    elif cfg == "qwen":
        n = 1000
        x = synth.make_clip_block(n, 240000, 0xB200, device=dev)
        rb = R.RaggedBatch.from_dense(x)
        out, ms, prof = timed(lambda: R.qwen_post_process_batch(rb, 24000), steps)
        s_in = 4.0 * n * 240000
        alg = {"k_qwen_moments": s_in, "k_qwen_apply": 2 * s_in}
        line(cfg, f"{n} x 10 s clips, QwenTTS._post_process_audio (windowed decay correction, -23 dBFS, tanh)",
             n * 10.0, ms, prof, alg)
        del x, rb, out
    elif cfg.startswith("pitch"):
        n = 1000
        n_steps = float(cfg[5:]) if len(cfg) > 5 else 2.0
        x = synth.make_clip_block(n, 240000, 0xB200, device=dev)
        rb = R.RaggedBatch.from_dense(x)
        out, ms, prof = timed(lambda: R.pitch_shift_batch(rb, 24000, n_steps), steps)
        rate = 2.0 ** (-n_steps / 12.0)
        T = 1 + 240000 // 128
        J = math.ceil(T / rate)
        LS = round(240000 / rate)
        s_in = 4.0 * n * 240000
        spec_b, plane_b = 8.0 * 257 * T * n, 4.0 * 257 * J * n
        alg = {"k_pv_stft": s_in + spec_b, "k_pv_phase": spec_b + 2 * plane_b, "k_pv_cumsum": 2 * plane_b,
               "k_pv_istft": 2 * plane_b + 4.0 * LS * n, "k_resample_windowed": 4.0 * LS * n + s_in}
        line(cfg, f"{n} x 10 s clips, torchaudio.functional.pitch_shift n_steps={n_steps:+g} "
             "(stft 512/128 -> phase vocoder -> istft -> resample -> crop)", n * 10.0, ms, prof, alg)
        del x, rb, out
    elif cfg == "mfcc":
        n = 1000
        x = synth.make_clip_block(n, 160000, 0xB200, device=dev)          # 10 s at 16 kHz
        rb = R.RaggedBatch.from_dense(x)
        out, ms, prof = timed(lambda: R.mfcc_stats_batch(rb), steps)
        T = 1 + 160000 // 512
        alg = {"k_mfcc_frames": 4.0 * n * 160000 + 4.0 * 128 * T * n, "k_mfcc_stats": 4.0 * 128 * T * n}
        line(cfg, f"{n} x 10 s clips at 16 kHz, mean / std of librosa.feature.mfcc(n_mfcc=13) (stft 2048/512 -> mel 128 -> dB -> DCT)",
             n * 10.0, ms, prof, alg)
        del x, rb, out
    elif cfg == "spk":
        from rho_tts_b200 import speaker as SP
        n = 1000
        x = synth.make_clip_block(n, 160000, 0xB200, device=dev) * 0.05    # 10 s at 16 kHz, quiet: every clip is raised
        rb = R.RaggedBatch.from_dense(x)

        def run():
            gain = SP.volume_gains(rb, -30.0, increase_only=True)
            return SP.partial_mels(rb, gain=gain)
        out, ms, prof = timed(run, steps)
        P = int(out[1][-1])
        s_in = 4.0 * n * 160000
        alg = {"k_spk_sumsq": s_in, "k_spk_mel": s_in + 4.0 * 160 * 40 * P}
        line(cfg, f"{n} x 10 s clips at 16 kHz, resemblyzer front end: -30 dBFS gains + 40-band mel spectrogram -> "
             f"{P} partial utterances [160, 40]", n * 10.0, ms, prof, alg)
        del x, rb, out
    elif cfg in ("c3", "c3v", "c3vg"):
        rb, first = ragged_c3()
        audio_s = rb.total_samples / SR
        if cfg == "c3":
            out, ms, prof = timed(lambda: R.join_batch(rb, first, p, want_seg_info=False), steps)
            rec = out.records_host()
            s_in, s_out = 4.0 * rb.total_samples, 4.0 * float(rec["out_len"].astype(np.int64).sum())
            alg = {"k_scan": s_in, "k_gather": s_in + s_out}
            line(cfg, f"{rb.n} ragged clips 1..30 s in {len(first) - 1} items of 2..6 segments: trim + crossfade joins "
                 f"({rb.total_samples * 4 / 1e9:.2f} GB in)", audio_s, ms, prof, alg)
        else:
            emb, ref = synth.make_embeddings(len(first) - 1, device=dev)
            plan = R.ValidatePlan(rb, first, p, 80, False, gather_first=(cfg == "c3vg"))   # c3vg: k_gather, then features from y
            out, ms, prof = timed(lambda: plan.run(rb, emb, ref), steps)
            rec = out.records_host()
            s_in, s_out = 4.0 * rb.total_samples, 4.0 * float(rec["out_len"].astype(np.int64).sum())
            len16 = (2 * rec["out_len"].astype(np.int64) + 2) // 3
            mel_b = 4.0 * 80 * float((len16 // 160).sum())
            alg = {"k_scan": s_in, "k_gather": s_in + s_out, "k_resample3to2": s_out + s_out * 2 / 3,
                   "k_logmel_frames": s_out * 2 / 3 + mel_b, "k_logmel_norm": mel_b}
            line(cfg, f"{rb.n} ragged clips in {len(first) - 1} joined items -> resample -> log-mel 80 (unpadded) -> cosine",
                 audio_s, ms, prof, alg)
            del plan
        del rb, out
