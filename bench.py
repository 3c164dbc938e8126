#!/usr/bin/env python
"""bench.py -- audio-seconds processed per second through the rho-tts validation front end on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W        (N > 1: launched by torchrun)
One JSON line on rank 0.  A "step" is one pass of the hot path (post-process -> resample 24k->16k ->
80-bin Whisper log-mel -> cosine) over one batch of synthetic clips per GPU.

Headline workload at every N: BASELINE.json configs[1] per GPU -- 1000 synthetic 10 s 24 kHz clips, full
post-process + 80-bin log-mel (30 s Whisper padding) + cosine vs a reference embedding.  Clips shard across GPUs
with no data-path collective; the 48-byte per-clip records of every step are gathered on all ranks by peer stores
issued from the kernel that assembles them (rho_tts_b200.dist.RecordExchange; NCCL all-gather on a side stream where
peer mapping is unavailable) -- weak scaling.

  value    : whole-job audio-seconds / second, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e      : the same through the host-buffer C-ABI call (rho_b200_validate_host): pinned host clips in,
             processed audio + records + complete [80][3000] feature rows back in pinned host memory, copies inside
             the timed region; next to it the copy-only ceiling of the box's host link measured in the same run
  roofline : dominant kernel, algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
  configs  : the other BASELINE configurations in the same run -- C3 (ragged joins, 1 GPU), C4 (8000 clips per GPU,
             128-bin, record gather), C5 (256 k clips split over the ranks, device-resident waves, strong scaling)
  cpu_baseline : the reference's own CPU path on the host cores (baseline/_ref when it is installed, else the numpy
             oracle port), bounded sample

--impl reference times that CPU path alone, on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP_SECONDS = 10.0
SR = 24000
CLIP_LEN = int(CLIP_SECONDS * SR)
N_MELS = 80
PAD_FRAMES = 3000
EMB_DIM = 256
METRIC = "audio-seconds processed/sec (postproc+log-mel+cosine)"
UNIT = "audio-s/s"
REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_CLIPS = None
_CPU_EMB = None
_CPU_REF = None          # (Ref object with the reference's unbound methods, WhisperFeatureExtractor, torchaudio) or None


def _load_reference():
    """The UNMODIFIED reference (baseline/_ref: `pip install --target` of /root/reference) plus the two
    third-party functions its validation front end calls.  None when any of it is missing on this box."""
    if not os.path.isdir(os.path.join(REF_INSTALL, "rho_tts")):
        return None
    try:
        import logging
        if REF_INSTALL not in sys.path:
            sys.path.insert(0, REF_INSTALL)
        import torch
        import torchaudio
        from transformers import WhisperFeatureExtractor
        from rho_tts.base_tts import BaseTTS
        logging.getLogger("rho_tts").setLevel(logging.ERROR)
        logging.getLogger("rho_tts.base_tts").setLevel(logging.ERROR)

        class Ref:      # the reference's own test idiom (tests/test_audio_processing.py:7-29): unbound methods on a plain object
            device = "cpu"; silence_threshold_db = -50.0; crossfade_duration_sec = 0.05; trim_silence = True
            fade_duration_sec = 0.02; force_sentence_split = True; inter_sentence_pause_sec = 0.1
            sound_decay_threshold = 0.3; sample_rate = SR

        for n in ("_trim_silence", "_remove_dc_offset", "_apply_fades", "_smooth_segment_join", "_validate_sound_decay",
                  "_post_process_audio"):
            setattr(Ref, n, getattr(BaseTTS, n))
        return Ref(), WhisperFeatureExtractor(feature_size=N_MELS), torchaudio, torch
    except Exception:       # noqa: BLE001
        return None


def _cpu_one(i: int):
    """One clip through the CPU path, the way the reference handles it (base_tts.py:912-926, then the 16 kHz resample,
    Whisper features and the embedding cosine of the validation front end)."""
    x = _CPU_CLIPS[i]
    emb, ref = _CPU_EMB
    if _CPU_REF is not None:
        import numpy as np
        R, fe, torchaudio, torch = _CPU_REF
        y = R._smooth_segment_join([torch.from_numpy(x.copy())])
        y = R._post_process_audio(y)
        ratio, ok = R._validate_sound_decay(y)
        w16 = torchaudio.functional.resample(y.reshape(1, -1), SR, 16000)[0]
        mel = fe(w16.numpy(), sampling_rate=16000, return_tensors="np")["input_features"][0]
        cs = np.dot(ref, emb[i]) / (np.linalg.norm(ref) * np.linalg.norm(emb[i]))       # base_tts.py:341-344
        return float(mel[0, 0]) + float(cs) + float(ratio) + y.shape[-1]
    import oracle
    c = oracle.derive_constants()
    o = oracle.post_process_clip(x, c)
    w16 = oracle.resample(o["audio"])
    mel = oracle.log_mel(w16, N_MELS, True)
    cs = oracle.cosine_similarity(ref, emb[i])
    return float(mel[0, 0]) + float(cs) + o["out_len"]


def _cpu_worker_init():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:       # noqa: BLE001
        pass
    os.environ["OMP_NUM_THREADS"] = "1"
    if _CPU_REF is not None:
        _CPU_REF[3].set_num_threads(1)


class CpuArm:
    """Bounded-sample CPU timing on all host cores (fork pool, one thread per worker): the reference's own
    implementation when baseline/_ref is present (`kind: reference`), else the numpy oracle port (`kind: port`)."""

    def __init__(self, n_clips: int):
        global _CPU_CLIPS, _CPU_EMB, _CPU_REF
        import multiprocessing as mp
        import torch
        torch.set_num_threads(1)        # no OpenMP pool in the parent before the fork
        from rho_tts_b200 import synth
        self.cores = os.cpu_count() or 1
        self.n_clips = n_clips
        _CPU_REF = None if os.environ.get("RHO_BENCH_CPU_PORT") else _load_reference()
        self.kind = "reference" if _CPU_REF is not None else "port"
        x = synth.make_clip_block(n_clips, CLIP_LEN, 1234 + 1)           # CPU generator, config C2's seed
        emb, ref = synth.make_embeddings(n_clips)
        _CPU_CLIPS = [x[i].numpy() for i in range(n_clips)]
        _CPU_EMB = (emb.numpy(), ref.numpy())
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_worker_init)

    def step(self) -> float:
        t0 = time.perf_counter()
        self.pool.map(_cpu_one, range(self.n_clips), chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()

    @property
    def sample(self) -> str:
        what = ("rho_tts BaseTTS._smooth_segment_join -> _post_process_audio -> _validate_sound_decay, torchaudio resample "
                "24k->16k, transformers WhisperFeatureExtractor(80, 30 s pad), numpy cosine" if self.kind == "reference" else
                "numpy oracle port: post-process+decay, resample 24k->16k, 80-bin log-mel 30 s pad, cosine")
        return f"{self.n_clips} x {CLIP_SECONDS:.0f} s clips of the C2 workload per step, one worker per core ({what})"


def run_reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    n_clips = max(4 * (os.cpu_count() or 1), 64)
    arm = CpuArm(n_clips)
    for _ in range(max(args.warmup, 1)):
        arm.step()
    steps = max(1, min(args.steps, 20))
    times = [arm.step() for _ in range(steps)]
    arm.close()
    total = sum(times)
    value = n_clips * CLIP_SECONDS * steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, clips: int = 1000) -> dict:
    return {
        "workload": f"C2 per GPU: {clips} x 10 s synthetic 24 kHz clips, full post-process + 80-bin Whisper "
                    f"log-mel (30 s pad) + cosine vs reference embedding",
        "clips_per_gpu": clips, "clip_seconds": CLIP_SECONDS, "sample_rate": SR, "n_mels": N_MELS,
        "pad_frames": PAD_FRAMES, "emb_dim": EMB_DIM,
        "parallelism": f"clip-sharded x{n_gpus}, per-step gather of the 48 B records" if n_gpus > 1 else "single GPU",
        "l2_policy": "inputs larger than L2 (0.96 GB of clips per step vs 126 MB L2)",
    }


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the GPU is under load."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake"}

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.marks = []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:       # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:           # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0: float, t1: float) -> dict:
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "")}
        inside = [(m, r) for (t, m, r) in self.samples if t0 <= t <= t1]
        note = "sampled inside the timed region"
        if len(inside) < 3:             # very short region: fall back to every sample taken under load
            inside = [(m, r) for (_, m, r) in self.samples]
            note = "timed region shorter than the sampling period: all samples of this run (warm-up..e2e)"
        mhz = sorted(m for m, _ in inside)
        bits = 0
        for _, r in inside:
            bits |= r
        reasons = [name for bit, name in self.REASONS.items() if bits & bit]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "note": note}


# ------------------------------------------------------------------------------------------ GPU arm
def load_peaks() -> tuple:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:                   # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def load_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, if there is one."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(kernel)
    except Exception:                   # noqa: BLE001
        return None


def fused_alg_bytes(rec, n_mels: int, pad: bool, fused_fills: bool) -> dict:
    """Algorithmic HBM bytes per launch of the kernels of the one-segment-item path:
    4 B per input sample read by the scan; x[start:end] read + y written + the log-mel frames that see signal written by
    the fused kernel (+ the constant of the zero-padding frames when this build lets it write them); the normaliser
    reads the frames with signal (its writes are data dependent, a few per cent: not counted)."""
    import numpy as np
    out_len = rec["out_len"].astype(np.int64)
    len16 = (2 * out_len + 2) // 3
    n = len(out_len)
    if pad:
        t_real = np.minimum(PAD_FRAMES, np.maximum(2, (np.minimum(len16, PAD_FRAMES * 160) + 200 + 159) // 160))
        t_lo = np.minimum(PAD_FRAMES, (t_real + 3) // 4 * 4)
        fill = 4.0 * n_mels * float((PAD_FRAMES - t_lo).sum())
        full = 4.0 * n_mels * PAD_FRAMES * n
    else:
        t_real = np.where(len16 > 200, len16 // 160, 0)
        fill, full = 0.0, 4.0 * n_mels * float(t_real.sum())
    mel_real = 4.0 * n_mels * float(t_real.sum())
    s_out = 4.0 * float(out_len.sum())
    return {"k_fused_features": 2 * s_out + mel_real + (fill if fused_fills else 0.0),
            "k_logmel_norm": mel_real if (fused_fills or not pad) else full,
            "_mel_full": full, "_s_out": s_out}


def kernel_table(prof: dict, alg: dict, prof_steps: int, peak: float) -> dict:
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        per = tot_ms / cnt
        k = {"ms_per_launch": per, "launches_per_step": cnt / prof_steps}
        if name in alg:
            k["algorithmic_bytes"] = alg[name]
            k["achieved_gbs"] = alg[name] / (per * 1e-3) / 1e9
            k["frac"] = k["achieved_gbs"] / peak
        kernels[name] = k
    return kernels


def dominant_roofline(kernels: dict, alg: dict, peak: float, peak_src: str) -> dict:
    dominant = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"])
    dk = kernels[dominant]
    return {"kernel": dominant, "bound": "hbm", "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": dk["frac"], "traffic": load_traffic(dominant), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": dk["algorithmic_bytes"], "ms_per_launch": dk["ms_per_launch"],
            "share_of_step": dk["ms_per_launch"] * dk["launches_per_step"] /
            sum(v["ms_per_launch"] * v["launches_per_step"] for v in kernels.values())}


class DeviceTimer:
    """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""

    def __init__(self, dev, dist):
        self.dev, self.dist = dev, dist

    def run(self, step, steps: int, warmup: int, finish=None):
        import torch
        for _ in range(warmup):
            step()
        if finish:
            finish()
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            step()
        if finish:
            finish()                        # e.g. the wait for the last step's gathered records
        ev1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if self.dist is not None:
            self.dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1


def make_fixed_clips(n: int, seed: int, dev):
    """n x 10 s clips generated on the owning GPU in blocks of 1000 (the generator's temporaries stay small)."""
    import torch
    from rho_tts_b200 import synth
    if n <= 1000:
        return synth.make_clip_block(n, CLIP_LEN, seed, device=dev)
    x = torch.empty((n, CLIP_LEN), dtype=torch.float32, device=dev)
    for b0 in range(0, n, 1000):
        nb = min(1000, n - b0)
        x[b0:b0 + nb] = synth.make_clip_block(nb, CLIP_LEN, seed + 977 * (b0 // 1000), device=dev)
    return x


def make_ragged_c3(n: int, seed: int, dev, R):
    """C3: n clips with lengths U[1, 30] s, the C2 clip model; groups of similar length share one generator call."""
    import numpy as np
    from rho_tts_b200 import synth
    lens = synth.make_ragged_lengths(n, seed)
    rb = R.RaggedBatch.empty_like_lengths(lens, dev)
    order = np.argsort(lens)
    for g0 in range(0, n, 50):
        idx = order[g0:g0 + 50]
        blk = synth.make_clip_block(len(idx), int(lens[idx].max()), seed * 7919 + g0, device=dev)
        for j, i in enumerate(idx):
            rb.clip(int(i)).copy_(blk[j, :int(lens[i])])
    return rb, synth.make_item_partition(n, seed)


def run_gpu_arm(args, rank: int, local_rank: int, world: int) -> None:
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU leg first, before CUDA is initialised in this process (fork pool)
        arm = CpuArm(max(4 * (os.cpu_count() or 1), 64))
        arm.step()
        reps = 0
        t_total = 0.0
        while reps < 3 or (t_total * arm.cores < 15.0 and reps < 40):
            t_total += arm.step()
            reps += 1
        arm.close()
        cpu_baseline = {"value": arm.n_clips * CLIP_SECONDS * reps / t_total, "unit": UNIT, "cores": arm.cores,
                        "kind": arm.kind, "sample": f"{reps} passes over " + arm.sample}

    import numpy as np
    import torch
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned host buffers of the e2e leg
    # (first touch) and the copies to / from them stay on the GPU's side of the host
    numa_cpus = R.dist.bind_to_gpu_numa(local_rank) if world > 1 else None
    dist = None
    saved_stdout = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries the one JSON line only: NCCL prints its version banner there on some boxes, so the file
        # descriptor is pointed at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    timer = DeviceTimer(dev, dist)
    peak, peak_src = load_peaks()
    handle = R._lib.Handle.get(local_rank)
    fused_fills = bool(R._lib.load().rho_b200_build_flags() & 1)   # who writes the zero-padding frames' constant
    p = R.make_params()
    configs = {}
    want = set(args.configs.split(",")) if args.configs else set()

    # =================================================================== C2 (headline): 1000 x 10 s per GPU, 80 bins
    n = args.clips
    x = make_fixed_clips(n, 0xB200 + rank, dev)                                  # generated on the owning GPU
    emb, ref = synth.make_embeddings(n, EMB_DIM, 4321 + rank, device=dev)
    rb = R.RaggedBatch.from_dense(x)
    plan = R.ValidatePlan(rb, np.arange(n + 1, dtype=np.int32), p, N_MELS, True)
    ex = R.dist.RecordExchange(local_rank, n, force_nccl=bool(os.environ.get("RHO_BENCH_NCCL_GATHER"))) if world > 1 else None
    gather_mode = ex.mode if ex else "none"
    gather_why = ex.why if ex else ""

    def step():
        out = plan.run(rb, emb, ref)
        if ex:
            ex.after_step(out.records)              # p2p: flow control only (the stores are in the kernel); nccl: side stream
        return out

    gathered_last = {}

    def finish():
        if ex:
            gathered_last["rec"] = ex.gathered()    # every step's gather completes inside the timed region

    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = None

    def step_counted():
        return step()

    for _ in range(max(args.warmup, 3)):
        step()
    finish()
    torch.cuda.synchronize()
    l0 = handle.launch_count
    ms_total, t_host0, t_host1 = timer.run(step_counted, args.steps, 0, finish)
    launches = handle.launch_count - l0
    out = plan.run(rb, emb, ref) if ex is None else step()
    if ex:
        finish()
    torch.cuda.synchronize()
    rec = out.records_host()
    gather_ok = None
    if ex:
        g = gathered_last["rec"].cpu().numpy().view(R.REC_DTYPE).reshape(world, n)
        gather_ok = bool(g[rank].tobytes() == rec.tobytes())
        cnt = torch.tensor([int(g["out_len"].astype(np.int64).sum())], dtype=torch.int64, device=dev)
        lo, hi = cnt.clone(), cnt.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        gather_ok = gather_ok and int(lo.item()) == int(hi.item())      # every rank holds the same gathered block
        ex.close()                                  # the profiling / e2e passes below do not exchange records
        ex = None
    audio_s_per_rank_step = n * CLIP_SECONDS
    value = world * audio_s_per_rank_step * args.steps / (ms_total * 1e-3)

    # ---- per-kernel device time (separate pass, events around every kernel on the launching stream)
    prof_steps = max(1, min(args.steps, 20))
    handle.profile_begin()
    for _ in range(prof_steps):
        plan.run(rb, emb, ref)
    prof = handle.profile_end()
    alg = fused_alg_bytes(rec, N_MELS, True, fused_fills)
    alg["k_scan"] = 4.0 * n * CLIP_LEN
    kernels = kernel_table(prof, alg, prof_steps, peak)
    roofline = dominant_roofline(kernels, alg, peak, peak_src)
    pipeline_bytes = 4.0 * n * CLIP_LEN + alg["_s_out"] + alg["_mel_full"] + 48 * n + 4 * EMB_DIM * n
    pipeline = {"algorithmic_bytes_per_step": pipeline_bytes,
                "achieved_gbs": pipeline_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                "frac_of_hbm_peak": pipeline_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak}
    # compact feature rows (RHO_V_COMPACT_PAD): the frames that can see signal + one constant per clip
    plan_c = R.ValidatePlan(rb, np.arange(n + 1, dtype=np.int32), p, N_MELS, True, compact=True)
    ms_c, _, _ = timer.run(lambda: plan_c.run(rb, emb, ref), max(5, args.steps // 4), 3)
    compact = {"value": world * audio_s_per_rank_step * max(5, args.steps // 4) / (ms_c * 1e-3), "unit": UNIT,
               "frames_per_row": int(plan_c.T), "note": "rows of the frames that can see signal + pad_value per clip"}
    del plan_c

    # =================================================================== e2e: HOST buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        from tools.link_ceiling import measure as link_measure
        xh = x.cpu().pin_memory()
        embh, refh = emb.cpu().pin_memory(), ref.cpu().pin_memory()
        yh = torch.empty_like(xh).pin_memory()
        melh = torch.empty((n, N_MELS, PAD_FRAMES), dtype=torch.float32).pin_memory()
        rech = torch.empty((n, 48), dtype=torch.uint8).pin_memory()
        e2e_steps = max(1, min(args.steps, 10))

        def host_timed(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())

        dt_full = host_timed(lambda: R.validate_host(xh, p, embh, refh, N_MELS, y=yh, mel=melh, rec=rech, device=local_rank))
        # the host path must agree with the device path
        rh = rech.numpy().view(R.REC_DTYPE).reshape(-1)
        assert np.array_equal(rh["out_len"], rec["out_len"]) and np.array_equal(rh["ok"], rec["ok"])
        assert torch.equal(melh[::97], out.mel[::97].cpu()), "host feature rows differ from the device path"
        # compact rows on the host as well (what an embedder that feeds its own encoder input buffer would take)
        seg_off = np.arange(n, dtype=np.int64) * CLIP_LEN
        seg_len = np.full(n, CLIP_LEN, dtype=np.int32)
        first = np.arange(n + 1, dtype=np.int32)
        T_c = int(R._lib.load().rho_b200_compact_frames(CLIP_LEN, PAD_FRAMES))
        melc = torch.empty((n, N_MELS, T_c), dtype=torch.float32).pin_memory()
        dt_comp = host_timed(lambda: R.validate_host_ragged(xh.reshape(-1), seg_off, seg_len, first, p, embh, refh, N_MELS,
                                                            compact=True, y=yh.reshape(-1), mel=melc, device=local_rank))
        # features left in HBM (mel is device memory): what the call costs when the Whisper encoder runs on the same GPU
        # (SURVEY 8f NEXT-2); audio and records still come back to the host
        mel_dev = torch.empty((n, N_MELS, PAD_FRAMES), dtype=torch.float32, device=dev)
        dt_hbm = host_timed(lambda: R.validate_host(xh, p, embh, refh, N_MELS, y=yh, mel=mel_dev, rec=rech, device=local_rank))
        assert torch.equal(mel_dev[::97], out.mel[::97]), "device-resident feature rows differ from the device path"
        del mel_dev
        h2d = xh.numel() * 4 + embh.numel() * 4 + refh.numel() * 4
        d2h_link = yh.numel() * 4 + rech.numel() + melc.numel() * 4 + 4 * n     # what crosses the link
        link = link_measure(dev, h2d, d2h_link, reps=3, dist=dist)
        ceiling = world * audio_s_per_rank_step / link["seconds_per_step_copy_only"]
        e2e_value = world * audio_s_per_rank_step * e2e_steps / dt_full
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_link,
               "steps": e2e_steps,
               "api": "rho_b200_validate_host: pinned host clips in; processed audio, records and complete [80][3000] "
                      "feature rows back in pinned host memory (only the frames that can see signal cross the link, the "
                      "constant tail of every row is written by host threads from the per-clip value)",
               "result_bytes_in_host_memory_per_step": yh.numel() * 4 + rech.numel() + melh.numel() * 4,
               "timer": "host perf_counter around the synchronous C call, max over ranks",
               "value_compact_rows": world * audio_s_per_rank_step * e2e_steps / dt_comp,
               "host_fill_threads": int(R._lib.load().rho_b200_host_fill_threads(R._lib.Handle.get(local_rank).ptr)),
               "value_features_in_hbm": world * audio_s_per_rank_step * e2e_steps / dt_hbm,
               "value_features_in_hbm_note": "same call with `mel` a device buffer: complete feature rows stay in HBM for a "
                                             "consumer on the GPU, audio and records come back (not the headline: the "
                                             "reference arm leaves its features in host memory)",
               "link_ceiling_gbs": {"h2d": link["h2d_gbs_concurrent"], "d2h": link["d2h_gbs_concurrent"],
                                    "h2d_alone": link["h2d_gbs_alone"], "d2h_alone": link["d2h_gbs_alone"]},
               "link_ceiling_value": ceiling, "frac_of_link_ceiling": e2e_value / ceiling,
               "link_ceiling_note": "copy-only run of the same bytes per step on this box in this run (tools/link_ceiling.py): "
                                    "per rank, all ranks at once, max over ranks",
               "host_numa_binding": (f"rank 0 bound to {len(numa_cpus)} CPUs local to its GPU (NVML)" if numa_cpus
                                     else "none")}
        del xh, yh, melh, melc
    del plan, rb, x, out
    torch.cuda.empty_cache()

    # =================================================================== C4 / C5: 8000 clips per GPU
    if "c4" in want or "c5" in want:
        n4 = 8000
        x4 = make_fixed_clips(n4, 0xC400 + 31 * rank, dev)
        emb4, ref4 = synth.make_embeddings(n4, EMB_DIM, 99 + rank, device=dev)
        rb4 = R.RaggedBatch.from_dense(x4)
        first4 = np.arange(n4 + 1, dtype=np.int32)
        if "c4" in want:
            plan4 = R.ValidatePlan(rb4, first4, p, 128, True)
            ex4 = R.dist.RecordExchange(local_rank, n4) if world > 1 else None

            def step4():
                o = plan4.run(rb4, emb4, ref4)
                if ex4:
                    ex4.after_step(o.records)
                return o

            g4 = {}

            def fin4():
                if ex4:
                    g4["rec"] = ex4.gathered()
            k4 = 5
            ms4, _, _ = timer.run(step4, k4, 2, fin4)
            o4 = step4(); fin4(); torch.cuda.synchronize()
            rec4 = o4.records_host()
            if ex4:
                gg = g4["rec"].cpu().numpy().view(R.REC_DTYPE).reshape(world, n4)
                assert gg[rank].tobytes() == rec4.tobytes(), "C4: gathered records differ from the local ones"
                ex4.close()
            handle.profile_begin()
            plan4.run(rb4, emb4, ref4)
            prof4 = handle.profile_end()
            alg4 = fused_alg_bytes(rec4, 128, True, fused_fills)
            alg4["k_scan"] = 4.0 * n4 * CLIP_LEN
            kern4 = kernel_table(prof4, alg4, 1, peak)
            configs["C4"] = {
                "workload": f"{world} x {n4} x 10 s clips ({world * n4} clips), post-process + 128-bin log-mel (30 s pad) + "
                            f"cosine, per-step gather of all records on every rank ({ex4.mode if ex4 else 'single GPU'})"
                            + (" -- BASELINE configs[3] at N = 8" if world == 8 else ""),
                "value": world * n4 * CLIP_SECONDS * k4 / (ms4 * 1e-3), "unit": UNIT, "ms_per_step": ms4 / k4, "steps": k4,
                "scaling": "weak", "clips_per_gpu": n4, "roofline": dominant_roofline(kern4, alg4, peak, peak_src),
                "kernels_ms": {k: round(v["ms_per_launch"] * v["launches_per_step"], 4) for k, v in kern4.items()},
                "accept_rate": float(rec4["ok"].mean())}
            del plan4, o4
            torch.cuda.empty_cache()
        if "c5" in want:
            total5 = 256000
            lo5, hi5 = R.dist.shard_range(total5, rank, world)
            mine = hi5 - lo5
            waves = [(w0, min(n4, mine - w0)) for w0 in range(0, mine, n4)]
            plan5 = R.ValidatePlan(rb4, first4, p, N_MELS, True)
            plans_tail = {}
            ex5 = R.dist.RecordExchange(local_rank, n4) if world > 1 else None

            def run_waves():
                for _, cnt in waves:
                    if cnt == n4:
                        o = plan5.run(rb4, emb4, ref4)
                    else:                       # the ragged last wave of this rank
                        if cnt not in plans_tail:
                            sub = R.RaggedBatch.from_dense(x4[:cnt])
                            plans_tail[cnt] = (sub, R.ValidatePlan(sub, np.arange(cnt + 1, dtype=np.int32), p, N_MELS, True))
                        sub, pl = plans_tail[cnt]
                        o = pl.run(sub, emb4[:cnt], ref4)
                    if ex5:
                        ex5.after_step(o.records)

            def fin5():
                if ex5:
                    ex5.gathered()
            ms5, _, _ = timer.run(run_waves, 1, 1, fin5)
            if ex5:
                ex5.close()
            configs["C5"] = {
                "workload": f"{total5} x 10 s clips (2.56 M audio-s, ~711 h) split over {world} rank(s) with dist.shard_range, "
                            f"processed in device-resident waves of {n4} clips (each wave re-reads this rank's resident 7.7 GB "
                            f"wave buffer, 60x L2), post-process + 80-bin log-mel (30 s pad) + cosine, records gathered per wave",
                "value": total5 * CLIP_SECONDS / (ms5 * 1e-3), "unit": UNIT, "seconds": ms5 * 1e-3, "scaling": "strong",
                "clips_per_rank": mine, "waves_per_rank": len(waves)}
            del plan5, plans_tail
        del rb4, x4
        torch.cuda.empty_cache()

    # =================================================================== C3: ragged joins, one GPU
    if "c3" in want and world == 1:
        rb3, first3 = make_ragged_c3(4000, 1234 + 3, dev, R)
        n_items3 = len(first3) - 1
        audio3 = rb3.total_samples / SR
        emb3, ref3 = synth.make_embeddings(n_items3, EMB_DIM, 7, device=dev)
        k3 = 5
        msj, _, _ = timer.run(lambda: R.join_batch(rb3, first3, p, want_seg_info=False), k3, 2)
        oj = R.join_batch(rb3, first3, p, want_seg_info=False)
        recj = oj.records_host()
        handle.profile_begin()
        R.join_batch(rb3, first3, p, want_seg_info=False)
        profj = handle.profile_end()
        s_in3, s_out3 = 4.0 * rb3.total_samples, 4.0 * float(recj["out_len"].astype(np.int64).sum())
        kernj = kernel_table(profj, {"k_scan": s_in3, "k_gather": s_in3 + s_out3}, 1, peak)
        del oj
        plan3 = R.ValidatePlan(rb3, first3, p, N_MELS, False)
        msv, _, _ = timer.run(lambda: plan3.run(rb3, emb3, ref3), k3, 2)
        handle.profile_begin()
        plan3.run(rb3, emb3, ref3)
        profv = handle.profile_end()
        len16 = (2 * recj["out_len"].astype(np.int64) + 2) // 3
        mel3 = 4.0 * N_MELS * float((len16 // 160).sum())
        # one kernel joins, writes y and computes the features: x read once, y and the features written once
        kernv = kernel_table(profv, {"k_scan": s_in3, "k_fused_features": s_in3 + s_out3 + mel3, "k_logmel_norm": mel3}, 1, peak)
        c3 = {"workload": f"4000 ragged clips of 1..30 s ({rb3.total_samples * 4 / 1e9:.2f} GB, {audio3:.0f} audio-s) in "
                          f"{n_items3} items of 2..6 segments",
              "join": {"what": "trim + DC + crossfade joins + pauses + fades + decay check into batch outputs (BASELINE configs[2])",
                       "value": audio3 * k3 / (msj * 1e-3), "unit": UNIT, "ms_per_step": msj / k3,
                       "roofline": dominant_roofline(kernj, {"k_scan": 1, "k_gather": 1}, peak, peak_src)},
              "join_and_features": {"what": "the same items through the whole front end in one pass over the segments: join "
                                            "(y written) + resample + 80-bin log-mel (unpadded) in ONE kernel -> cosine",
                                    "value": audio3 * k3 / (msv * 1e-3), "unit": UNIT, "ms_per_step": msv / k3,
                                    "kernels_ms": {k: round(v["ms_per_launch"] * v["launches_per_step"], 4) for k, v in kernv.items()},
                                    "frac": {k: round(v["frac"], 3) for k, v in kernv.items() if "frac" in v}}}
        del plan3
        torch.cuda.empty_cache()
        if not args.no_e2e:
            # C3 end to end: ragged segments in pinned host memory -> joined items + records back in host memory
            lens3 = rb3.h_lengths.astype(np.int32)
            off3 = rb3.h_offsets.astype(np.int64)
            xh3 = rb3.data.cpu().pin_memory()
            y_off3, cap3, y_total3 = R.host_item_layout(lens3, first3, p)
            yh3 = torch.empty(y_total3, dtype=torch.float32).pin_memory()
            R.validate_host_ragged(xh3, off3, lens3, first3, p, features=False, y=yh3, device=local_rank)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps3 = 2
            for _ in range(reps3):
                h3 = R.validate_host_ragged(xh3, off3, lens3, first3, p, features=False, y=yh3, device=local_rank)
            dt3 = (time.perf_counter() - t0) / reps3
            assert np.array_equal(h3.records["out_len"], recj["out_len"]) and np.array_equal(h3.records["ok"], recj["ok"])
            c3["join"]["e2e"] = {"value": audio3 / dt3, "unit": UNIT, "h2d_bytes_per_step": int(xh3.numel() * 4),
                                 "d2h_bytes_per_step": int(y_total3 * 4 + 48 * n_items3),
                                 "api": "rho_b200_validate_host_ragged (mel = NULL)"}
            del xh3, yh3
        configs["C3"] = c3
        del rb3
        torch.cuda.empty_cache()

    sampler.stop()
    sampler.join(timeout=1.0)
    clocks = sampler.summary(t_host0, t_host1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, n), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "pipeline_roofline": pipeline, "kernels": kernels,
            "cpu_baseline": cpu_baseline, "impl": "rho_tts_b200",
            "accept_rate": float(rec["ok"].mean()),
            "record_gather": {"mode": gather_mode, "verified": gather_ok, "fallback_reason": gather_why or None},
            "value_compact_rows": compact, "configs": configs,
        }
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=1000, help="clips per GPU per step (C2: 1000)")
    ap.add_argument("--configs", default="c3,c4,c5", help="other BASELINE configurations to run in the same job ('' = none)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
