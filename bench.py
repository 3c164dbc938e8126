#!/usr/bin/env python
"""bench.py -- audio-seconds processed per second through the rho-tts validation front end on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W        (N > 1: launched by torchrun)
One JSON line on rank 0.  A "step" is one pass of the hot path (post-process -> resample 24k->16k ->
80-bin Whisper log-mel -> cosine) over one batch of synthetic clips per GPU.

Workload at every N: BASELINE.json configs[1] per GPU -- 1000 synthetic 10 s 24 kHz clips, full
post-process + 80-bin log-mel (30 s Whisper padding) + cosine vs a reference embedding.  Clips shard
across GPUs with no data-path collective; each step ends with the one NCCL all-gather of the 48-byte
per-clip records (weak scaling).

  value    : whole-job audio-seconds / second, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e      : the same through the host-buffer C-ABI call (rho_b200_validate_host): pinned host clips in,
             processed audio + records + features back to pinned host memory, copies inside the timed region
  roofline : dominant kernel, algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline : the numpy oracle (a port of the reference's torch/numpy path) on the host cores, bounded sample

--impl reference times that CPU oracle alone, on all host cores (the reference is pure Python: there is
no oracle/_ref binary).
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


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_CLIPS = None
_CPU_EMB = None


def _cpu_one(i: int):
    """The oracle pipeline for one clip: what the reference does per clip on the CPU
    (_smooth_segment_join([c]) -> _post_process_audio -> _validate_sound_decay, then the 16 kHz resample,
    Whisper log-mel and the embedding cosine of the validation front end)."""
    import oracle
    x = _CPU_CLIPS[i]
    emb, ref = _CPU_EMB
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
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = "1"


class CpuArm:
    """Bounded-sample CPU timing of the oracle on all host cores (fork pool, one BLAS thread each)."""

    def __init__(self, n_clips: int):
        global _CPU_CLIPS, _CPU_EMB
        import multiprocessing as mp
        from rho_tts_b200 import synth
        self.cores = os.cpu_count() or 1
        self.n_clips = n_clips
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
        return (f"{self.n_clips} x {CLIP_SECONDS:.0f} s clips of the C2 workload per step "
                f"(numpy oracle: post-process+decay, resample 24k->16k, 80-bin log-mel 30 s pad, cosine)")


def run_reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    n_clips = max(2 * (os.cpu_count() or 1), 32)
    arm = CpuArm(n_clips)
    for _ in range(max(args.warmup, 1)):
        arm.step()
    times = [arm.step() for _ in range(args.steps)]
    arm.close()
    total = sum(times)
    value = n_clips * CLIP_SECONDS * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.sample},
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
        "parallelism": f"clip-sharded x{n_gpus}, one NCCL all-gather of 48 B records per step" if n_gpus > 1
        else "single GPU",
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


def run_gpu_arm(args, rank: int, local_rank: int, world: int) -> None:
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU leg first, before CUDA is initialised in this process (fork pool)
        arm = CpuArm(max(2 * (os.cpu_count() or 1), 32))
        arm.step()
        reps = 0
        t_total = 0.0
        while reps < 3 or (t_total * arm.cores < 15.0 and reps < 40):
            t_total += arm.step()
            reps += 1
        arm.close()
        cpu_baseline = {"value": arm.n_clips * CLIP_SECONDS * reps / t_total, "unit": UNIT, "cores": arm.cores,
                        "kind": "port", "sample": f"{reps} passes over " + arm.sample}

    import numpy as np
    import torch
    import rho_tts_b200 as R
    from rho_tts_b200 import synth

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned host buffers of the e2e leg
    # (first touch) and the copies to / from them stay on the GPU's side of the host
    numa_cpus = R.dist.bind_to_gpu_numa(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL's version / debug banner goes to stderr: stdout carries the one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    n = args.clips
    x = synth.make_clip_block(n, CLIP_LEN, 0xB200 + rank, device=dev)           # generated on the owning GPU
    emb, ref = synth.make_embeddings(n, EMB_DIM, 4321 + rank, device=dev)
    rb = R.RaggedBatch.from_dense(x)
    p = R.make_params()
    plan = R.ValidatePlan(rb, np.arange(n + 1, dtype=np.int32), p, N_MELS, True)
    handle = R._lib.Handle.get(local_rank)
    gathered = torch.empty((world * n, 48), dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        out = plan.run(rb, emb, ref)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.records)      # the path's only collective
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = handle.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    torch.cuda.synchronize()
    t_host1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = handle.launch_count - l0

    audio_s_per_rank_step = n * CLIP_SECONDS
    value = world * audio_s_per_rank_step * args.steps / (ms_total * 1e-3)

    # ---- per-kernel device time (separate pass, events around every kernel on the launching stream)
    rec = out.records_host()
    prof_steps = max(1, min(args.steps, 20))
    handle.profile_begin()
    for _ in range(prof_steps):
        plan.run(rb, emb, ref)
    prof = handle.profile_end()
    sum_in = float(n) * CLIP_LEN
    sum_out = float(rec["out_len"].astype(np.int64).sum())
    len16 = (2 * rec["out_len"].astype(np.int64) + 2) // 3
    sum16 = float(len16.sum())
    t_real = np.minimum(PAD_FRAMES, np.maximum(2, (len16 + 200 + 159) // 160))
    sum_mel_real = float(t_real.sum()) * N_MELS
    fused_fills = bool(R._lib.load().rho_b200_build_flags() & 1)   # who writes the zero-padding frames' constant
    t_lo = np.minimum(PAD_FRAMES, (t_real + 3) // 4 * 4)
    fill_bytes = 4.0 * N_MELS * float((PAD_FRAMES - t_lo).sum())
    alg = {
        "k_scan": 4 * sum_in,
        "k_gather": 8 * sum_out,
        "k_resample3to2": 4 * sum_out + 4 * sum16,
        "k_logmel_frames": 4 * sum16 + 4 * sum_mel_real,
        # the normaliser reads every frame with signal and writes back only what the clamp changes (data-dependent, a few
        # per cent: not counted) -- plus the padding constant when the fused kernel does not write it
        "k_logmel_norm": 4 * sum_mel_real if fused_fills else 4.0 * N_MELS * PAD_FRAMES * n,
        # fused apply + resample + log-mel: x[start:end] in, y out, raw log-mel frames out (+ the padding constant)
        "k_fused_features": 8 * sum_out + 4 * sum_mel_real + (fill_bytes if fused_fills else 0.0),
    }
    peak, peak_src = load_peaks()
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        per = tot_ms / cnt
        k = {"ms_per_launch": per, "launches_per_step": cnt / prof_steps}
        if name in alg:
            k["algorithmic_bytes"] = alg[name]
            k["achieved_gbs"] = alg[name] / (per * 1e-3) / 1e9
            k["frac"] = k["achieved_gbs"] / peak
        kernels[name] = k
    dominant = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_per_launch"])
    dk = kernels[dominant]
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dk["frac"], "traffic": load_traffic(dominant), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dk["algorithmic_bytes"], "ms_per_launch": dk["ms_per_launch"],
                "share_of_step": dk["ms_per_launch"] * dk["launches_per_step"] /
                sum(v["ms_per_launch"] * v["launches_per_step"] for v in kernels.values())}
    pipeline_bytes = 4 * sum_in + 4 * sum_out + 4.0 * N_MELS * PAD_FRAMES * n + 48 * n + 4 * EMB_DIM * n
    pipeline = {"algorithmic_bytes_per_step": pipeline_bytes,
                "achieved_gbs": pipeline_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                "frac_of_hbm_peak": pipeline_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak}

    # ---- end to end through the HOST-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        xh = x.cpu().pin_memory()
        embh, refh = emb.cpu().pin_memory(), ref.cpu().pin_memory()
        yh = torch.empty_like(xh).pin_memory()
        melh = torch.empty((n, N_MELS, PAD_FRAMES), dtype=torch.float32).pin_memory()
        rech = torch.empty((n, 48), dtype=torch.uint8).pin_memory()
        e2e_steps = max(1, min(args.steps, 10))
        res = {}
        for label, mel_buf in (("all results to host", melh), ("features stay in HBM", None)):
            for _ in range(2):
                R.validate_host(xh, p, embh, refh, N_MELS, y=yh, mel=mel_buf, rec=rech, device=local_rank)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                R.validate_host(xh, p, embh, refh, N_MELS, y=yh, mel=mel_buf, rec=rech, device=local_rank)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            res[label] = world * audio_s_per_rank_step * e2e_steps / float(dt.item())
        h2d = xh.numel() * 4 + embh.numel() * 4 + refh.numel() * 4
        d2h = yh.numel() * 4 + rech.numel() + melh.numel() * 4
        e2e = {"value": res["all results to host"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "api": "rho_b200_validate_host (pinned host buffers, chunked copy/compute overlap)",
               "timer": "host perf_counter around the synchronous C call, max over ranks",
               "value_features_stay_in_hbm": res["features stay in HBM"],
               "d2h_bytes_per_step_features_stay_in_hbm": yh.numel() * 4 + rech.numel(),
               "host_numa_binding": (f"rank 0 bound to {len(numa_cpus)} CPUs local to its GPU (NVML)" if numa_cpus
                                     else "none")}
        # the host path must agree with the device path
        rh = rech.numpy().view(R.REC_DTYPE).reshape(-1)
        assert np.array_equal(rh["out_len"], rec["out_len"]) and np.array_equal(rh["ok"], rec["ok"])

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
        }
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
