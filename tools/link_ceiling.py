"""Copy-only ceiling of the host link for the end-to-end (host-buffer) path.

Every rank moves the bytes one e2e step moves -- `h2d_bytes` from pinned host memory to the device and `d2h_bytes`
back, in chunks, on two streams at once -- and nothing else: no kernels.  The slower direction bounds what
rho_b200_validate_host can reach on this box; with N ranks (one process per GPU) all N links share the host, which is
what the 2 / 4 / 8-GPU e2e numbers have to be read against.

    python tools/link_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/link_ceiling.py

bench.py imports `measure` and puts the result next to `e2e` in its JSON line.
"""
from __future__ import annotations

import json
import os
import sys
import time


def measure(dev, h2d_bytes: int, d2h_bytes: int, reps: int = 5, chunk_bytes: int = 64 << 20, dist=None,
            h_src=None, h_dst=None) -> dict:
    """Concurrent pinned H2D + D2H of the given sizes, `reps` times; seconds per repetition = max over ranks.
    h_src / h_dst: existing pinned uint8 / float tensors to reuse (at least that many bytes), else allocated here."""
    import torch
    h2d_bytes, d2h_bytes = int(h2d_bytes), int(d2h_bytes)
    if h_src is None:
        h_src = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    if h_dst is None:
        h_dst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    h_src = h_src.view(torch.uint8).reshape(-1)[:h2d_bytes]
    h_dst = h_dst.view(torch.uint8).reshape(-1)[:d2h_bytes]
    d_in = torch.empty(min(h2d_bytes, 4 * chunk_bytes), dtype=torch.uint8, device=dev)
    d_out = torch.empty(min(d2h_bytes, 4 * chunk_bytes), dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def one(do_in: bool, do_out: bool):
        if do_in:
            with torch.cuda.stream(s_in):
                for k, o in enumerate(range(0, h2d_bytes, chunk_bytes)):
                    n = min(chunk_bytes, h2d_bytes - o)
                    so = (k % 4) * chunk_bytes if d_in.numel() > chunk_bytes else 0
                    d_in[so:so + n].copy_(h_src[o:o + n], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                for k, o in enumerate(range(0, d2h_bytes, chunk_bytes)):
                    n = min(chunk_bytes, d2h_bytes - o)
                    so = (k % 4) * chunk_bytes if d_out.numel() > chunk_bytes else 0
                    h_dst[o:o + n].copy_(d_out[so:so + n], non_blocking=True)

    def timed(do_in: bool, do_out: bool) -> float:
        one(do_in, do_out)
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            one(do_in, do_out)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    t_both = timed(True, True)
    t_in = timed(True, False)
    t_out = timed(False, True)
    return {
        "seconds_per_step_copy_only": t_both,
        "h2d_gbs_alone": h2d_bytes / t_in / 1e9, "d2h_gbs_alone": d2h_bytes / t_out / 1e9,
        "h2d_gbs_concurrent": h2d_bytes / t_both / 1e9, "d2h_gbs_concurrent": d2h_bytes / t_both / 1e9,
        "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes, "chunk_bytes": chunk_bytes, "reps": reps,
        "note": "per rank; seconds = max over ranks; pinned host memory, two streams, no kernels",
    }


def main() -> None:
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    # the bytes of one C2 step (1000 x 10 s clips): clips + embeddings in; processed audio, records and the feature
    # frames that can see signal (1004 of 3000 per row) out
    n, L = 1000, 240000
    h2d = 4 * n * L + 4 * 256 * (n + 1)
    d2h = 4 * n * L + 48 * n + 4 * 80 * 1004 * n + 4 * n
    out = measure(dev, h2d, d2h, dist=dist)
    out["n_gpus"] = world
    out["audio_s_per_s_copy_only"] = world * n * 10.0 / out["seconds_per_step_copy_only"]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
