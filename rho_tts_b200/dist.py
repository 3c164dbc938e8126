"""Clip sharding across the GPUs of one box, and the one collective of the path.

The reference has no multi-GPU code.  Clips and join items are independent,
so the path shards with NO data-path collective: each rank (one process per GPU) owns a contiguous
range of items, processes it locally, and the 48-byte per-item records are all-gathered once per
batch over NCCL (NVLink 5 / NVSwitch).  Audio and features stay sharded.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Equal-count contiguous ranges (fixed-length configs): the first n % world ranks get one more."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_samples(seg_lengths: Sequence[int], item_first_seg: Sequence[int], world: int) -> np.ndarray:
    """Balanced contiguous item ranges for ragged batches: cut points on the prefix sum of samples per
    item, items are never split.  Returns item boundaries b[0..world] (rank r owns items [b[r], b[r+1]))."""
    lens = np.asarray(seg_lengths, dtype=np.int64)
    first = np.asarray(item_first_seg, dtype=np.int64)
    n_items = len(first) - 1
    seg_pre = np.concatenate([[0], np.cumsum(lens)])
    item_pre = seg_pre[first]                      # samples before item i
    total = int(item_pre[-1])
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = n_items
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(item_pre, target, side="left"))
        # choose the neighbouring boundary closer to the target
        if i > 0 and (i > n_items or abs(item_pre[i - 1] - target) <= abs(item_pre[min(i, n_items)] - target)):
            i -= 1
        bounds[r] = min(max(i, bounds[r - 1]), n_items)
    return bounds


def gather_records(records: torch.Tensor, world: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gather of the local [n_local, 48] uint8 record block; every rank must contribute the same
    n_local (pad with zero rows otherwise).  Enqueued on the current stream right after the kernel that
    wrote the records; NCCL on CUDA tensors, gloo on CPU tensors (tests)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    out = torch.empty((world * records.shape[0],) + tuple(records.shape[1:]), dtype=records.dtype,
                      device=records.device)
    if records.is_cuda:
        dist.all_gather_into_tensor(out, records.contiguous(), group=group)
    else:
        parts = list(out.chunk(world, dim=0))
        dist.all_gather(parts, records.contiguous(), group=group)
    return out


def gather_records_ragged(records: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Same, for unequal per-rank counts: pads to max(counts), gathers, strips the padding."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    mx = int(max(counts))
    pad = torch.zeros((mx,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    pad[:records.shape[0]] = records
    allr = gather_records(pad, group=group).reshape(len(counts), mx, *records.shape[1:])
    return torch.cat([allr[r, :int(c)] for r, c in enumerate(counts)], dim=0)


def bind_to_gpu_numa(local_rank: int) -> Optional[list]:
    """Pin this process (one process per GPU) to the CPUs NVML reports as local to GPU `local_rank`, BEFORE it
    allocates pinned host buffers: first-touch then places them on the GPU's NUMA node, and the host <-> device copies
    of rho_b200_validate_host do not cross the socket interconnect.  With 8 ranks on a two-socket box this is the
    difference between every rank sharing one socket's memory controllers and each GPU using its own.
    Returns the CPU list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:                   # noqa: BLE001  (no NVML, no permission: keep the default placement)
        return None
