"""Clip sharding across the GPUs of one box, and the one exchange of the path.

The reference has no multi-GPU code.  Clips and join items are independent,
so the path shards with NO data-path collective: each rank (one process per GPU) owns a contiguous
range of items, processes it locally, and the 48-byte per-item records are gathered once per batch.
Audio and features stay sharded.

The gather is not a collective call on the critical path: `RecordExchange` maps every rank's gathered
buffer into every process (CUDA IPC over NVLink 5 / NVSwitch peer memory) and the kernel that assembles
a record stores it into all of them (csrc/exchange.cu, records_dev.cuh).  Where peer mapping is not
available the records are all-gathered with NCCL on a side stream, overlapped with the next batch.
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


class RecordExchange:
    """The per-batch record gather of one process per GPU.

    mode "p2p"  : rho_b200_exchange_* -- every rho_b200_validate call on this device also stores its records into all
                  ranks' gathered buffers (NVLink peer stores issued by the record-assembling kernel) and publishes an
                  epoch flag; `after_step()` enqueues the flow control (wait for the previous epoch: the block of epoch e
                  is reused by epoch e + 2), `gathered()` waits for an epoch and returns its block.
    mode "nccl" : fallback when the peer mapping fails -- an `all_gather_into_tensor` of the local records on a side
                  stream, overlapped with the next batch (two buffers), joined in `after_step()` one step later.
    Every rank makes the same sequence of calls.  n_per_rank = records per rank and batch (equal on all ranks)."""

    def __init__(self, device: int, n_per_rank: int, group=None, force_nccl: bool = False):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = int(n_per_rank)
        self.device = int(device)
        self.h = _lib.Handle.get(self.device)
        self.mode = "nccl"
        self.why = "forced" if force_nccl else ""
        self.step = 0
        ok = 0
        handles = None
        if not force_nccl:
            mine = ctypes.create_string_buffer(64)
            rc = self.h.lib.rho_b200_exchange_create(self.h.ptr, self.world, self.rank, self.n, mine)
            if rc != 0:
                self.why = _lib.last_error()
            flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=torch.device("cuda", self.device))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(mine.raw), group=group)
                rc = self.h.lib.rho_b200_exchange_connect(self.h.ptr, b"".join(handles))
                if rc != 0:
                    self.why = _lib.last_error()
                flag.fill_(1 if rc == 0 else 0)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                ok = int(flag.item())
            if not ok:
                self.h.lib.rho_b200_exchange_destroy(self.h.ptr)     # all ranks fall back together
        if ok:
            self.mode = "p2p"
            dist.barrier(group=group)          # every rank is connected before anybody's kernels store to its peers
        else:
            dev = torch.device("cuda", self.device)
            self.side = torch.cuda.Stream(device=dev)
            self.bufs = [torch.empty((self.world * self.n, 48), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.stage = [torch.empty((self.n, 48), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.done_ev = [None, None]

    def after_step(self, records: torch.Tensor) -> None:
        """Call right after the rho_b200_validate launch of a batch, on the stream it was enqueued on."""
        import ctypes
        cur = torch.cuda.current_stream(self.device)
        if self.mode == "p2p":
            self.step = int(self.h.lib.rho_b200_exchange_epoch(self.h.ptr))   # = rho_b200_validate calls since connect
            if self.step >= 2:                  # flow control: the parity of this step's successor is free once epoch
                from . import _lib              # step - 1 has fully arrived everywhere (one step of slack: never stalls in lockstep)
                _lib.check(self.h.lib.rho_b200_exchange_wait(self.h.ptr, self.step - 1, ctypes.c_void_p(cur.cuda_stream)),
                           "exchange_wait")
            return
        self.step += 1
        b = self.step & 1
        if self.done_ev[b] is not None:         # the gather that used this pair of buffers two steps ago
            cur.wait_event(self.done_ev[b])
        self.stage[b][:records.shape[0]].copy_(records, non_blocking=True)   # the records buffer is rewritten next step
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.dist.all_gather_into_tensor(self.bufs[b], self.stage[b], group=self.group)
            done = torch.cuda.Event()
            done.record(self.side)
        self.done_ev[b] = done

    def gathered(self) -> torch.Tensor:
        """[world * n_per_rank, 48] uint8 view of the LAST step's gathered records, valid for work enqueued on the
        current stream after this call (and for the host after a synchronize)."""
        import ctypes
        from . import _lib
        cur = torch.cuda.current_stream(self.device)
        if self.mode == "p2p":
            _lib.check(self.h.lib.rho_b200_exchange_wait(self.h.ptr, self.step, ctypes.c_void_p(cur.cuda_stream)), "exchange_wait")
            out = torch.empty((self.world * self.n, 48), dtype=torch.uint8, device=torch.device("cuda", self.device))
            bad = ctypes.c_int(0)
            _lib.check(self.h.lib.rho_b200_exchange_read(self.h.ptr, self.step, ctypes.c_void_p(out.data_ptr()),
                                                         ctypes.byref(bad), ctypes.c_void_p(cur.cuda_stream)), "exchange_read")
            if bad.value:
                raise RuntimeError(f"record exchange: rank {bad.value - 1} did not publish epoch {self.step} within the time-out")
            return out
        b = self.step & 1
        cur.wait_event(self.done_ev[b])
        return self.bufs[b]

    def close(self) -> None:
        if self.mode == "p2p":
            torch.cuda.synchronize(self.device)
            self.dist.barrier(group=self.group)           # nobody unmaps while a peer may still store
            self.h.lib.rho_b200_exchange_destroy(self.h.ptr)
            self.mode = "closed"


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
