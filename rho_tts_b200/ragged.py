"""Ragged clip batches in HBM.

Layout (DESIGN.md "Data layout"): one fp32 buffer holding every clip back to back, each clip
start padded to a multiple of ALIGN samples (128 B) so that 128-bit loads are legal and rows do
not share cache lines; `offsets` (int64) and `lengths` (int32) live both on the device (kernel
arguments) and on the host (grid sizing, slicing) so no call needs a device->host sync.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

ALIGN = 32  # samples (128 bytes)


def _round_up(v: int, a: int = ALIGN) -> int:
    return (v + a - 1) // a * a


@dataclass
class RaggedBatch:
    data: torch.Tensor        # float32 [total], device
    offsets: torch.Tensor     # int64 [n], device
    lengths: torch.Tensor     # int32 [n], device
    h_offsets: np.ndarray     # int64 [n]
    h_lengths: np.ndarray     # int32 [n]

    @property
    def n(self) -> int:
        return int(self.h_lengths.shape[0])

    @property
    def device(self) -> torch.device:
        return self.data.device

    @property
    def max_len(self) -> int:
        return int(self.h_lengths.max()) if self.n else 0

    @property
    def total_samples(self) -> int:
        return int(self.h_lengths.astype(np.int64).sum())

    def clip(self, i: int, length: int | None = None) -> torch.Tensor:
        o = int(self.h_offsets[i])
        n = int(self.h_lengths[i]) if length is None else int(length)
        return self.data[o:o + n]

    @staticmethod
    def plan_offsets(lengths: Sequence[int]) -> np.ndarray:
        lens = np.asarray(lengths, dtype=np.int64)
        padded = (lens + ALIGN - 1) // ALIGN * ALIGN
        off = np.zeros(len(lens), dtype=np.int64)
        if len(lens) > 1:
            off[1:] = np.cumsum(padded[:-1])
        return off

    @classmethod
    def empty_like_lengths(cls, lengths: Sequence[int], device) -> "RaggedBatch":
        lens = np.asarray(lengths, dtype=np.int32)
        off = cls.plan_offsets(lens)
        total = int(off[-1] + _round_up(int(lens[-1]))) if len(lens) else 0
        data = torch.zeros(max(total, ALIGN), dtype=torch.float32, device=device)
        return cls(data, torch.from_numpy(off).to(device), torch.from_numpy(lens).to(device), off, lens)

    @classmethod
    def from_list(cls, clips: List[torch.Tensor], device) -> "RaggedBatch":
        flat = [c.reshape(-1) for c in clips]
        rb = cls.empty_like_lengths([int(c.numel()) for c in flat], device)
        for i, c in enumerate(flat):
            if c.numel():
                rb.clip(i).copy_(c.to(device=device, dtype=torch.float32), non_blocking=True)
        return rb

    @classmethod
    def from_dense(cls, x: torch.Tensor) -> "RaggedBatch":
        """(B, L) contiguous fp32 on the device; zero-copy when L is a multiple of ALIGN."""
        assert x.dim() == 2 and x.dtype == torch.float32
        B, L = x.shape
        if L % ALIGN == 0 and x.is_contiguous():
            off = np.arange(B, dtype=np.int64) * L
            lens = np.full(B, L, dtype=np.int32)
            return cls(x.reshape(-1), torch.from_numpy(off).to(x.device), torch.from_numpy(lens).to(x.device), off, lens)
        return cls.from_list(list(x), x.device)
