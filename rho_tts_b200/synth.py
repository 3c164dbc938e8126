"""Synthetic 24 kHz clip generator.

x[t] = env(t) * am(t) * sum_h a_h sin(2 pi h f0 t / sr) * gate(t) + noise + dc
  f0 ~ U[90, 300] Hz, a = (0.25, 0.12, 0.06), am = 0.6 + 0.4 sin(2 pi r t), r ~ U[3, 6] Hz,
  env linear 1 -> rho, rho ~ max(0.02, U[-0.15, 1.2]) (about a fifth of the clips fail the 0.3 decay check),
  noise ~ N(0, 1e-3^2) (-60 dBFS), dc ~ U[-2e-3, 2e-3], leading silence U[0, 0.5] s,
  trailing silence U[0, 1.0] s (scaled down for clips under 3 s), 5 ms linear onset / offset.

Parity sets: device="cpu" with torch.Generator().manual_seed(seed) -> bit-identical inputs for the
oracle and the GPU.  Throughput sets: the same code on the CUDA device (Philox), generated outside
any timed region and never crossing PCIe.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch

SR = 24000


def make_clip_block(n: int, length: int, seed: int, device="cpu", sr: int = SR) -> torch.Tensor:
    """(n, length) fp32 clips."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))

    def U(lo, hi):
        return lo + (hi - lo) * torch.rand(n, 1, generator=g, device=dev, dtype=torch.float32)

    f0, r, rho, dc = U(90.0, 300.0), U(3.0, 6.0), torch.clamp(U(-0.15, 1.2), min=0.02), U(-2e-3, 2e-3)
    dur = length / sr
    lead_s = U(0.0, 0.5) * min(1.0, dur / 3.0)
    trail_s = U(0.0, 1.0) * min(1.0, dur / 3.0)
    t = torch.arange(length, device=dev, dtype=torch.float32).unsqueeze(0) / sr       # (1, L)
    two_pi = 2.0 * math.pi
    tone = 0.25 * torch.sin(two_pi * f0 * t) + 0.12 * torch.sin(two_pi * 2.0 * f0 * t) \
        + 0.06 * torch.sin(two_pi * 3.0 * f0 * t)
    am = 0.6 + 0.4 * torch.sin(two_pi * r * t)
    env = 1.0 + (rho - 1.0) * (t / max(dur, 1e-9))
    on, off_t = lead_s, dur - trail_s
    ramp = 0.005
    gate = torch.clamp((t - on) / ramp, 0.0, 1.0) * torch.clamp((off_t - t) / ramp, 0.0, 1.0)
    x = env * am * tone * gate
    x = x + 1e-3 * torch.randn(n, length, generator=g, device=dev, dtype=torch.float32) + dc
    return x.contiguous()


def make_ragged_lengths(n: int, seed: int, lo_s: float = 1.0, hi_s: float = 30.0, sr: int = SR) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.rint(rng.uniform(lo_s, hi_s, n) * sr).astype(np.int32)


def make_clips(lengths: Sequence[int], seed: int, device="cpu", sr: int = SR) -> List[torch.Tensor]:
    """One 1-D fp32 tensor per requested length."""
    return [make_clip_block(1, int(L), seed * 100003 + i, device, sr)[0] for i, L in enumerate(lengths)]


def make_item_partition(n_segments: int, seed: int, lo: int = 2, hi: int = 6) -> np.ndarray:
    """item_first_seg for grouping consecutive segments into items of k ~ U{lo..hi} segments."""
    rng = np.random.default_rng(seed)
    first = [0]
    while first[-1] < n_segments:
        first.append(min(n_segments, first[-1] + int(rng.integers(lo, hi + 1))))
    return np.asarray(first, dtype=np.int32)


def make_embeddings(n: int, dim: int = 256, seed: int = 4321, device="cpu"):
    """ReLU + L2-normed speaker-embedding look-alikes: (emb [n, dim], ref [dim])."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    e = torch.relu(torch.randn(n + 1, dim, generator=g, device=dev, dtype=torch.float32))
    e = e / e.norm(dim=1, keepdim=True)
    return e[1:].contiguous(), e[0].contiguous()
