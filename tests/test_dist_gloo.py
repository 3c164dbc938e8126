"""CPU, world_size 2 over gloo: the N > 1 host logic -- contiguous clip shards per rank, no data-path
collective, one all-gather of the 48-byte records, rank-order concatenation."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _worker(rank, world, port, n_items, q):
    import torch.distributed as dist
    from rho_tts_b200 import REC_DTYPE
    from rho_tts_b200 import dist as rdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = rdist.shard_range(n_items, rank, world)
        rec = np.zeros(hi - lo, dtype=REC_DTYPE)
        rec["start"] = np.arange(lo, hi)            # stands in for what the kernels write for these clips
        rec["ok"] = (np.arange(lo, hi) % 3 == 0)
        rec["decay_ratio"] = np.arange(lo, hi) * 0.5
        local = torch.from_numpy(rec.view(np.uint8).reshape(-1, 48).copy())
        counts = [b - a for a, b in (rdist.shard_range(n_items, r, world) for r in range(world))]
        allr = rdist.gather_records_ragged(local, counts)
        got = allr.numpy().view(REC_DTYPE).reshape(-1)
        ok = (got.shape[0] == n_items and np.array_equal(got["start"], np.arange(n_items))
              and np.array_equal(got["ok"], (np.arange(n_items) % 3 == 0).astype(np.int32))
              and np.array_equal(got["decay_ratio"], np.arange(n_items) * 0.5))
        if counts[0] == counts[-1]:
            same = rdist.gather_records(local).numpy().view(REC_DTYPE).reshape(-1)
            ok = ok and np.array_equal(same["start"], np.arange(n_items))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [10, 11])
def test_two_rank_record_gather(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_items
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
