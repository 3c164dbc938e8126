"""GPU, needs TWO devices (skipped on a one-GPU box; run with `gpurun --gpus 2`):
  * a handle on device 1 while the process's current device is 0 (every entry point guards the device), and
  * the record exchange of one process per GPU: records stored into every rank's gathered buffer by the kernel that
    assembles them (CUDA IPC peer memory), against an NCCL all-gather of the same records.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 CUDA devices, this box has {torch.cuda.device_count() if torch.cuda.is_available() else 0}")


def test_handle_on_device_1_while_current_device_is_0():
    _need_two()
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    torch.cuda.set_device(0)
    x = synth.make_clip_block(24, 120000, 5)
    emb, ref = synth.make_embeddings(24)
    p = R.make_params()
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        rb = R.RaggedBatch.from_dense(x.to(dev))
        o = R.validate_batch(rb, p, emb.to(dev), ref.to(dev))
        assert torch.cuda.current_device() == 0              # the caller's current device is left alone
        outs.append((o.records_host(), o.mel.cpu(), o.audio.data.cpu()))
        j = R.join_batch(rb, [0, 3, 10, 24], p)
        assert torch.cuda.current_device() == 0 and j.records.device == dev
    assert outs[0][0].tobytes() == outs[1][0].tobytes()
    assert torch.equal(outs[0][1], outs[1][1])
    # the host entry point on device 1
    xh = x.pin_memory()
    y, _, rec = R.validate_host(xh, p, None, None, 80, device=1)
    assert torch.cuda.current_device() == 0
    rh = rec.numpy().view(R.REC_DTYPE).reshape(-1)
    assert np.array_equal(rh["out_len"], outs[0][0]["out_len"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange_worker(rank, world, port, force_nccl, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        import rho_tts_b200 as R
        from rho_tts_b200 import synth
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        n = 96
        x = synth.make_clip_block(n, 60000, 100 + rank, device=dev)
        emb, ref = synth.make_embeddings(n, 256, 7 + rank, device=dev)
        rb = R.RaggedBatch.from_dense(x)
        plan = R.ValidatePlan(rb, np.arange(n + 1, dtype=np.int32), R.make_params(), 80, True)
        ex = R.dist.RecordExchange(rank, n, force_nccl=force_nccl)
        ok = True
        for step in range(5):
            if step == 3:                                   # different data on the way: the epochs must not mix
                x.mul_(0.5)
            out = plan.run(rb, emb, ref)
            ex.after_step(out.records)
            got = ex.gathered()
            want = R.dist.gather_records(out.records)        # NCCL all-gather of the same records
            torch.cuda.synchronize()
            ok = ok and torch.equal(got.cpu(), want.cpu())
        mode, why = ex.mode, ex.why
        ex.close()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, mode, why))
    except Exception as e:          # noqa: BLE001
        q.put((rank, False, "error", repr(e)))


@pytest.mark.parametrize("force_nccl", [False, True])
def test_record_exchange_two_ranks(force_nccl):
    _need_two()
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, force_nccl, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(timeout=60) for p in procs]
    for rank, ok, mode, why in res:
        assert ok, (rank, mode, why)
    modes = {m for _, _, m, _ in res}
    assert len(modes) == 1
    print(f"record exchange mode: {modes.pop()} (fallback reason: {res[0][3] or 'none'})")
    if force_nccl:
        assert res[0][2] == "nccl"


def test_record_exchange_single_rank(cuda_device):
    """World size 1 (runs on a one-GPU box): the exchange with only the local sink -- records stored by the assembling
    kernel into the gathered buffer, epoch flag, wait kernel, read-back -- against the records themselves."""
    import torch.distributed as dist
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    if dist.is_initialized():
        pytest.skip("a process group is already initialised in this process")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), RANK="0", WORLD_SIZE="1")
    dist.init_process_group("nccl", device_id=cuda_device)
    try:
        n = 64
        x = synth.make_clip_block(n, 60000, 3, device=cuda_device)
        rb = R.RaggedBatch.from_dense(x)
        plan = R.ValidatePlan(rb, np.arange(n + 1, dtype=np.int32), R.make_params(), 80, True)
        ex = R.dist.RecordExchange(0, n)
        assert ex.mode == "p2p", ex.why
        for step in range(4):
            x.mul_(0.9)
            out = plan.run(rb, None, None)
            ex.after_step(out.records)
            got = ex.gathered()
            torch.cuda.synchronize()
            assert torch.equal(got, out.records), step
        ex.close()
    finally:
        dist.destroy_process_group()
