"""End-to-end (host-buffer) throughput of rho_b200_validate_host on the C2 workload for a sweep of the staging knobs
(steady-state chunk size -- the chunk sizes ramp up from 1/16 of it and down again --, host fill threads, complete /
compact feature rows, no features).  One subprocess per setting (the knobs are read once
per process).    python tools/e2e_sweep.py  >  profiles/e2e_sweep_rNN.log
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    n, L = 1000, 240000
    x = synth.make_clip_block(n, L, 0xB200, device="cuda").cpu().pin_memory()
    emb, ref = synth.make_embeddings(n)
    emb, ref = emb.pin_memory(), ref.pin_memory()
    p = R.make_params()
    y = torch.empty_like(x).pin_memory()
    rec = torch.empty((n, 48), dtype=torch.uint8).pin_memory()
    mode = os.environ.get("SWEEP_MODE", "full")
    if mode in ("full", "full_nofill"):
        mel = torch.empty((n, 80, 3000), dtype=torch.float32).pin_memory()
        fn = lambda: R.validate_host(x, p, emb, ref, 80, y=y, mel=mel, rec=rec)         # noqa: E731
    elif mode == "nomel":
        fn = lambda: R.validate_host(x, p, emb, ref, 80, y=y, mel=None, rec=rec)        # noqa: E731
    else:
        T = int(R._lib.load().rho_b200_compact_frames(L, 3000))
        mel = torch.empty((n, 80, T), dtype=torch.float32).pin_memory()
        so, sl, fi = np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int32), np.arange(n + 1, dtype=np.int32)
        fn = lambda: R.validate_host_ragged(x.reshape(-1), so, sl, fi, p, emb, ref, 80, compact=True, y=y.reshape(-1), mel=mel)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(json.dumps({"mode": mode, "chunk_clips": os.environ.get("SWEEP_CHUNK"), "fill_threads": os.environ.get("RHO_HOST_FILL_THREADS"), "slots": os.environ.get("RHO_HOST_SLOTS"),
                      "ms_median": 1e3 * ts[len(ts) // 2], "ms_best": 1e3 * ts[0], "audio_s_per_s_median": n * 10.0 / ts[len(ts) // 2]}))


if __name__ == "__main__":
    if os.environ.get("SWEEP_CHILD"):
        child()
    else:
        settings = [(m, c, 3, sl) for m in ("full", "compact", "nomel") for c in (32, 64, 128) for sl in (3, 4, 6)]
        settings += [("full", 128, ft, 4) for ft in (1, 2, 4, 6)] + [("full_nofill", 128, 3, 4)]
        if True:
            for mode, chunk, ft, slots in settings:
                if True:
                    env = dict(os.environ, SWEEP_CHILD="1", SWEEP_MODE=mode, SWEEP_CHUNK=str(chunk), RHO_HOST_SLOTS=str(slots),
                               RHO_HOST_CHUNK_SAMPLES=str(chunk * 240000), RHO_HOST_FILL_THREADS=str(ft))
                    if mode == "full_nofill":           # timing experiment only: the rows' constant tails stay unwritten
                        env["RHO_HOST_DEBUG_SKIP_FILL"] = "1"
                    r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
                    print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
