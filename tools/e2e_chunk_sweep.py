"""e2e (host buffers through rho_b200_validate_host) for several chunk sizes: one bench.py subprocess per value of
RHO_HOST_CHUNK.  Developer tool.    python tools/e2e_chunk_sweep.py 8 16 32 64 128"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for ch in sys.argv[1:] or ["16", "32", "64"]:
    env = dict(os.environ, RHO_HOST_CHUNK=ch)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"chunk {ch:>4s}: e2e {d['e2e']['value']:.0f} audio-s/s, features stay in HBM {d['e2e']['value_features_stay_in_hbm']:.0f}", flush=True)
    except Exception as e:   # noqa: BLE001
        print(ch, "FAILED", repr(e), r.stderr[-300:], flush=True)
