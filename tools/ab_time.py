"""Times every rho_tts_b200/variants/lib_*.so on the C2 workload (one subprocess per variant; the
product library last).  Developer tool: `tools/ab_fused.sh` builds the variants."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "rho_tts_b200", "variants", "lib_*.so"))) + [""]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["RHO_B200_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "50", "--warmup", "5", "--no-e2e",
                        "--no-cpu-baseline"] + sys.argv[1:], env=env, capture_output=True, text=True)
    name = os.path.basename(lib) or "product"
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        ks = {k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items() if v["ms_per_launch"] > 0.05}
        print(f"{name:28s} step {d['ms_per_step']:.4f} ms  {ks}", flush=True)
    except Exception as e:   # noqa: BLE001
        print(name, "FAILED", repr(e), r.stderr[-500:], flush=True)
