"""A/B of library variants on config c3 (ragged joins) and the scan kernel of C2: one subprocess per variant."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "rho_tts_b200", "variants", "lib_*.so"))) + [""]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["RHO_B200_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_configs.py"), "c3", "--steps", "10"],
                       env=env, capture_output=True, text=True)
    name = os.path.basename(lib) or "product"
    for ln in r.stdout.strip().splitlines():
        try:
            d = json.loads(ln)
            print(f"{name:24s} {d['config']:4s} step {d['ms_per_step']:.4f}  " +
                  "  ".join(f"{k}={v['ms']}" for k, v in d["kernels"].items() if v["ms"] > 0.05), flush=True)
        except Exception as e:  # noqa: BLE001
            print(name, "FAILED", ln[:200], r.stderr[-300:])
