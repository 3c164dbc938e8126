"""Compact view of tools/bench_configs.py lines on stdin: config, ms per step, per-kernel (ms, fraction of HBM peak)."""
import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print(d["config"], d["ms_per_step"], {k: (v["ms"], v.get("frac_hbm")) for k, v in d["kernels"].items()})
