"""does processing C2 in waves that fit the 126 MB L2 (so that the fused kernel's read of x and the
normaliser's read of the frames hit L2 after the scan / the fused kernel touched them) pay?  Times one C2 step as
1000 / wave clips per rho_b200_validate call for several wave sizes (CUDA events); DRAM bytes come from ncu on the same
command (`--metrics dram__bytes_read.sum,dram__bytes_write.sum`).   python tools/wave_traffic.py [wave sizes]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rho_tts_b200 as R
from rho_tts_b200 import synth

dev = torch.device("cuda", 0)
n, L = 1000, 240000
x = synth.make_clip_block(n, L, 0xB200, device=dev)
emb, ref = synth.make_embeddings(n, device=dev)
p = R.make_params()
for wave in [int(a) for a in sys.argv[1:]] or [1000, 500, 250, 125, 100, 64]:
    subs = []
    for w0 in range(0, n, wave):
        sub = R.RaggedBatch.from_dense(x[w0:w0 + wave])
        subs.append((sub, R.ValidatePlan(sub, np.arange(sub.n + 1, dtype=np.int32), p, 80, True), emb[w0:w0 + wave]))

    def step():
        for sub, plan, e in subs:
            plan.run(sub, e, ref)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"wave_clips": wave, "waves": len(subs), "ms_per_step": e0.elapsed_time(e1) / 20,
                      "x_bytes_per_wave_MB": wave * L * 4 / 1e6}), flush=True)
    del subs
