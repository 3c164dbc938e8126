"""Error of the GPU pitch shift against (a) the oracle (fp32, torch's roundings) and (b) torchaudio on the same box,
in fp32 and in fp64 -- how much of the distance is the reference's own fp32 noise.  GPU box only (torchaudio is in
the image; /root/reference is not needed).    python tests/diagnostics/pitch_error.py"""
import os
import sys
import time

import numpy as np
import torch
import torchaudio

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "golden"))
import rho_tts_b200 as R  # noqa: E402
from oracle import pitch as OP  # noqa: E402
from pitch_inputs import pitch_input  # noqa: E402
from rho_tts_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
for name, x in (("voiced 3 s", pitch_input(72000, 7)), ("voiced 10 s", pitch_input(240000, 8)),
                ("synthetic bench clip 10 s", synth.make_clips([240000], 5)[0].numpy())):
    for steps in (2.0, -3.0, 0.5):
        rb = R.RaggedBatch.from_list([torch.from_numpy(x)], dev)
        y = R.pitch_shift_batch(rb, 24000, steps).clip(0).cpu().numpy()
        t0 = time.perf_counter()
        t32 = torchaudio.functional.pitch_shift(torch.from_numpy(x)[None], 24000, steps)[0].numpy()
        dt = time.perf_counter() - t0
        t64 = torchaudio.functional.pitch_shift(torch.from_numpy(x)[None].double(), 24000, steps)[0].numpy()
        o = OP.pitch_shift(x, 24000, steps)
        print(f"{name:28s} n_steps {steps:+.1f}: gpu-torch32 {np.abs(y - t32).max():.2e}  gpu-oracle {np.abs(y - o).max():.2e}  "
              f"oracle-torch32 {np.abs(o - t32).max():.2e}  torch32-torch64 {np.abs(t32 - t64).max():.2e}  "
              f"(peak {np.abs(t32).max():.2f}; torchaudio CPU {dt:.2f} s)", flush=True)
