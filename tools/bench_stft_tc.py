"""Times rho_b200_stft_power_tc (the windowed DFT as two tcgen05 GEMMs, in isolation) on the 16 kHz signal of the C2
workload (1000 clips x 10 s -> ~1 M frames), next to the product path's FFT kernel on the same input
(k_logmel_frames: FFT + mel + log of the same frames), CUDA events, inputs larger than L2.  One JSON line.
Tensor-pipe utilisation comes from ncu (sm__pipe_tensor_cycles_active) on the same command: profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rho_tts_b200 as R
from rho_tts_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
x = synth.make_clip_block(n, 240000, 0xB200, device=dev)
rb16 = R.resample_batch(R.RaggedBatch.from_dense(x))
del x


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# the tile list and the output buffer are built once (host work outside the timed region)
h = R._lib.Handle.get(0)
n16 = rb16.h_lengths.astype(np.int64)
T = n16 // 160
base = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
tiles = np.concatenate([np.stack([np.full(len(t0), c), t0, np.minimum(8, T[c] - t0), base[c] + t0], axis=1)
                        for c in range(n) for t0 in [np.arange(0, T[c], 8)]]).astype(np.int32)
tl = torch.from_numpy(tiles).to(dev)
power = torch.empty((int(base[-1]), 208), dtype=torch.float32, device=dev)
from rho_tts_b200.batch import _ptr, _stream
from rho_tts_b200 import _lib


def run_tc():
    _lib.check(h.lib.rho_b200_stft_power_tc(h.ptr, _ptr(rb16.data), _ptr(rb16.offsets), _ptr(rb16.lengths), 0, _ptr(tl),
                                            int(tl.shape[0]), _ptr(power), 208, _stream(0)), "stft_power_tc")


ms_tc = timed(run_tc)
h.profile_begin()
for _ in range(3):
    R.logmel_batch(rb16, 80, False, lengths=rb16.lengths)
prof = h.profile_end()
ms_fft = prof["k_logmel_frames"][0] / prof["k_logmel_frames"][1]
frames = int(base[-1])
mma_flops = frames / 8 * 24 * 2.0 * 128 * 32 * 8            # 24 M128 N32 K8 MMAs per 8 frames
print(json.dumps({"kernel": "k_stft_tc", "clips": n, "frames": frames, "ms": ms_tc, "frames_per_s": frames / ms_tc * 1e3,
                  "issued_tf32_tflops": mma_flops / ms_tc / 1e9, "issued_frac_of_tf32_nominal_1100": mma_flops / ms_tc / 1e9 / 1100.0,
                  "hbm_gbs": (4.0 * float(n16.sum()) + 4.0 * 201 * frames) / ms_tc / 1e6,
                  "product_path_k_logmel_frames_ms": ms_fft,
                  "note": "k_logmel_frames = shared-memory FFT + sparse mel + log of the same frames in ONE kernel; "
                          "k_stft_tc produces only the power spectra (0.8 GB written) and still needs rho_b200_mel_project"}))
