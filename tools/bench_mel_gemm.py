"""Times rho_b200_mel_project (tcgen05 mel projection, in isolation) on 1 M frames (the C2 step has
1000 clips x ~1000 frames), CUDA events, inputs larger than L2.  Prints one JSON line.
Tensor-pipe utilisation comes from ncu (sm__pipe_tensor_cycles_active) on the same command: profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rho_tts_b200 as R

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_mels = int(sys.argv[2]) if len(sys.argv) > 2 else 80
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
fpi = int(sys.argv[4]) if len(sys.argv) > 4 else 1000          # frames per clip (0: one [n_mels, n_frames] matrix)
dev = torch.device("cuda", 0)
p = torch.rand(n_frames, 204, device=dev)
buf = torch.empty(((n_frames + fpi - 1) // fpi, n_mels, fpi), device=dev) if fpi > 0 else None
for _ in range(3):
    out = R.mel_project(p, n_mels, fpi, buf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = R.mel_project(p, n_mels, fpi, buf)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_alg = n_frames * (201 * 4 + n_mels * 4)
flops_useful = 2.0 * n_frames * 201 * n_mels
flops_issued = 2.0 * 128 * 208 * 3 * n_frames            # M padded to 128, K to 208, 3 TF32 products
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
print(json.dumps({"kernel": "k_mel_gemm", "n_frames": n_frames, "n_mels": n_mels, "frames_per_item": fpi, "ms": ms,
                  "frames_per_s": n_frames / ms * 1e3, "hbm_gbs": bytes_alg / ms / 1e6,
                  "hbm_frac": bytes_alg / ms / 1e6 / peaks["hbm_gbs"],
                  "useful_tflops": flops_useful / ms / 1e9, "issued_tf32_tflops": flops_issued / ms / 1e9,
                  "issued_frac_of_tf32_nominal_1100": flops_issued / ms / 1e9 / 1100.0}))
