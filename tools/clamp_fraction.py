import torch, sys
sys.path.insert(0, ".")
import rho_tts_b200 as R
from rho_tts_b200 import synth
dev = torch.device("cuda", 0)
x = synth.make_clip_block(200, 240000, 0xB200, device=dev)
out = R.validate_batch(R.RaggedBatch.from_dense(x), R.make_params())
mel = out.mel[:, :, :1000]
floor = (out.mel.amax(dim=(1, 2)) - 2.0)[:, None, None]
cl = (mel <= floor + 1e-7)
print("clamped elements: %.4f" % cl.float().mean().item())
v4 = cl.reshape(200, 80, 250, 4).any(dim=3)
print("float4 vectors with a clamped element: %.4f" % v4.float().mean().item())
print("by mel row (first 8, last 4):", [round(v, 3) for v in cl.float().mean(dim=(0, 2))[:8].tolist()], [round(v, 3) for v in cl.float().mean(dim=(0, 2))[-4:].tolist()])
fr = cl.any(dim=1)                                   # [clips, 1000] frames with a clamped value
print("frames with a clamped value: %.4f" % fr.float().mean().item())
for bs in (8, 16, 32, 64, 128):
    nb = 1000 // bs
    blk = fr[:, :nb * bs].reshape(200, nb, bs).any(dim=2)
    print("blocks of %3d frames with a clamped value: %.4f" % (bs, blk.float().mean().item()))
