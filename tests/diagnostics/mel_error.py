"""Developer diagnostic: worst log-mel error of the fused and the unfused GPU paths against the fp64 oracle
chain, and against each other (40 clips x 5 s)."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
import rho_tts_b200 as R
from rho_tts_b200 import synth
dev = torch.device("cuda", 0)
n = 24
x = synth.make_clip_block(n, 120000, 31)
emb, ref = synth.make_embeddings(n)
rb = R.RaggedBatch.from_dense(x.to(dev))
res = {}
for fuse in (True, False):
    out = R.validate_batch(rb, R.make_params(), emb.to(dev), ref.to(dev), n_mels=80, pad_to_30s=True, fuse=fuse)
    res[fuse] = out.mel.cpu().numpy().astype(np.float64)
c = oracle.derive_constants()


def truth64(y):
    """Same tables (fp32 taps / window / filterbank), all arithmetic in float64."""
    from oracle.resample import sinc_resample_kernel
    from oracle.logmel import hann_periodic, slaney_mel_filterbank
    taps, width, orig, new = sinc_resample_kernel(24000, 16000)
    L = y.size
    xp = np.zeros(L + 2 * width + orig); xp[width:width + L] = y
    fr = np.lib.stride_tricks.sliding_window_view(xp, taps.shape[1])[::orig]
    w = (fr @ taps.astype(np.float64).T).reshape(-1)[:-(-new * L // orig)]
    buf = np.zeros(480000); buf[:min(w.size, 480000)] = w[:480000]
    p = np.concatenate([buf[200:0:-1], buf, buf[-2:-202:-1]])
    frames = np.lib.stride_tricks.sliding_window_view(p, 400)[::160][:3000]
    power = np.abs(np.fft.rfft(frames * hann_periodic().astype(np.float64)[None, :], axis=1)) ** 2
    mel = slaney_mel_filterbank(80).astype(np.float32).astype(np.float64).T @ power.T
    ls = np.log10(np.maximum(mel, 1e-10)); ls = np.maximum(ls, ls.max() - 8.0)
    return (ls + 4.0) / 4.0


worst = {True: 0.0, False: 0.0}
w64 = {True: 0.0, False: 0.0, "oracle": 0.0}
for i in range(n):
    o = oracle.post_process_clip(x[i].numpy(), c)
    m = oracle.log_mel(oracle.resample(o["audio"]), 80, True).astype(np.float64)
    t64 = truth64(o["audio"].astype(np.float64))
    w64["oracle"] = max(w64["oracle"], float(np.abs(m - t64).max()))
    for fuse in (True, False):
        worst[fuse] = max(worst[fuse], float(np.abs(res[fuse][i] - m).max()))
        w64[fuse] = max(w64[fuse], float(np.abs(res[fuse][i] - t64).max()))
print("max |gpu - oracle|: fused %.3e  unfused %.3e   fused vs unfused %.3e" %
      (worst[True], worst[False], float(np.abs(res[True] - res[False]).max())))
print("max |. - fp64 truth|: fused %.3e  unfused %.3e  numpy oracle %.3e" % (w64[True], w64[False], w64["oracle"]))
