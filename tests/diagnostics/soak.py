"""Randomised soak of the batched entry points against the oracle: fresh seeds, ragged lengths from a few samples to
several seconds, random item partitions, both feature sizes.  Developer diagnostic (GPU box):
    python tests/diagnostics/soak.py [rounds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
import rho_tts_b200 as R  # noqa: E402
from rho_tts_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
c = oracle.derive_constants()
p = R.make_params()
worst = {"audio": 0.0, "mel": 0.0, "ratio": 0.0}
bad = 0
for r in range(rounds):
    rng = np.random.default_rng(900000 + r)
    n = int(rng.integers(3, 14))
    lens = np.where(rng.random(n) < 0.25, rng.integers(1, 3000, n), rng.integers(3000, 150000, n)).astype(np.int32)
    clips = [t.numpy() for t in synth.make_clips(lens, 5000 + r)]
    rb = R.RaggedBatch.from_list([torch.from_numpy(x) for x in clips], dev)
    # (a) joins of random partitions
    first = synth.make_item_partition(n, 77 + r, 1, 4)
    out = R.join_batch(rb, first, p, want_seg_info=False)
    rec = out.records_host()
    for i in range(len(first) - 1):
        o = oracle.smooth_segment_join(clips[first[i]:first[i + 1]], c)
        L = int(rec["out_len"][i])
        if o.audio is None:
            continue
        if L != o.audio.size:
            bad += 1; print("round", r, "item", i, "length", L, o.audio.size); continue
        y = out.audio.clip(i, L).cpu().numpy()
        if L:
            worst["audio"] = max(worst["audio"], float(np.abs(y - o.audio.reshape(-1)).max()))
        ratio, ok, _, _ = oracle.sound_decay(o.audio.reshape(-1), 0.3)
        worst["ratio"] = max(worst["ratio"], abs(ratio - rec["decay_ratio"][i]) / max(1.0, abs(ratio)))
        if bool(rec["ok"][i]) != ok and abs(ratio - 0.3) > 1e-4:
            bad += 1; print("round", r, "item", i, "decision", rec["ok"][i], ok, ratio)
    # (a2) the same joined items through the whole front end (join -> features read from the joined audio -> cosine)
    nmj = 128 if r % 2 == 0 else 80
    padj = (r % 3 != 1)
    vj = R.validate_batch(rb, p, n_mels=nmj, pad_to_30s=padj, item_first_seg=first)
    vjr = vj.records_host()
    for i in range(len(first) - 1):
        o = oracle.smooth_segment_join(clips[first[i]:first[i + 1]], c)
        if o.audio is None:
            continue
        if int(vjr["out_len"][i]) != o.audio.size or int(vjr["out_len"][i]) != int(rec["out_len"][i]):
            bad += 1; print("round", r, "item", i, "joined-features length", vjr["out_len"][i], o.audio.size); continue
        w = oracle.resample(o.audio.reshape(-1)) if o.audio.size else np.zeros(0, np.float32)
        if not padj and w.size <= 200:
            continue
        m = oracle.log_mel(w, nmj, padj)
        g = vj.mel[i, :, :m.shape[1]].cpu().numpy()
        worst["mel"] = max(worst["mel"], float(np.abs(g - m).max()))
    # (b) one-segment items through the fused path, both feature sizes
    nm = 80 if r % 2 == 0 else 128
    v = R.validate_batch(rb, p, n_mels=nm, pad_to_30s=(r % 3 != 0))
    vr = v.records_host()
    for i in range(n):
        o = oracle.post_process_clip(clips[i], c)
        if (vr["start"][i], vr["end"][i]) != (o["start"], o["end"]):
            bad += 1; print("round", r, "clip", i, "bounds", vr["start"][i], vr["end"][i], o["start"], o["end"]); continue
        w = oracle.resample(o["audio"])
        if w.size <= 200 and (r % 3 == 0):
            continue                                     # unpadded reflect padding needs > 200 samples
        m = oracle.log_mel(w, nm, r % 3 != 0)
        T = m.shape[1]
        g = v.mel[i, :, :T].cpu().numpy()
        worst["mel"] = max(worst["mel"], float(np.abs(g - m).max()))
print("rounds", rounds, "mismatches", bad, "worst abs errors", worst)
sys.exit(1 if bad or worst["audio"] > 1e-4 or worst["mel"] > 1e-4 or worst["ratio"] > 1e-4 else 0)
