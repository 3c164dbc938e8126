"""Generates tests/golden/golden_cf0_v1.npz: BaseTTS._smooth_segment_join with the crossfade DISABLED
(crossfade_duration_sec = 0, and a 0.0004 s one that gives 9 samples <= 10), run through the REFERENCE itself.

With crossfade_samples == 0 the reference's `current_segment[..., :-crossfade_samples]` (base_tts.py:485) is the
empty slice [..., :0]: segment 0 is dropped from the result.  The drop-in keeps that bug for bug; these vectors pin it.

    python tests/golden/make_golden_cf0.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/src")

import torch  # noqa: E402
from make_golden import Ref, make_inputs  # noqa: E402


def main():
    clips, _ = make_inputs()
    items = [[0, 1, 2], [3, 4], [5, 3], [5, 5, 5], [3, 5], [6, 7, 8, 0], [2, 2]]
    out = {"n_items": np.int32(len(items)), "xfade_secs": np.asarray([0.0, 0.0004])}
    for v, xf in enumerate(out["xfade_secs"]):
        ref = Ref()
        ref.crossfade_duration_sec = float(xf)
        for k, idx in enumerate(items):
            y = ref._smooth_segment_join([torch.from_numpy(clips[j].copy()) for j in idx])
            out[f"v{v}_item{k}_idx"] = np.asarray(idx, dtype=np.int32)
            out[f"v{v}_item{k}"] = y.numpy().reshape(-1).copy()
            out[f"v{v}_item_dim{k}"] = np.int32(y.dim())
    path = os.path.join(HERE, "golden_cf0_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
