"""CPU oracle for the rho-tts audio hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's algorithm for the path
named by BASELINE.json `north_star`.  It exists so that the
CUDA path can be checked against something that runs on a box where
`/root/reference` is absent.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
    `--impl reference` legs may import this package.
  * Nothing under `rho_tts_b200/` may import, call or link it.  The product path
    has no CPU fallback: it raises when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * in-tree stages (trim / DC / fades / join / decay / cosine): the reference
    ships NO golden vectors.  The oracle is pinned against
    outputs of the reference itself, generated in the authoring container by
    `tests/golden/make_golden.py` (imports `rho_tts` from /root/reference/src)
    and committed under `tests/golden/`.
  * resample (torchaudio, unpinned `>=2.0` in the reference's pyproject) and
    Whisper log-mel (transformers, unpinned `>=4.40`): the algorithm lives in
    third-party packages that are not under /root/reference.  Installed
    versions torchaudio 2.11.0 / transformers 5.5.0 are the pin; golden vectors
    from those exact functions are committed, and when the packages import the
    tests also compare live.
"""

from .dsp import (  # noqa: F401
    TrimResult,
    derive_constants,
    frame_energy,
    trim_bounds,
    trim_silence,
    remove_dc_offset,
    fade_curves,
    apply_fades,
    crossfade_curves,
    join_plan,
    smooth_segment_join,
    sound_decay,
    cosine_similarity,
    post_process_clip,
)
from .resample import sinc_resample_kernel, resample  # noqa: F401
from .logmel import slaney_mel_filterbank, hann_periodic, stft_power, log_mel, log_mel_truth64  # noqa: F401
