"""Shared helpers for the parity tests."""
import numpy as np

TOL = 1e-4   # north_star: values within 1e-4 relative (fp32); abs-or-rel form from SURVEY.md 8(d)


def assert_close(a, b, tol=TOL, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return 0.0
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    worst = float(err.max())
    assert worst <= tol, f"{what}: max |a-b|/max(1,|b|) = {worst:.3e} > {tol:g} at {int(err.argmax())}"
    return worst


def tone_clip(rng, L, lead=0, trail=0, amp=0.3, dc=1e-3, sr=24000, decay_to=None):
    t = np.arange(L) / sr
    env = np.linspace(1.0, rng.uniform(0.1, 1.2) if decay_to is None else decay_to, L) if L else np.zeros(0)
    x = amp * np.sin(2 * np.pi * rng.uniform(90, 300) * t) * env
    if lead:
        x[:lead] = 0
    if trail:
        x[L - trail:] = 0
    x = x + rng.normal(0, 1e-3, L) + dc
    return x.astype(np.float32)
