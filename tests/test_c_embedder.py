"""The C ABI from plain C: tests/c/host_embedder.c (gcc, include/rho_b200.h only -- no torch, no CUDA headers) drives
rho_b200_validate_host with pageable host buffers; its outputs must be what the Python host mirror produces.
CPU part: the program compiles and links against the in-tree library.  GPU part: it runs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "host_embedder.c")


def _build(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    from rho_tts_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    exe = str(tmp_path / "host_embedder")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run([gcc, "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", libdir,
                        "-l:librho_b200.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_embedder_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)          # no arguments: usage, exit code 2 (no GPU touched)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_c_embedder_matches_python_mirror(tmp_path, cuda_device):
    import torch
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    exe = _build(tmp_path)
    n, L = 70, 48000
    x = synth.make_clip_block(n, L, 321)
    clips = tmp_path / "clips.f32"
    x.numpy().tofile(clips)
    r = subprocess.run([exe, str(clips), str(n), str(L), str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rec_c = np.fromfile(tmp_path / "out.rec", dtype=R.REC_DTYPE)
    y_c = np.fromfile(tmp_path / "out.y", dtype=np.float32).reshape(n, L)
    mel_c = np.fromfile(tmp_path / "out.mel", dtype=np.float32).reshape(n, 80, 3000)
    out = R.validate_batch(R.RaggedBatch.from_dense(x.to(cuda_device)), R.make_params())
    rec = out.records_host()
    for f in ("start", "end", "out_len", "flags", "ok", "n_segments"):
        assert np.array_equal(rec[f], rec_c[f]), f
    assert np.allclose(rec["decay_ratio"], rec_c["decay_ratio"], rtol=1e-6)
    assert np.array_equal(out.mel.cpu().numpy(), mel_c)
    for i in range(n):
        k = int(rec["out_len"][i])
        assert np.array_equal(out.audio.clip(i, k).cpu().numpy(), y_c[i, :k])
    assert f"{n} clips" in r.stdout and "accepted" in r.stdout
