"""CPU: the C-ABI library loads, exports every symbol include/rho_b200.h declares, builds the same
constant tables as torchaudio / transformers, and fails loudly without a GPU.  No kernels run here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import oracle
from rho_tts_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rho_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rho_b200_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rho_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)


def test_abi_version_and_struct_sizes(lib):
    assert lib.rho_b200_abi_version() == 2 == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.RhoRecord) == 48 and ctypes.sizeof(_lib.RhoSegInfo) == 16
    assert ctypes.sizeof(_lib.RhoParams) == 48
    from rho_tts_b200 import REC_DTYPE, SEG_DTYPE
    assert REC_DTYPE.itemsize == 48 and SEG_DTYPE.itemsize == 16
    for (name, _), f in zip(_lib.RhoRecord._fields_, REC_DTYPE.names):
        assert name == f
        assert getattr(_lib.RhoRecord, name).offset == REC_DTYPE.fields[f][1]


def _table(lib, kind, arg, n):
    a = np.zeros(n, np.float32)
    assert lib.rho_b200_host_table(kind, arg, a.ctypes.data, n) == n
    return a


def test_resample_taps_match_torchaudio(lib):
    taps = _table(lib, 0, 0, 46).reshape(2, 23)
    k, width, orig, new = oracle.sinc_resample_kernel(24000, 16000)
    assert (width, orig, new) == (10, 3, 2)
    assert np.abs(taps - k).max() < 1e-7
    ta = pytest.importorskip("torchaudio")
    kt, wt = ta.functional.functional._get_sinc_resample_kernel(24000, 16000, 8000, dtype=torch.float32)
    assert wt == 10 and np.abs(taps - kt.numpy()[:, 0, :]).max() < 1e-7
    # the clamp positions the kernel may skip are numerically nil
    nil = np.abs(taps) < 1e-20
    assert nil[0].nonzero()[0].tolist() == [0, 20, 21, 22] and nil[1].nonzero()[0].tolist() == [0, 1, 2, 21, 22]


def test_hann_and_mel_tables(lib):
    h = _table(lib, 1, 0, 400)
    assert np.abs(h - torch.hann_window(400).numpy()).max() < 5e-7
    for nm in (80, 128):
        m = _table(lib, 2, nm, nm * 201).reshape(nm, 201)
        want = oracle.slaney_mel_filterbank(nm).astype(np.float32).T
        assert np.array_equal(m, want)
        try:
            from transformers import WhisperFeatureExtractor
        except Exception:            # noqa: BLE001
            continue
        assert np.array_equal(m, WhisperFeatureExtractor(feature_size=nm).mel_filters.astype(np.float32).T)
    assert lib.rho_b200_host_table(2, 64, h.ctypes.data, 400) < 0
    assert b"n_mels" in lib.rho_b200_last_error()


def test_compact_frames(lib):
    """Row length of a compact feature tensor: the frames of the 30 s window that can see signal (SURVEY App. A.9:
    t < ceil((L16 + 200) / 160)), rounded up to whole 128-bit pieces; the unpadded frame count with pad_frames = 0."""
    for L in (1, 100, 24000, 240000, 240001, 719999, 720000, 2000000):
        L16 = -(-2 * L // 3)
        t_real = min(3000, max(2, -(-(min(L16, 480000) + 200) // 160)))
        assert lib.rho_b200_compact_frames(L, 3000) == min(3000, (t_real + 3) // 4 * 4), L
        assert lib.rho_b200_compact_frames(L, 0) == L16 // 160
    assert lib.rho_b200_compact_frames(240000, 3000) == 1004


def test_host_entry_points_reject_bad_layouts_without_a_gpu(lib):
    """Argument / layout checks of the host entry points run before anything touches the device."""
    x = np.zeros(1000, np.float32)
    off = np.array([0, 400], np.int64); ln = np.array([500, 400], np.int32)          # overlapping segments
    first = np.array([0, 1, 2], np.int32); yoff = np.array([0, 512], np.int64)
    rec = np.zeros(2 * 48, np.uint8)
    p = _lib.RhoParams(24000, 1, -50.0, 0.02, 0.05, 0.1, 0.3)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)     # noqa: E731
    assert lib.rho_b200_validate_host_ragged(None, vp(x), vp(off), vp(ln), 2, vp(first), 2, ctypes.byref(p), vp(x), vp(yoff),
                                             80, 3000, None, 3000, None, None, None, 0, vp(rec)) == -1
    assert b"handle is NULL" in lib.rho_b200_last_error()
    assert lib.rho_b200_exchange_wait(None, 1, None) == -1 and lib.rho_b200_exchange_epoch(None) == -1


def test_workspace_bytes_monotone(lib):
    a = lib.rho_b200_workspace_bytes(10, 10, 240000)
    b = lib.rho_b200_workspace_bytes(1000, 1000, 240000)
    c = lib.rho_b200_workspace_bytes(1000, 1000, 720000)
    assert 0 < a < b < c


@pytest.mark.skipif(torch.cuda.is_available(), reason="box has a GPU")
def test_create_fails_loudly_without_gpu(lib):
    h = ctypes.c_void_p()
    rc = lib.rho_b200_create(ctypes.byref(h), 0)
    assert rc == -3 and not h.value
    assert b"no CPU fallback" in lib.rho_b200_last_error()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _lib.Handle(0)


def test_baked_mel_weights_match_runtime_table(lib):
    """mel_sparse_gen.inc bakes the filterbank into FFMA immediates; they must be the table, bit for bit."""
    text = open(os.path.join(ROOT, "rho_tts_b200", "csrc", "mel_sparse_gen.inc")).read()
    for nm in (80, 128, 40):
        body = text[text.index(f"void mel_sparse_{nm}("):]
        body = body[:body.index("static const unsigned int")]
        table = _table(lib, 2, nm, nm * 201).reshape(nm, 201)
        seen = np.zeros_like(table)
        rows = re.findall(r"\{ float a = 0\.f;(.*?) emit\((\d+), a\); \}", body)
        assert len(rows) == nm
        for taps, m in rows:
            for hx, k in re.findall(r"__uint_as_float\((0x[0-9a-f]{8})u\), p\[(\d+)\]", taps):
                seen[int(m), int(k)] = np.array([int(hx, 16)], dtype=np.uint32).view(np.float32)[0]
        assert np.array_equal(seen.view(np.uint32), table.view(np.uint32))


def test_fused_mel_stream_is_the_filterbank(lib):
    """The filterbank as the fused kernel walks it (a stream of float4 in constant memory: bundles of rows, headers,
    interleaved zero-padded groups of four weights, x 1/4) decodes back to the table of rho_b200_host_table(2), every row
    exactly once, split over the ten warps of a half; the padding never leaves the 201 bins."""
    for nm in (80, 128):
        stream = np.zeros(2560, np.float32)
        part = np.zeros(11, np.int32)
        part4 = np.zeros(11, np.int32)
        rows = ctypes.c_int(0)
        n = lib.rho_b200_host_mel_stream(nm, stream.ctypes.data, stream.size, part.ctypes.data, part4.ctypes.data,
                                         ctypes.byref(rows))
        assert n > 0 and n % 4 == 0 and rows.value in (1, 2, 4)
        R = rows.value
        hdr4 = (R + 1 + 3) // 4
        table = _table(lib, 2, nm, nm * 201).reshape(nm, 201)
        got = np.zeros_like(table)
        seen = np.zeros(nm, np.int32)
        assert part[0] == 0 and part[10] == nm and np.all(np.diff(part) >= 0) and part4[0] == 0 and part4[10] == n // 4
        words = stream.view(np.uint32)
        for w in range(10):
            pos = int(part4[w])
            for m0 in range(int(part[w]), int(part[w + 1]), R):
                first = words[4 * pos:4 * pos + R] // 4
                nbytes = int(words[4 * pos + R])
                assert nbytes % 16 == 0
                groups = nbytes // 16
                body = stream[4 * (pos + hdr4):4 * (pos + hdr4 + R * groups)].reshape(groups, R, 4)
                for r in range(R):
                    if m0 + r < part[w + 1]:
                        f = int(first[r])
                        assert 0 <= f and f + 4 * groups <= 201
                        got[m0 + r, f:f + 4 * groups] += body[:, r, :].reshape(-1)
                        seen[m0 + r] += 1
                    else:
                        assert not body[:, r, :].any()                      # filler rows of the last bundle: all zero
                pos += hdr4 + R * groups
            assert pos == part4[w + 1]
        assert np.all(seen == 1)
        assert np.array_equal(got, (0.25 * table).astype(np.float32))          # exact: a power of two
    assert lib.rho_b200_host_mel_stream(40, None, 0, None, None, None) < 0


def test_pitch_and_mfcc_host_tables(lib):
    """The tables the pitch-shift / MFCC kernels get from the host, against torch / the oracle (no GPU needed)."""
    import math
    from oracle import mfcc as OM
    from oracle import pitch as OP
    padv = _table(lib, 3, 0, 257)
    assert np.array_equal(padv, torch.linspace(0, math.pi * 128, 257).numpy())       # bit for bit: the phases depend on it
    assert np.array_equal(padv, OP.linspace_f32(math.pi * 128, 257))
    dct = _table(lib, 4, 0, 13 * 128).reshape(13, 128)
    assert float(np.abs(dct - OM.dct_matrix()).max()) <= 1e-7
    fb = _table(lib, 5, 0, 128 * 1025).reshape(128, 1025)
    assert float(np.abs(fb - OM.mel_filterbank()).max()) <= 1e-8
    # windowed taps of 26939 -> 24000 (pitch +2 semitones): every tap torchaudio's dense fp32 kernel holds above 1e-20
    # is in the window, with the same value to an ulp
    ta = pytest.importorskip("torchaudio")
    orig, new = 11, 10                                                                 # 26400 -> 24000: small dense kernel
    taps = _table(lib, 6, 26400, new * 16).reshape(new, 16)
    kt, width = ta.functional.functional._get_sinc_resample_kernel(26400, 24000, 2400, dtype=torch.float32)
    dense = kt.numpy()[:, 0, :]
    assert width == 7
    for p in range(new):
        lo = max(0, int(math.floor(p * orig / new + width - 6 * orig / (min(orig, new) * 0.99))))
        row = np.zeros(dense.shape[1] + 16, np.float32)
        row[:dense.shape[1]] = dense[p]
        assert float(np.abs(taps[p] - row[lo:lo + 16]).max()) <= 1.2e-7
        outside = np.delete(dense[p], np.arange(lo, min(lo + 16, dense.shape[1])))
        assert float(np.abs(outside).max(initial=0.0)) < 1e-20
