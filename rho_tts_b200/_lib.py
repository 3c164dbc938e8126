"""ctypes binding of librho_b200.so (C ABI declared in include/rho_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is
visible when a handle is requested, this module raises RuntimeError (never ValueError:
the reference pipeline treats ValueError as a configuration error and stops retrying,
src/rho_tts/base_tts.py:786-787).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# RHO_B200_LIB: developer override used by tools/ab_fused.sh to A/B kernel variants
LIB_PATH = os.environ.get("RHO_B200_LIB") or os.path.join(_HERE, "librho_b200.so")

ABI_VERSION = 2

# flags (include/rho_b200.h)
F_ALL_SILENT = 1
F_FALLBACK = 2
F_TWO_D = 4
F_UNTOUCHED = 8
V_ONE_SEGMENT_ITEMS = 1
V_NO_FUSION = 2
V_COMPACT_PAD = 4
V_GATHER_FIRST = 8

EXPORTS = (
    "rho_b200_abi_version", "rho_b200_create", "rho_b200_destroy", "rho_b200_last_error",
    "rho_b200_host_table", "rho_b200_host_mel_stream", "rho_b200_workspace_bytes", "rho_b200_trim_scan", "rho_b200_join",
    "rho_b200_remove_dc", "rho_b200_apply_fades", "rho_b200_sound_decay", "rho_b200_sound_decay_batch", "rho_b200_resample3to2",
    "rho_b200_resample_out_len", "rho_b200_resample", "rho_b200_pitch_workspace_bytes", "rho_b200_pitch_shift", "rho_b200_mfcc_workspace_bytes", "rho_b200_mfcc_stats", "rho_b200_spk_slices", "rho_b200_normalize_volume", "rho_b200_spk_mel", "rho_b200_spk_pool", "rho_b200_pcm16", "rho_b200_logmel", "rho_b200_mel_project", "rho_b200_stft_power_tc", "rho_b200_qwen_workspace_bytes", "rho_b200_qwen_postprocess",
    "rho_b200_cosine", "rho_b200_validate", "rho_b200_compact_frames", "rho_b200_exchange_create", "rho_b200_exchange_connect",
    "rho_b200_exchange_wait", "rho_b200_exchange_epoch", "rho_b200_exchange_read", "rho_b200_exchange_destroy",
    "rho_b200_validate_host", "rho_b200_validate_host_ragged", "rho_b200_host_fill_threads",
    "rho_b200_build_flags", "rho_b200_launch_count", "rho_b200_profile_begin", "rho_b200_profile_end", "rho_b200_kernel_name",
)


class RhoParams(ctypes.Structure):
    """rho_params: the BaseTTS attributes read at call time (base_tts.py:72-81)."""
    _fields_ = [
        ("sr", c_int32), ("trim_enabled", c_int32), ("silence_db", c_double), ("fade_sec", c_double),
        ("xfade_sec", c_double), ("pause_sec", c_double), ("decay_thr", c_double),
    ]


class RhoSegInfo(ctypes.Structure):
    _fields_ = [("start", c_int32), ("end", c_int32), ("dc", c_float), ("flags", c_uint32)]


class RhoRecord(ctypes.Structure):
    _fields_ = [
        ("start", c_int32), ("end", c_int32), ("out_len", c_int32), ("flags", c_uint32),
        ("dc", c_float), ("first_rms", c_float), ("last_rms", c_float), ("cosine", c_float),
        ("decay_ratio", c_double), ("ok", c_int32), ("n_segments", c_int32),
    ]


assert ctypes.sizeof(RhoRecord) == 48 and ctypes.sizeof(RhoSegInfo) == 16 and ctypes.sizeof(RhoParams) == 48

_lib = None
_lib_lock = threading.Lock()


def load():
    """Load librho_b200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C rho_tts_b200/csrc`. rho_tts_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        vp, i64, i32 = c_void_p, c_int64, c_int
        P = POINTER(RhoParams)
        sig = {
            "rho_b200_abi_version": (c_int, []),
            "rho_b200_create": (c_int, [POINTER(c_void_p), c_int]),
            "rho_b200_destroy": (c_int, [vp]),
            "rho_b200_last_error": (c_char_p, []),
            "rho_b200_host_table": (c_int, [c_int, c_int, vp, c_size_t]),
            "rho_b200_host_mel_stream": (c_int, [c_int, vp, c_size_t, vp, vp, POINTER(c_int)]),
            "rho_b200_workspace_bytes": (c_size_t, [c_int, c_int, i64]),
            "rho_b200_trim_scan": (c_int, [vp, vp, vp, vp, vp, i32, i64, P, vp, vp, c_size_t, vp]),
            "rho_b200_join": (c_int, [vp, vp, vp, vp, i32, i64, vp, i32, i64, P, vp, vp, vp, vp, vp, c_size_t, vp]),
            "rho_b200_remove_dc": (c_int, [vp, vp, i64, vp, vp, c_size_t, vp]),
            "rho_b200_apply_fades": (c_int, [vp, vp, i64, c_int, c_int, P, vp]),
            "rho_b200_sound_decay": (c_int, [vp, vp, i64, P, vp, vp, c_size_t, vp]),
            "rho_b200_sound_decay_batch": (c_int, [vp, vp, vp, vp, c_int, i32, i64, P, vp, vp, c_size_t, vp]),
            "rho_b200_resample3to2": (c_int, [vp, vp, vp, vp, c_int, i32, i64, vp, vp, vp, vp]),
            "rho_b200_resample_out_len": (c_int64, [i64, c_int, c_int]),
            "rho_b200_resample": (c_int, [vp, vp, vp, vp, c_int, i32, i64, c_int, c_int, vp, vp, vp, vp]),
            "rho_b200_pitch_workspace_bytes": (c_size_t, [c_int, i64, c_double]),
            "rho_b200_pitch_shift": (c_int, [vp, vp, vp, vp, c_int, i32, i64, i64, c_int, c_double, c_int, vp, vp, vp,
                                             c_size_t, vp]),
            "rho_b200_pcm16": (c_int, [vp, vp, vp, vp, c_int, i32, i64, vp, vp, vp]),
            "rho_b200_mfcc_workspace_bytes": (c_size_t, [c_int, i64]),
            "rho_b200_mfcc_stats": (c_int, [vp, vp, vp, vp, c_int, i32, i64, vp, vp, c_size_t, vp]),
            "rho_b200_spk_slices": (c_int, [i64, c_int, c_double, POINTER(c_int64)]),
            "rho_b200_normalize_volume": (c_int, [vp, vp, vp, vp, c_int, i32, i64, c_float, c_int, vp, vp, vp, vp,
                                                  c_size_t, vp]),
            "rho_b200_spk_mel": (c_int, [vp, vp, vp, vp, c_int, i32, i64, c_int, c_double, c_uint32, vp, vp, vp, vp, vp,
                                         vp]),
            "rho_b200_spk_pool": (c_int, [vp, vp, vp, i32, c_int, vp, vp]),
            "rho_b200_logmel": (c_int, [vp, vp, vp, vp, i32, i64, c_int, c_int, vp, i64, vp, vp, c_size_t, vp]),
            "rho_b200_mel_project": (c_int, [vp, vp, i64, i64, c_int, vp, i64, i64, i64, vp]),
            "rho_b200_stft_power_tc": (c_int, [vp, vp, vp, vp, c_int, vp, c_int, vp, i64, vp]),
            "rho_b200_qwen_workspace_bytes": (c_size_t, [c_int, i64, c_int]),
            "rho_b200_qwen_postprocess": (c_int, [vp, vp, vp, vp, c_int, c_int, i64, c_int, vp, vp, vp, c_size_t, vp]),
            "rho_b200_cosine": (c_int, [vp, vp, vp, i32, c_int, vp, c_int, vp]),
            "rho_b200_validate": (c_int, [vp, vp, vp, vp, i32, i64, vp, i32, i64, P, vp, vp, c_int, c_int, vp, i64, vp,
                                          vp, vp, c_int, vp, vp, c_uint32, vp, c_size_t, vp]),
            "rho_b200_compact_frames": (c_int64, [i64, c_int]),
            "rho_b200_exchange_create": (c_int, [vp, c_int, c_int, i64, vp]),
            "rho_b200_exchange_connect": (c_int, [vp, vp]),
            "rho_b200_exchange_wait": (c_int, [vp, i64, vp]),
            "rho_b200_exchange_epoch": (c_int64, [vp]),
            "rho_b200_exchange_read": (c_int, [vp, i64, vp, POINTER(c_int), vp]),
            "rho_b200_exchange_destroy": (c_int, [vp]),
            "rho_b200_validate_host_ragged": (c_int, [vp, vp, vp, vp, c_int, vp, c_int, P, vp, vp, c_int, c_int, vp, i64,
                                                      vp, vp, vp, c_int, vp]),
            "rho_b200_validate_host": (c_int, [vp, vp, c_int, c_int32, P, vp, c_int, c_int, vp, vp, vp, c_int, vp]),
            "rho_b200_host_fill_threads": (c_int, [vp]),
            "rho_b200_build_flags": (c_int, []),
            "rho_b200_launch_count": (c_int64, [vp]),
            "rho_b200_profile_begin": (c_int, [vp]),
            "rho_b200_profile_end": (c_int, [vp, POINTER(c_double), POINTER(c_int64), c_int]),
            "rho_b200_kernel_name": (c_char_p, [c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.rho_b200_abi_version() != ABI_VERSION:
            raise RuntimeError(f"librho_b200 ABI {lib.rho_b200_abi_version()} != expected {ABI_VERSION}")
        _lib = lib
        return lib


def last_error() -> str:
    return (load().rho_b200_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"librho_b200 {what} failed ({rc}): {last_error()}")


class Handle:
    """One library handle per (process, device).  Creation fails loudly without a B200."""

    _cache: dict = {}
    _cache_lock = threading.Lock()

    def __init__(self, device: int = 0):
        self.lib = load()
        self.device = int(device)
        h = c_void_p()
        check(self.lib.rho_b200_create(ctypes.byref(h), self.device), "create")
        self.ptr = h

    @classmethod
    def get(cls, device: int = 0) -> "Handle":
        with cls._cache_lock:
            hd = cls._cache.get(device)
            if hd is None:
                hd = cls._cache[device] = Handle(device)
            return hd

    @property
    def launch_count(self) -> int:
        return int(self.lib.rho_b200_launch_count(self.ptr))

    def profile_begin(self) -> None:
        check(self.lib.rho_b200_profile_begin(self.ptr), "profile_begin")

    def profile_end(self) -> dict:
        """{kernel name: (total ms, launches)} for the kernels issued since profile_begin."""
        ms = (c_double * 32)()
        cnt = (c_int64 * 32)()
        n = self.lib.rho_b200_profile_end(self.ptr, ms, cnt, 32)
        if n < 0:
            check(n, "profile_end")
        return {self.lib.rho_b200_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}

    def close(self):
        if self.ptr:
            self.lib.rho_b200_destroy(self.ptr)
            self.ptr = None
