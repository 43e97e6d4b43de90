"""ctypes binding of libam_b200.so (C ABI: include/am_b200.h).

This is the stand-in for the Haskell `foreign import ccall` stubs shown in INTEGRATION.md: GHC
is not available in this image, so the host-side mirror of the reference API is Python and
binds the same symbols.  There is no CPU fallback: if the library is missing, import fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AM_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libam_b200.so")  # AM_LIB: kernel-variant builds (development)

AM_OK, AM_E_BADARG, AM_E_OOM, AM_E_CUDA, AM_E_OVERFLOW, AM_E_NODEVICE, AM_E_UNSUPPORTED, AM_E_INTERNAL = range(8)
_NAMES = ["AM_OK", "AM_E_BADARG", "AM_E_OOM", "AM_E_CUDA", "AM_E_OVERFLOW", "AM_E_NODEVICE", "AM_E_UNSUPPORTED", "AM_E_INTERNAL"]


class AmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (_NAMES[code] if 0 <= code < len(_NAMES) else code, msg))
        self.code = code


class NoDeviceError(AmError):
    pass


class U8Slice(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("off", C.c_int64), ("len", C.c_int64)]


class Match(C.Structure):
    _fields_ = [("end_pos", C.c_uint64), ("needle_id", C.c_uint32), ("reserved", C.c_uint32)]


class LowerPair(C.Structure):
    _fields_ = [("from_cp", C.c_uint32), ("to_cp", C.c_uint32)]


class LowerTable(C.Structure):
    _fields_ = [("pairs", C.POINTER(LowerPair)), ("n", C.c_size_t)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("force_kernel", C.c_int32), ("reserved", C.c_uint64 * 6)]


class DevText(C.Structure):
    _fields_ = [("dev_text", C.c_void_p), ("text_len", C.c_uint64), ("report_begin", C.c_uint64), ("pos_base", C.c_uint64)]


class ShardResult(C.Structure):
    _fields_ = [("n_local", C.c_uint64), ("global_offset", C.c_uint64), ("total", C.c_uint64)]


COMM_ID_BYTES = 128
_P = C.POINTER
# Every symbol include/am_b200.h declares: (restype, argtypes).  Structs travel by pointer (ABI version 2).
SYMBOLS = {
    "am_last_error": (C.c_char_p, []),
    "am_last_error_copy": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "am_abi_version": (C.c_int, []),
    "am_device_count": (C.c_int, []),
    "am_automaton_build": (C.c_int, [_P(U8Slice), C.c_size_t, _P(LowerTable), _P(Options), _P(C.c_void_p)]),
    "am_automaton_free": (None, [C.c_void_p]),
    "am_automaton_prepare": (C.c_int, [C.c_void_p, C.c_int]),
    "am_debug_host_filter": (C.c_int, [C.c_void_p, C.c_int, _P(U8Slice), C.c_uint32, C.c_void_p]),
    "am_automaton_info": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint64), _P(C.c_int)]),
    "am_contains_any": (C.c_int, [C.c_void_p, C.c_int, _P(U8Slice), _P(C.c_int)]),
    "am_count_matches": (C.c_int, [C.c_void_p, C.c_int, _P(U8Slice), _P(C.c_uint64)]),
    "am_find_all": (C.c_int, [C.c_void_p, C.c_int, _P(U8Slice), C.c_void_p, C.c_size_t, _P(C.c_uint64)]),
    "am_contains_all": (C.c_int, [C.c_void_p, C.c_int, _P(U8Slice), _P(C.c_int)]),
    "am_count_matches_dev": (C.c_int, [C.c_void_p, C.c_int, _P(DevText), C.c_void_p, _P(C.c_uint64)]),
    "am_contains_any_dev": (C.c_int, [C.c_void_p, C.c_int, _P(DevText), C.c_void_p, _P(C.c_int)]),
    "am_find_all_dev": (C.c_int, [C.c_void_p, C.c_int, _P(DevText), C.c_void_p, C.c_void_p, C.c_size_t, _P(C.c_uint64)]),
    "am_shard_plan": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint64)]),
    "am_comm_unique_id": (C.c_int, [C.c_void_p]),
    "am_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, _P(C.c_void_p)]),
    "am_comm_free": (None, [C.c_void_p]),
    "am_count_sharded": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _P(DevText), C.c_void_p, _P(ShardResult)]),
    "am_find_all_sharded": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _P(DevText), C.c_void_p, C.c_void_p, C.c_size_t, _P(ShardResult)]),
    "am_contains_any_sharded": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _P(DevText), C.c_void_p, _P(C.c_int)]),
    "am_shard_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "am_comm_allreduce_u64": (C.c_int, [C.c_void_p, _P(C.c_uint64), C.c_int, C.c_void_p]),
    "am_replacer_build": (C.c_int, [_P(U8Slice), _P(U8Slice), C.c_size_t, C.c_int, _P(LowerTable), _P(Options), _P(C.c_void_p)]),
    "am_replacer_build_stored": (C.c_int, [_P(U8Slice), _P(C.c_uint32), _P(C.c_uint32), _P(U8Slice), C.c_size_t, C.c_int, _P(LowerTable), _P(Options), _P(C.c_void_p)]),
    "am_replacer_free": (None, [C.c_void_p]),
    "am_replacer_run": (C.c_int, [C.c_void_p, C.c_int, _P(U8Slice), C.c_uint64, _P(C.c_void_p), _P(C.c_uint64), _P(C.c_int)]),
    "am_replacer_run_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, _P(C.c_void_p), _P(C.c_uint64), _P(C.c_int)]),
    "am_replacer_last_passes": (C.c_uint64, []),
    "am_replacer_last_rescans": (C.c_uint64, []),
    "am_replacer_last_profile": (C.c_int, [_P(C.c_float), _P(C.c_uint64)]),
    "am_free": (None, [C.c_void_p]),
    "am_dev_free": (None, [C.c_void_p]),
    "am_lower_utf8": (C.c_int, [_P(LowerTable), _P(U8Slice), C.c_void_p, C.c_size_t, _P(C.c_uint64)]),
    "am_skip_code_points_backwards": (C.c_int, [_P(U8Slice), C.c_int64, C.c_int64, _P(C.c_int64)]),
    "am_profile_enable": (C.c_int, [C.c_int]),
    "am_profile_last_scan_ms": (C.c_int, [_P(C.c_float)]),
    "am_profile_kernel_launches": (C.c_uint64, []),
    "am_synth_fill_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint32, C.c_void_p]),
    "am_synth_plant_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, _P(U8Slice), C.c_size_t, C.c_uint32, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libam_b200.so is missing at %s -- build it with `make -C alfred-margaret_b200` "
                "(or __graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc == AM_OK:
        return
    msg = (lib().am_last_error() or b"").decode("utf-8", "replace")
    if rc == AM_E_NODEVICE:
        raise NoDeviceError(rc, msg)
    raise AmError(rc, msg)
