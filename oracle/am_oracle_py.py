"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see am_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libam_oracle.so")
# AM_ORACLE_NATIVE=1 (bench.py's CPU-timing legs): a -march=native build made on THIS machine, kept
# apart from the portable one and rebuilt when the CPU model differs from the one it was built on.
_NATIVE = os.environ.get("AM_ORACLE_NATIVE") == "1"
if _NATIVE:
    _LIB_PATH = os.path.join(_HERE, "_build", "native", "libam_oracle.so")


def _cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class U8Slice(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("off", C.c_int64), ("len", C.c_int64)]


class Match(C.Structure):
    _fields_ = [("pos", C.c_int64), ("value", C.c_int64)]


MATCH_DTYPE = np.dtype([("pos", "<i8"), ("value", "<i8")])

_lib = None


def build_lib(force: bool = False) -> str:
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "am_oracle.c"))
    if _NATIVE:
        stamp = _LIB_PATH + ".cpu"
        if force or stale or not os.path.exists(stamp) or open(stamp).read() != _cpu_model():
            subprocess.check_call(["make", "-C", _HERE, "native"], stdout=subprocess.DEVNULL)
            with open(stamp, "w") as f:
                f.write(_cpu_model())
    elif force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if _NATIVE or not os.path.exists(_LIB_PATH):
            build_lib()
        L = C.CDLL(_LIB_PATH)
        L.amo_build.argtypes = [C.POINTER(U8Slice), C.c_size_t, C.POINTER(C.c_void_p)]
        L.amo_build.restype = C.c_int
        L.amo_free.argtypes = [C.c_void_p]
        L.amo_num_states.argtypes = [C.c_void_p]; L.amo_num_states.restype = C.c_int64
        L.amo_num_transitions.argtypes = [C.c_void_p]; L.amo_num_transitions.restype = C.c_int64
        L.amo_transitions.argtypes = [C.c_void_p]; L.amo_transitions.restype = C.POINTER(C.c_uint64)
        L.amo_offsets.argtypes = [C.c_void_p]; L.amo_offsets.restype = C.POINTER(C.c_uint32)
        L.amo_root_ascii.argtypes = [C.c_void_p]; L.amo_root_ascii.restype = C.POINTER(C.c_uint64)
        L.amo_lower_code_point.argtypes = [C.c_void_p, C.c_uint32]; L.amo_lower_code_point.restype = C.c_uint32
        L.amo_lower_utf8.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]; L.amo_lower_utf8.restype = C.c_int64
        L.amo_length_code_points.argtypes = [C.c_char_p, C.c_int64]; L.amo_length_code_points.restype = C.c_int64
        L.amo_skip_code_points_backwards.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.amo_skip_code_points_backwards.restype = C.c_int64
        L.amo_count.argtypes = [C.c_void_p, C.c_int, C.c_void_p, U8Slice]; L.amo_count.restype = C.c_uint64
        L.amo_count_parallel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, U8Slice, C.c_int]; L.amo_count_parallel.restype = C.c_uint64
        L.amo_contains_any.argtypes = [C.c_void_p, C.c_int, C.c_void_p, U8Slice]; L.amo_contains_any.restype = C.c_int
        L.amo_contains_all.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, U8Slice]; L.amo_contains_all.restype = C.c_int
        L.amo_find_all.argtypes = [C.c_void_p, C.c_int, C.c_void_p, U8Slice, C.c_void_p, C.c_int64]; L.amo_find_all.restype = C.c_int64
        L.amo_find_all_parallel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, U8Slice, C.c_void_p, C.c_int64, C.c_int]; L.amo_find_all_parallel.restype = C.c_int64
        L.amo_replacer_build.argtypes = [C.POINTER(U8Slice), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(U8Slice), C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.amo_replacer_build.restype = C.c_int
        L.amo_replacer_free.argtypes = [C.c_void_p]
        L.amo_replacer_run.argtypes = [C.c_void_p, C.c_void_p, U8Slice, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        L.amo_replacer_run.restype = C.c_int
        L.amo_buf_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


CASE_SENSITIVE, IGNORE_CASE = 0, 1


def _as_bytes(x) -> bytes:
    if isinstance(x, str):
        return x.encode("utf-8")
    return bytes(x)


class _Slices:
    """Keeps the python buffers alive while C sees (ptr, off, len) triples."""

    def __init__(self, items):
        self.bufs = [_as_bytes(x) for x in items]
        self.keep = [C.create_string_buffer(b, len(b)) if len(b) else C.create_string_buffer(1) for b in self.bufs]
        self.arr = (U8Slice * max(1, len(self.bufs)))()
        for i, (b, k) in enumerate(zip(self.bufs, self.keep)):
            self.arr[i] = U8Slice(C.cast(k, C.c_void_p).value, 0, len(b))


def _hay_slice(hay):
    """hay: bytes/str, a numpy uint8 array, or a (buffer, off, len) triple (a Text slice)."""
    off = None
    if isinstance(hay, tuple):
        hay, off, ln = hay
    if isinstance(hay, np.ndarray):
        arr = np.ascontiguousarray(hay, dtype=np.uint8)
    else:
        arr = np.frombuffer(_as_bytes(hay), dtype=np.uint8)
    if off is None:
        off, ln = 0, arr.size
    return arr, U8Slice(arr.ctypes.data if arr.size else 0, off, ln)


def lower_table_dense(pairs=None):
    """Dense 0x110000-entry table from (cp, lower cp) pairs; the table is DATA."""
    t = np.arange(0x110000, dtype=np.uint32)
    if pairs is not None:
        p = np.asarray(pairs, dtype=np.uint32).reshape(-1, 2)
        t[p[:, 0]] = p[:, 1]
    return t


class Machine:
    """`AcMachine` with needle indices as values (Automaton.hs:108-123, build :176-200)."""

    def __init__(self, needles):
        self._needles = _Slices(needles)
        self.n = len(self._needles.bufs)
        h = C.c_void_p()
        rc = lib().amo_build(self._needles.arr, self.n, C.byref(h))
        if rc != 0:
            raise MemoryError("amo_build failed")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib().amo_free(self.h)
            self.h = None

    @property
    def num_states(self):
        return lib().amo_num_states(self.h)

    def packed(self):
        ns, nt = self.num_states, lib().amo_num_transitions(self.h)
        tr = np.ctypeslib.as_array(lib().amo_transitions(self.h), (nt,)).copy()
        of = np.ctypeslib.as_array(lib().amo_offsets(self.h), (ns + 1,)).copy()
        ra = np.ctypeslib.as_array(lib().amo_root_ascii(self.h), (128,)).copy()
        return tr, of, ra

    @staticmethod
    def _lower_ptr(lower):
        return lower.ctypes.data if lower is not None else None

    def count(self, hay, cs=CASE_SENSITIVE, lower=None, threads=1):
        arr, s = _hay_slice(hay)
        if threads > 1:
            return int(lib().amo_count_parallel(self.h, cs, self._lower_ptr(lower), s, threads))
        return int(lib().amo_count(self.h, cs, self._lower_ptr(lower), s))

    def contains_any(self, hay, cs=CASE_SENSITIVE, lower=None):
        arr, s = _hay_slice(hay)
        return bool(lib().amo_contains_any(self.h, cs, self._lower_ptr(lower), s))

    def contains_all(self, hay, cs=CASE_SENSITIVE, lower=None):
        arr, s = _hay_slice(hay)
        return bool(lib().amo_contains_all(self.h, self.n, cs, self._lower_ptr(lower), s))

    def find_all(self, hay, cs=CASE_SENSITIVE, lower=None, threads=1, cap=1024):
        """All matches in the reference's callback order, as a structured array (pos, value)."""
        arr, s = _hay_slice(hay)
        while True:
            out = np.empty(cap, dtype=MATCH_DTYPE)
            if threads > 1:
                n = lib().amo_find_all_parallel(self.h, cs, self._lower_ptr(lower), s, out.ctypes.data, cap, threads)
            else:
                n = lib().amo_find_all(self.h, cs, self._lower_ptr(lower), s, out.ctypes.data, cap)
            if n <= cap:
                return out[:n]
            cap = int(n)


def lower_utf8(data, lower=None) -> bytes:
    b = _as_bytes(data)
    out = C.create_string_buffer(4 * len(b) + 4)
    n = lib().amo_lower_utf8(lower.ctypes.data if lower is not None else None, b, len(b), out, len(out))
    assert n >= 0
    return out.raw[:n]


def length_code_points(data) -> int:
    b = _as_bytes(data)
    return int(lib().amo_length_code_points(b, len(b)))


def skip_code_points_backwards(data, index, n, off=0, length=None):
    b = _as_bytes(data)
    if length is None:
        length = len(b) - off
    return int(lib().amo_skip_code_points_backwards(b, off, length, index, n))


class Replacer:
    """`Replacer.build` / `run` / `runWithLimit` (Replacer.hs:97-116, :200-274)."""

    def __init__(self, pairs, cs=CASE_SENSITIVE, lower=None):
        self.cs, self.lower = cs, lower
        needles = [_as_bytes(n) for n, _ in pairs]
        repls = [_as_bytes(r) for _, r in pairs]
        lb = (C.c_int64 * max(1, len(pairs)))(*[len(n) for n in needles])
        lc = (C.c_int64 * max(1, len(pairs)))(*[length_code_points(n) for n in needles])
        built = [lower_utf8(n, lower) if cs == IGNORE_CASE else n for n in needles]  # Replacer.hs:105-107
        self._n, self._r = _Slices(built), _Slices(repls)
        h = C.c_void_p()
        rc = lib().amo_replacer_build(self._n.arr, lb, lc, self._r.arr, len(pairs), cs, C.byref(h))
        if rc != 0:
            raise MemoryError("amo_replacer_build failed")
        self.h = h
        self.passes = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().amo_replacer_free(self.h)
            self.h = None

    def run_with_limit(self, hay, max_len=-1):
        arr, s = _hay_slice(hay)
        out, out_len, exceeded, passes = C.c_void_p(), C.c_int64(), C.c_int(), C.c_int64()
        rc = lib().amo_replacer_run(self.h, self.lower.ctypes.data if self.lower is not None else None, s, max_len,
                                    C.byref(out), C.byref(out_len), C.byref(exceeded), C.byref(passes))
        if rc != 0:
            raise RuntimeError("amo_replacer_run: error %d (the reference would call `error`)" % rc)
        self.passes = passes.value
        if exceeded.value:
            return None
        res = C.string_at(out.value, out_len.value) if out_len.value else b""
        lib().amo_buf_free(out)
        return res

    def run(self, hay):
        return self.run_with_limit(hay, -1)
