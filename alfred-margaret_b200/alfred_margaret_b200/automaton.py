"""Host mirror of `Data.Text.AhoCorasick.Automaton` (src/Data/Text/AhoCorasick/Automaton.hs).

Same surface -- `build`, `runText`, `runLower`, `runWithCase`, `Match`, `Next` (export list :32-44) --
but the state-transition loop (:442-534) runs in hand-written sm_100a CUDA kernels behind the C ABI.
The native side deals in needle indices; this layer maps index -> payload `v` and folds the
caller's function over the ordered match array, honouring `Done` (early exit, :398, :530-532).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, Sequence, Tuple

import numpy as np

from . import _ffi
from .case_sensitivity import CaseSensitivity
from .utf8 import LowerTableArg, Text, as_text, default_lower_table

MATCH_DTYPE = np.dtype([("end_pos", "<u8"), ("needle_id", "<u4"), ("reserved", "<u4")])


@dataclass(frozen=True)
class Match:
    """`Match { matchPos, matchValue }` (:98-105): pos is the code unit index one past the match."""
    pos: int
    value: Any


@dataclass(frozen=True)
class Done:  # `Next a = Done !a | Step !a` (:398)
    acc: Any


@dataclass(frozen=True)
class Step:
    acc: Any


class _Needles:
    def __init__(self, needles: Sequence):
        self.bufs = [n.encode("utf-8") if isinstance(n, str) else (n.tobytes() if isinstance(n, Text) else bytes(n)) for n in needles]
        self.keep = [C.create_string_buffer(b, len(b)) if b else C.create_string_buffer(1) for b in self.bufs]
        self.arr = (_ffi.U8Slice * max(1, len(self.bufs)))()
        for i, (b, k) in enumerate(zip(self.bufs, self.keep)):
            self.arr[i] = _ffi.U8Slice(C.cast(k, C.c_void_p).value, 0, len(b))


class AcMachine:
    """`AcMachine v` (:108-123).  Immutable and case-agnostic like the reference's: `run_text` and `run_lower` run
    the SAME machine (:539-553); the device image of a case mode is built when that mode is first used.
    `case_sensitivity` is only the DEFAULT mode of the convenience methods below (and the image built eagerly)."""

    def __init__(self, needles_with_values: Sequence[Tuple[Any, Any]], case_sensitivity=CaseSensitivity.CaseSensitive,
                 lower_table: LowerTableArg | None = None, device: int = -1, force_kernel: int = 0):
        self.values = [v for _, v in needles_with_values]
        self._needles = _Needles([n for n, _ in needles_with_values])
        self.case_sensitivity = CaseSensitivity(case_sensitivity)
        self._lower = lower_table or default_lower_table()
        opts = _ffi.Options(device, force_kernel, (C.c_uint64 * 6)())
        h = C.c_void_p()
        _ffi.check(_ffi.lib().am_automaton_build(self._needles.arr, len(self.values), self._lower.ptr(), C.byref(opts), C.byref(h)))
        self.handle = h
        self._owner = True
        try:
            _ffi.check(_ffi.lib().am_automaton_prepare(h, int(self.case_sensitivity)))   # errors of the default mode surface here
        except Exception:
            self.__del__()
            raise

    def with_values(self, values):
        """`fmap` on AcMachine (Functor, :123): same device image, remapped payloads."""
        return self._share(values=list(values))

    def with_case(self, case_sensitivity):
        """The same machine with another default case mode (what `Searcher.setCaseSensitivity` needs: no rebuild)."""
        return self._share(case_sensitivity=CaseSensitivity(case_sensitivity))

    def _share(self, **changes):
        m = object.__new__(AcMachine)
        m.__dict__.update(self.__dict__)
        m.__dict__.update(changes)
        m._owner = False
        m._parent = self  # keeps the handle alive
        return m

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and getattr(self, "_owner", False):
            try:
                _ffi.lib().am_automaton_free(h)
            except Exception:
                pass
            self.handle = None

    def _cs(self, case) -> int:
        return int(self.case_sensitivity if case is None else CaseSensitivity(case))

    def info(self, case=None):
        ns, mx, halo, kind = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        _ffi.check(_ffi.lib().am_automaton_info(self.handle, self._cs(case), C.byref(ns), C.byref(mx), C.byref(halo), C.byref(kind)))
        return {"num_states": ns.value, "max_needle_bytes": mx.value, "halo_bytes": halo.value, "kernel_kind": kind.value}

    def host_filter_flags(self, text, align: int = 0, case=None) -> np.ndarray:
        """Introspection (works on a host image, device=-2): the fast path's q-gram filter evaluated on the host for every
        start position; bit 0 = shared-memory bitmap passes, bit 1 = second level passes (am_debug_host_filter)."""
        t = as_text(text)
        out = np.zeros(max(1, t.len), dtype=np.uint8)
        sl = t.slice()
        _ffi.check(_ffi.lib().am_debug_host_filter(self.handle, self._cs(case), C.byref(sl), align, out.ctypes.data))
        return out[: t.len]

    # ---- raw results (needle indices) -------------------------------------------------------------
    def find_all(self, text, case=None) -> np.ndarray:
        """All matches in the reference's callback order as a structured array (end_pos, needle_id)."""
        t = as_text(text)
        sl = t.slice()
        cap = 4096 + t.len // 256
        while True:
            out = np.empty(cap, dtype=MATCH_DTYPE)
            n = C.c_uint64()
            rc = _ffi.lib().am_find_all(self.handle, self._cs(case), C.byref(sl), out.ctypes.data, cap, C.byref(n))
            if rc == _ffi.AM_E_OVERFLOW:       # the call reports the capacity it needs: one retry
                cap = int(n.value)
                continue
            _ffi.check(rc)
            return out[: n.value]

    def count_matches(self, text, case=None) -> int:
        t = as_text(text)  # keep the buffer alive across the call
        sl = t.slice()
        n = C.c_uint64()
        _ffi.check(_ffi.lib().am_count_matches(self.handle, self._cs(case), C.byref(sl), C.byref(n)))
        return n.value

    def contains_any(self, text, case=None) -> bool:
        t = as_text(text)
        sl = t.slice()
        b = C.c_int()
        _ffi.check(_ffi.lib().am_contains_any(self.handle, self._cs(case), C.byref(sl), C.byref(b)))
        return bool(b.value)

    def contains_all(self, text, case=None) -> bool:
        t = as_text(text)
        sl = t.slice()
        b = C.c_int()
        _ffi.check(_ffi.lib().am_contains_all(self.handle, self._cs(case), C.byref(sl), C.byref(b)))
        return bool(b.value)

    # ---- device-resident variants (dev_ptr: CUDA device pointer as int) -----------------------------
    def _dev_text(self, dev_ptr, text_len, report_begin=0, pos_base=0):
        return _ffi.DevText(dev_ptr, text_len, report_begin, pos_base)

    def count_matches_dev(self, dev_ptr, text_len, report_begin=0, pos_base=0, stream=None, case=None) -> int:
        n = C.c_uint64()
        t = self._dev_text(dev_ptr, text_len, report_begin, pos_base)
        _ffi.check(_ffi.lib().am_count_matches_dev(self.handle, self._cs(case), C.byref(t), stream, C.byref(n)))
        return n.value

    def contains_any_dev(self, dev_ptr, text_len, report_begin=0, pos_base=0, stream=None, case=None) -> bool:
        b = C.c_int()
        t = self._dev_text(dev_ptr, text_len, report_begin, pos_base)
        _ffi.check(_ffi.lib().am_contains_any_dev(self.handle, self._cs(case), C.byref(t), stream, C.byref(b)))
        return bool(b.value)

    def find_all_dev(self, dev_ptr, text_len, out_dev_ptr, cap, report_begin=0, pos_base=0, stream=None, case=None) -> int:
        """Sorted am_match records into device memory; returns n_found (raises AM_E_OVERFLOW if > cap)."""
        n = C.c_uint64()
        t = self._dev_text(dev_ptr, text_len, report_begin, pos_base)
        rc = _ffi.lib().am_find_all_dev(self.handle, self._cs(case), C.byref(t), stream, out_dev_ptr, cap, C.byref(n))
        if rc == _ffi.AM_E_OVERFLOW:
            raise OverflowError(n.value)
        _ffi.check(rc)
        return n.value


def build(needles_with_values: Sequence[Tuple[Any, Any]], **kw) -> AcMachine:
    """`build :: [(Text, v)] -> AcMachine v` (:176).  One machine for both case modes."""
    return AcMachine(needles_with_values, **kw)


def run_with_case(case_sensitivity, seed, f: Callable[[Any, Match], Any], machine: AcMachine, text):
    """`runWithCase` (:443): left fold of `f` over all matches; `f` returns Step(acc) or Done(acc).

    Any machine runs in either mode, as in the reference (its own test helper builds once and picks the mode per
    call, tests/Data/Text/AhoCorasickSpec.hs:252-261); for IgnoreCase the caller has lower-cased the needles (:543-546).
    """
    acc = seed
    ms = machine.find_all(text, case=CaseSensitivity(case_sensitivity))
    values = machine.values
    for pos, nid in zip(ms["end_pos"].tolist(), ms["needle_id"].tolist()):
        nxt = f(acc, Match(pos, values[nid]))
        if isinstance(nxt, Done):
            return nxt.acc
        acc = nxt.acc
    return acc


def run_text(seed, f, machine: AcMachine, text):
    """`runText = runWithCase CaseSensitive` (:539-541)."""
    return run_with_case(CaseSensitivity.CaseSensitive, seed, f, machine, text)


def run_lower(seed, f, machine: AcMachine, text):
    """`runLower = runWithCase IgnoreCase` (:551-553); the caller lower-cases the needles."""
    return run_with_case(CaseSensitivity.IgnoreCase, seed, f, machine, text)
