"""Host mirror of the parts of `Data.Text.Utf8` the hot path needs (src/Data/Text/Utf8.hs).

Texts are UTF-8 byte strings; indices are code UNIT (byte) indices (`CodeUnitIndex`, :106-114).
A `Text` slice (array, off, len) is modelled by `Text`.
"""
from __future__ import annotations

import ctypes as C
import functools
import sys
import unicodedata

import numpy as np

from . import _ffi


class Text:
    """`Text u8data off len` (text-2.x internal representation; Automaton.hs:449)."""

    __slots__ = ("array", "off", "len", "_keep")

    def __init__(self, array, off=None, length=None):
        if isinstance(array, str):
            array = array.encode("utf-8")
        if isinstance(array, Text):
            self.array, self.off, self.len = array.array, array.off, array.len
            return
        self.array = np.frombuffer(array, dtype=np.uint8) if not isinstance(array, np.ndarray) else np.ascontiguousarray(array, dtype=np.uint8)
        self.off = 0 if off is None else int(off)
        self.len = (self.array.size - self.off) if length is None else int(length)

    def slice(self) -> _ffi.U8Slice:
        return _ffi.U8Slice(self.array.ctypes.data if self.array.size else 0, self.off, self.len)

    def tobytes(self) -> bytes:
        return self.array[self.off:self.off + self.len].tobytes()

    def __len__(self):  # lengthUtf8 (:127-128)
        return self.len


def as_text(x) -> Text:
    return x if isinstance(x, Text) else Text(x)


@functools.lru_cache(maxsize=1)
def host_lower_pairs():
    """The host's `Data.Char.toLower` above ASCII as (from, to) pairs -- the table the ABI takes as DATA.

    A Haskell host would enumerate its own `Char.toLower`; this Python host derives the Unicode
    *simple* lower-case mapping from `str.lower()`, which is single-code-point for every scalar
    except U+0130 (simple mapping U+0069), cf. SURVEY.md section 8c.
    """
    pairs = []
    for cp in range(128, sys.maxunicode + 1):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        low = chr(cp).lower()
        to = 0x69 if cp == 0x130 else (ord(low) if len(low) == 1 else cp)
        if to != cp:
            pairs.append((cp, to))
    return np.asarray(pairs, dtype=np.uint32).reshape(-1, 2)


UNICODE_VERSION = unicodedata.unidata_version


class LowerTableArg:
    """Keeps the ctypes view of a lower table alive."""

    def __init__(self, pairs=None):
        self.pairs = host_lower_pairs() if pairs is None else np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        self.struct = _ffi.LowerTable(C.cast(self.pairs.ctypes.data, C.POINTER(_ffi.LowerPair)), self.pairs.shape[0])

    def ptr(self):
        return C.byref(self.struct)


@functools.lru_cache(maxsize=1)
def default_lower_table() -> LowerTableArg:
    return LowerTableArg()


def lower_code_point(cp: int) -> int:
    """`lowerCodePoint` (:145-151) for this host's table."""
    if cp < 128:
        return cp + 32 if 65 <= cp <= 90 else cp
    low = chr(cp).lower()
    return 0x69 if cp == 0x130 else (ord(low) if len(low) == 1 else cp)


def lower_utf8(text, table: LowerTableArg | None = None) -> bytes:
    """`lowerUtf8` (:138-140), computed by the library with the table passed as data."""
    t = as_text(text)
    table = table or default_lower_table()
    cap = 4 * t.len + 8
    out = (C.c_uint8 * cap)()
    n = C.c_uint64()
    sl = t.slice()
    _ffi.check(_ffi.lib().am_lower_utf8(table.ptr(), C.byref(sl), out, cap, C.byref(n)))
    return bytes(out[: n.value])


def skip_code_points_backwards(text, index: int, n: int) -> int:
    """`skipCodePointsBackwards` (:256-276); raises where the reference calls `error`."""
    t = as_text(text)
    out = C.c_int64()
    sl = t.slice()
    rc = _ffi.lib().am_skip_code_points_backwards(C.byref(sl), index, n, C.byref(out))
    if rc == _ffi.AM_E_BADARG:
        raise ValueError("Invalid use of skipCodePointsBackwards")
    _ffi.check(rc)
    return out.value
