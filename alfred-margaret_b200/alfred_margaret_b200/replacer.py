"""Host mirror of `Data.Text.AhoCorasick.Replacer` (src/Data/Text/AhoCorasick/Replacer.hs)."""
from __future__ import annotations

import ctypes as C
from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .automaton import _Needles
from .case_sensitivity import CaseSensitivity
from .utf8 import LowerTableArg, as_text, default_lower_table, lower_utf8

MAX_BOUND = (1 << 64) - 1


class Stored(NamedTuple):
    """One `(needle, Payload)` pair as the replacer's searcher holds it (:59-75): the needle lowered iff the replacer was
    BUILT with IgnoreCase (:105-107), the lengths those of the original needle (:111-113).  Priority = -index."""
    needle: bytes
    length_bytes: int
    length_code_points: int
    replacement: bytes


class _Handle:
    """Owns one am_replacer; shared by the replacers that differ only in their case flag."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        if self.h:
            try:
                _ffi.lib().am_replacer_free(self.h)
            except Exception:
                pass
            self.h = None


class Replacer:
    """`Replacer` (:78-80): a `Searcher Payload`.  Pair i has priority -i (:101-111)."""

    def __init__(self, case_sensitivity, stored: Sequence[Stored], handle: _Handle, lower_table: LowerTableArg, device: int):
        self._case = CaseSensitivity(case_sensitivity)
        self._stored = list(stored)
        self._h = handle
        self._lower = lower_table
        self._device = device
        self.last_passes = 0
        self.last_rescans = 0   # passes that scanned the whole text (1 when the match list is carried between passes)

    @property
    def handle(self):
        return self._h.h

    def __eq__(self, other):                       # derived Eq: the searcher's needles (with payloads) and case flag
        return isinstance(other, Replacer) and (self._case, self._stored) == (other._case, other._stored)

    def __hash__(self):
        return hash((self._case, tuple(self._stored)))


def _b(x) -> bytes:
    return x.encode("utf-8") if isinstance(x, str) else bytes(x)


def _from_stored(case_, stored: List[Stored], lower_table, device) -> Replacer:
    n = _Needles([s.needle for s in stored])
    r = _Needles([s.replacement for s in stored])
    lb = np.asarray([s.length_bytes for s in stored] or [0], dtype=np.uint32)
    lc = np.asarray([s.length_code_points for s in stored] or [0], dtype=np.uint32)
    opts = _ffi.Options(device, 0, (C.c_uint64 * 6)())
    h = C.c_void_p()
    _ffi.check(_ffi.lib().am_replacer_build_stored(n.arr, lb.ctypes.data_as(C.POINTER(C.c_uint32)), lc.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                   r.arr, len(stored), int(case_), lower_table.ptr(), C.byref(opts), C.byref(h)))
    return Replacer(case_, stored, _Handle(h), lower_table, device)


def build(case_sensitivity, replaces: Sequence[Tuple], lower_table: LowerTableArg | None = None, device: int = -1) -> Replacer:
    """`build :: CaseSensitivity -> [(Needle, Replacement)] -> Replacer` (:97-116): am_replacer_build lowers the needles
    of an IgnoreCase replacer and records the original lengths."""
    case_ = CaseSensitivity(case_sensitivity)
    lower_table = lower_table or default_lower_table()
    pairs = [(_b(n), _b(r)) for n, r in replaces]
    n = _Needles([p[0] for p in pairs])
    r = _Needles([p[1] for p in pairs])
    opts = _ffi.Options(device, 0, (C.c_uint64 * 6)())
    h = C.c_void_p()
    _ffi.check(_ffi.lib().am_replacer_build(n.arr, r.arr, len(pairs), int(case_), lower_table.ptr(), C.byref(opts), C.byref(h)))
    stored = [Stored(lower_utf8(nd, lower_table) if case_ == CaseSensitivity.IgnoreCase else nd, len(nd), len(nd.decode("utf-8")), rp) for nd, rp in pairs]
    return Replacer(case_, stored, _Handle(h), lower_table, device)


def compose(r1: Replacer, r2: Replacer) -> Optional[Replacer]:
    """`compose` (:120-133): r2 after r1; None if the case sensitivities differ.  The STORED needles are concatenated and
    renumbered -- nothing is lowered again."""
    if r1._case != r2._case:
        return None
    return _from_stored(r1._case, r1._stored + r2._stored, r1._lower, r1._device)


def map_replacement(f, r: Replacer) -> Replacer:
    """`mapReplacement` (:136-141): "It doesn't modify the needles"."""
    return _from_stored(r._case, [s._replace(replacement=_b(f(s.replacement))) for s in r._stored], r._lower, r._device)


def replacer_case_sensitivity(r: Replacer) -> CaseSensitivity:
    return r._case


def set_case_sensitivity(case_, r: Replacer) -> Replacer:
    """`setCaseSensitivity` (:151-153): "Does not change the capitilization of the needles" -- only the flag changes; the
    stored needles, the payload lengths and the device handle are shared."""
    return Replacer(case_, r._stored, r._h, r._lower, r._device)


def run_with_limit(r: Replacer, max_length: int, text) -> Optional[bytes]:
    """`runWithLimit :: Replacer -> CodeUnitIndex -> Text -> Maybe Text` (:203-242); None is `Nothing`."""
    t = as_text(text)
    sl = t.slice()
    out, out_len, exceeded = C.c_void_p(), C.c_uint64(), C.c_int()
    _ffi.check(_ffi.lib().am_replacer_run(r.handle, int(r._case), C.byref(sl), max_length, C.byref(out), C.byref(out_len), C.byref(exceeded)))
    r.last_passes = _ffi.lib().am_replacer_last_passes()
    r.last_rescans = _ffi.lib().am_replacer_last_rescans()
    if exceeded.value:
        return None
    n = out_len.value
    res = bytes(memoryview((C.c_char * n).from_address(out.value))) if n else b""   # (string_at takes a C int: no > 2 GiB)
    _ffi.lib().am_free(out)
    return res


def run(r: Replacer, text) -> bytes:
    """`run = fromJust . runWithLimit replacer maxBound` (:200-201)."""
    return run_with_limit(r, MAX_BOUND, text)


def run_dev(r: Replacer, dev_ptr: int, text_len: int, max_length: int = MAX_BOUND, stream=None):
    """Device-resident form (am_replacer_run_dev): returns (device pointer of the result, its length) or None; the
    caller releases the buffer with `free_dev`."""
    out, out_len, exceeded = C.c_void_p(), C.c_uint64(), C.c_int()
    _ffi.check(_ffi.lib().am_replacer_run_dev(r.handle, int(r._case), dev_ptr, text_len, max_length, stream, C.byref(out), C.byref(out_len), C.byref(exceeded)))
    r.last_passes = _ffi.lib().am_replacer_last_passes()
    r.last_rescans = _ffi.lib().am_replacer_last_rescans()
    if exceeded.value:
        return None
    return out.value, out_len.value


def free_dev(dev_ptr) -> None:
    _ffi.lib().am_dev_free(dev_ptr)


def last_profile():
    """(device ms, bytes read + written) of the last run on this thread (needs am_profile_enable(1))."""
    ms, b = C.c_float(), C.c_uint64()
    _ffi.check(_ffi.lib().am_replacer_last_profile(C.byref(ms), C.byref(b)))
    return ms.value, b.value


# ---- aeson-compatible JSON (generic instances of `Replacer` and `Payload`, :56-83) -----------------------------------
def to_json(r: Replacer) -> dict:
    """`Replacer { replacerSearcher :: Searcher Payload }` with the generic encodings: the searcher object of
    Searcher.hs:68-73 whose values are `Payload` records, exactly as stored."""
    needles = [[s.needle.decode("utf-8"), {"needlePriority": -i, "needleLengthBytes": s.length_bytes,
                                          "needleLengthCodePoints": s.length_code_points, "needleReplacement": s.replacement.decode("utf-8")}]
               for i, s in enumerate(r._stored)]
    return {"replacerSearcher": {"needles": needles, "caseSensitivity": r._case.name}}


def from_json(obj: dict, lower_table: LowerTableArg | None = None, device: int = -1) -> Replacer:
    """The derived FromJSON instance: `buildWithValues` over the stored (needle, Payload) pairs -- nothing is lowered,
    the payload lengths are taken as given.  Priorities must be 0, -1, -2, ... in list order (what `build` assigns,
    :101-111): the device keeps priority = -index."""
    try:
        s = obj["replacerSearcher"]
        case_ = CaseSensitivity[s["caseSensitivity"]]
        stored = []
        for i, (n, payload) in enumerate(s["needles"]):
            if payload["needlePriority"] != -i:
                raise ValueError("Replacer: needle %d has priority %r, expected %d" % (i, payload["needlePriority"], -i))
            stored.append(Stored(_b(n), int(payload["needleLengthBytes"]), int(payload["needleLengthCodePoints"]), _b(payload["needleReplacement"])))
    except (KeyError, TypeError) as e:
        raise ValueError("Replacer: malformed JSON (%s)" % e)
    return _from_stored(case_, stored, lower_table or default_lower_table(), device)
