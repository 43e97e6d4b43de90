"""Host mirror of `Data.Text.AhoCorasick.Replacer` (src/Data/Text/AhoCorasick/Replacer.hs)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

from . import _ffi
from .automaton import _Needles
from .case_sensitivity import CaseSensitivity
from .utf8 import LowerTableArg, as_text, default_lower_table

MAX_BOUND = (1 << 64) - 1


class Replacer:
    """`Replacer` (:78-80).  Pair i has priority -i (:101-111)."""

    def __init__(self, case_sensitivity, replaces: Sequence[Tuple], lower_table: LowerTableArg | None = None, device: int = -1):
        self._case = CaseSensitivity(case_sensitivity)
        self._replaces = [(self._b(n), self._b(r)) for n, r in replaces]
        self._lower = lower_table or default_lower_table()
        self._device = device
        self._n = _Needles([n for n, _ in self._replaces])
        self._r = _Needles([r for _, r in self._replaces])
        opts = _ffi.Options(device, 0, (C.c_uint64 * 6)())
        h = C.c_void_p()
        _ffi.check(_ffi.lib().am_replacer_build(self._n.arr, self._r.arr, len(self._replaces), int(self._case),
                                                self._lower.ptr(), C.byref(opts), C.byref(h)))
        self.handle = h
        self.last_passes = 0
        self.last_rescans = 0   # passes that scanned the whole text (1 when the match list is carried between passes)

    @staticmethod
    def _b(x) -> bytes:
        return x.encode("utf-8") if isinstance(x, str) else bytes(x)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                _ffi.lib().am_replacer_free(h)
            except Exception:
                pass
            self.handle = None

    def __eq__(self, other):
        return isinstance(other, Replacer) and (self._case, self._replaces) == (other._case, other._replaces)

    def __hash__(self):
        return hash((self._case, tuple(self._replaces)))


def build(case_sensitivity, replaces: Sequence[Tuple], **kw) -> Replacer:
    """`build :: CaseSensitivity -> [(Needle, Replacement)] -> Replacer` (:97-116)."""
    return Replacer(case_sensitivity, replaces, **kw)


def compose(r1: Replacer, r2: Replacer) -> Optional[Replacer]:
    """`compose` (:120-133): r2 after r1; None if the case sensitivities differ."""
    if r1._case != r2._case:
        return None
    return Replacer(r1._case, r1._replaces + r2._replaces, lower_table=r1._lower, device=r1._device)


def map_replacement(f, r: Replacer) -> Replacer:
    """`mapReplacement` (:136-141)."""
    return Replacer(r._case, [(n, f(rep)) for n, rep in r._replaces], lower_table=r._lower, device=r._device)


def replacer_case_sensitivity(r: Replacer) -> CaseSensitivity:
    return r._case


def set_case_sensitivity(case_, r: Replacer) -> Replacer:
    """`setCaseSensitivity` (:151-153)."""
    return Replacer(case_, r._replaces, lower_table=r._lower, device=r._device)


def run_with_limit(r: Replacer, max_length: int, text) -> Optional[bytes]:
    """`runWithLimit :: Replacer -> CodeUnitIndex -> Text -> Maybe Text` (:203-242); None is `Nothing`."""
    t = as_text(text)
    out, out_len, exceeded = C.c_void_p(), C.c_uint64(), C.c_int()
    _ffi.check(_ffi.lib().am_replacer_run(r.handle, t.slice(), max_length, C.byref(out), C.byref(out_len), C.byref(exceeded)))
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


# ---- aeson-compatible JSON (generic instances of `Replacer` and `Payload`, :56-83) -----------------------------------
def to_json(r: Replacer) -> dict:
    """`Replacer { replacerSearcher :: Searcher Payload }` with the generic encodings: the searcher object of
    Searcher.hs:68-73 whose values are `Payload` records.  The stored needles are the ones the automaton holds
    (lowered for IgnoreCase, :105-107); the lengths are those of the ORIGINAL needle (:111-113)."""
    from .utf8 import lower_utf8
    needles = []
    for i, (n, rep) in enumerate(r._replaces):
        stored = lower_utf8(n, r._lower) if r._case == CaseSensitivity.IgnoreCase else n
        needles.append([stored.decode("utf-8"), {"needlePriority": -i, "needleLengthBytes": len(n),
                                                 "needleLengthCodePoints": len(n.decode("utf-8")), "needleReplacement": rep.decode("utf-8")}])
    return {"replacerSearcher": {"needles": needles, "caseSensitivity": r._case.name}}


def from_json(obj: dict, **kw) -> Replacer:
    """Rebuild from the stored (needle, Payload) pairs.  Priorities must be 0, -1, -2, ... in list order (that is what
    `build` assigns, :101-111); lowering is idempotent, so the stored needles can go through `build` again."""
    try:
        s = obj["replacerSearcher"]
        case_ = CaseSensitivity[s["caseSensitivity"]]
        pairs = []
        for i, (n, payload) in enumerate(s["needles"]):
            if payload["needlePriority"] != -i:
                raise ValueError("Replacer: needle %d has priority %r, expected %d" % (i, payload["needlePriority"], -i))
            pairs.append((n, payload["needleReplacement"]))
    except (KeyError, TypeError) as e:
        raise ValueError("Replacer: malformed JSON (%s)" % e)
    return build(case_, pairs, **kw)
