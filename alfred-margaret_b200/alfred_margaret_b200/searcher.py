"""Host mirror of `Data.Text.AhoCorasick.Searcher` (src/Data/Text/AhoCorasick/Searcher.hs)."""
from __future__ import annotations

from typing import Any, Sequence, Tuple

from .automaton import AcMachine
from .case_sensitivity import CaseSensitivity


class Searcher:
    """`Searcher v` (:61-66): case flag, needles, count, automaton.  Eq/Hashable by needles only (:82-90)."""

    def __init__(self, case_sensitivity, needles_with_values: Sequence[Tuple[Any, Any]], **kw):
        self._case = CaseSensitivity(case_sensitivity)
        self._needles = list(needles_with_values)
        self._kw = kw
        self._automaton = AcMachine(self._needles, case_sensitivity=self._case, **kw)

    def __eq__(self, other):
        return isinstance(other, Searcher) and (len(self._needles), self._needles, self._case) == (len(other._needles), other._needles, other._case)

    def __hash__(self):
        return hash(tuple((bytes(n, "utf-8") if isinstance(n, str) else bytes(n), v) for n, v in self._needles))


def build(case_sensitivity, needles: Sequence, **kw) -> Searcher:
    """`build :: CaseSensitivity -> [Text] -> Searcher ()` (:110-111)."""
    return build_with_values(case_sensitivity, [(n, ()) for n in needles], **kw)


def build_with_values(case_sensitivity, needles_with_values, **kw) -> Searcher:
    """`buildWithValues` (:115-118).  For IgnoreCase the caller passes lower-case needles."""
    return Searcher(case_sensitivity, needles_with_values, **kw)


def build_needle_id_searcher(case_sensitivity, needles: Sequence, **kw) -> Searcher:
    """`buildNeedleIdSearcher` (:167-169)."""
    return build_with_values(case_sensitivity, [(n, i) for i, n in enumerate(needles)], **kw)


def needles(s: Searcher):
    return list(s._needles)


def num_needles(s: Searcher) -> int:
    return len(s._needles)


def automaton(s: Searcher) -> AcMachine:
    return s._automaton


def case_sensitivity(s: Searcher) -> CaseSensitivity:
    return s._case


def set_case_sensitivity(case_, s: Searcher) -> Searcher:
    """`setCaseSensitivity` (:142-145): flips the flag; needles and automaton are shared, nothing is rebuilt."""
    out = Searcher.__new__(Searcher)
    out._case, out._kw, out._needles = CaseSensitivity(case_), s._kw, s._needles
    out._automaton = s._automaton.with_case(out._case)
    return out


def map_searcher(f, s: Searcher) -> Searcher:
    """`mapSearcher` (:121-125): payloads live on the host, so only the value table changes."""
    out = Searcher.__new__(Searcher)
    out._case, out._kw = s._case, s._kw
    out._needles = [(n, f(v)) for n, v in s._needles]
    out._automaton = s._automaton.with_values([v for _, v in out._needles])
    return out


def contains_any(s: Searcher, text) -> bool:
    """`containsAny` (:156-164): the fold returns `Done True` on the first match == (count > 0)."""
    return s._automaton.contains_any(text, case=s._case)


def contains_all(s: Searcher, text) -> bool:
    """`containsAll` (:173-187), for searchers from build_needle_id_searcher."""
    return s._automaton.contains_all(text, case=s._case)


# ---- aeson-compatible JSON (`instance ToJSON / FromJSON (Searcher v)`, :68-77) -------------------------------------
def _text(x) -> str:
    return x if isinstance(x, str) else bytes(x).decode("utf-8")


def to_json(s: Searcher, value_to_json=lambda v: [] if v == () else v) -> dict:
    """`object ["needles" .= needles s, "caseSensitivity" .= caseSensitivity s]`: needles as [text, value] pairs (aeson
    encodes a tuple as an array and `()` as []), the case as the constructor name."""
    return {"needles": [[_text(n), value_to_json(v)] for n, v in s._needles], "caseSensitivity": s._case.name}


def from_json(obj: dict, value_from_json=lambda j: () if j == [] else j, **kw) -> Searcher:
    """`buildWithValues <$> o .: "caseSensitivity" <*> o .: "needles"`."""
    if not isinstance(obj, dict) or "needles" not in obj or "caseSensitivity" not in obj:
        raise ValueError("Searcher: expected an object with needles and caseSensitivity")
    return build_with_values(CaseSensitivity[obj["caseSensitivity"]], [(n, value_from_json(v)) for n, v in obj["needles"]], **kw)
