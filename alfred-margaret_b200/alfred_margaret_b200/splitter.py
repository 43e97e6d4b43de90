"""Host mirror of `Data.Text.AhoCorasick.Splitter` (src/Data/Text/AhoCorasick/Splitter.hs) -- SURVEY.md 8f rank 2.

A single-needle automaton; splitting is the reference's fold (`stepAccum` :158-170, `finalizeAccum` :140-147)
over the ordered match list the device returns.  Like the reference (:63-67) a Splitter holds ONE machine and runs
it in either case mode.
"""
from __future__ import annotations

from typing import List

from .automaton import AcMachine
from .case_sensitivity import CaseSensitivity
from .utf8 import as_text, skip_code_points_backwards


class Splitter:
    def __init__(self, separator, **kw):
        self._sep = separator.encode("utf-8") if isinstance(separator, str) else bytes(separator)
        self._kw = kw
        self._machine = AcMachine([(self._sep, ())], **kw)

    def machine(self, cs=CaseSensitivity.CaseSensitive) -> AcMachine:
        return self._machine

    def __eq__(self, other):                       # instance Eq Splitter (:175-177)
        return isinstance(other, Splitter) and self._sep == other._sep

    def __hash__(self):
        return hash(self._sep)


def build(sep, **kw) -> Splitter:
    """`build :: Text -> Splitter` (:63-67)."""
    return Splitter(sep, **kw)


def separator(s: Splitter) -> bytes:
    return s._sep


def automaton(s: Splitter) -> AcMachine:
    return s.machine(CaseSensitivity.CaseSensitive)


def _split_reverse(s: Splitter, text, ignore_case: bool) -> List[bytes]:
    t = as_text(text)
    hay = t.tobytes()
    ends = s.machine().find_all(t, case=CaseSensitivity.IgnoreCase if ignore_case else CaseSensitivity.CaseSensitive)["end_pos"].tolist()
    res, fragment_start = [], 0                    # zeroAccum (:150-152)
    if ignore_case:
        sep_len = sum((b & 0xC0) != 0x80 for b in s._sep)        # Text.length (separator s): code points (:113)
    else:
        sep_len = len(s._sep)                                     # lengthUtf8: bytes (:103)
    for new_fragment_start in ends:
        if ignore_case:
            sep_start = skip_code_points_backwards(hay, new_fragment_start - 1, sep_len - 1)   # (:116)
        else:
            sep_start = new_fragment_start - sep_len
        if sep_start < fragment_start:             # overlaps the previous separator: ignored (:163-164)
            continue
        res.append(hay[fragment_start:sep_start])  # the reference conses (:165); appended here, reversed once below
        fragment_start = new_fragment_start
    res.append(hay[fragment_start:])               # finalizeAccum (:140-147)
    res.reverse()
    return res


def split_reverse(s: Splitter, text) -> List[bytes]:
    """`splitReverse` (:98-106)."""
    return _split_reverse(s, text, False)


def split_reverse_ignore_case(s: Splitter, text) -> List[bytes]:
    """`splitReverseIgnoreCase` (:109-118); the separator must be lower case."""
    return _split_reverse(s, text, True)


def split(s: Splitter, text) -> List[bytes]:
    """`split = NonEmpty.reverse . splitReverse` (:84-85)."""
    return list(reversed(split_reverse(s, text)))


def split_ignore_case(s: Splitter, text) -> List[bytes]:
    """`splitIgnoreCase` (:95-96)."""
    return list(reversed(split_reverse_ignore_case(s, text)))


# ---- aeson-compatible JSON (`toJSON = toJSON . separator`, `parseJSON v = build <$> parseJSON v`, :54-60) -----------------
def to_json(s: Splitter) -> str:
    return s._sep.decode("utf-8")


def from_json(obj, **kw) -> Splitter:
    if not isinstance(obj, str):
        raise ValueError("Splitter: expected a JSON string")
    return build(obj, **kw)
