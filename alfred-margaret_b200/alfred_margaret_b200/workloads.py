"""The five configurations of BASELINE.json as reproducible synthetic workloads (bench / test tooling).

SURVEY.md section 8d fixes their shapes; every generator is a pure function of its seeds, so the CPU oracle can be
handed any window of a device-resident multi-GiB haystack.

  C1  3 needles [tshirt, shirts, shorts], CaseSensitive, 1 MB of README sentences           (plumbing)
  C2  1 000 random 4-16 B a-z needles, 4 GiB a-z text, one needle planted per 4 KiB          (the headline)
  C3  10 000 lower-case needles (20 % with non-ASCII code points), IgnoreCase, 8 GiB mixed-case UTF-8
  C4  Replacer of 5 000 (needle, replacement) pairs, 2 GiB text with plants from 64 needles
  C5  100 000 random 6-16 B a-z needles, 64 GiB sharded over the GPUs
"""
from __future__ import annotations

import numpy as np

from . import synth

GIB = 1 << 30
C2_SEEDS = (42, 43, 44)          # needles, text, plants
C4_SEEDS = (62, 63, 64)
C5_SEEDS = (72, 73, 74)
C3_UNIT = 16 << 20               # the C3 text is a 16 MiB unit (ending on a code point boundary) repeated


def c1():
    needles = [b"tshirt", b"shirts", b"shorts"]
    rng = np.random.default_rng(1)
    sentences = ["short tshirts ", "sweatshirts and shirtshirts ", "long shirt "]
    hay = "".join(sentences[int(i)] for i in rng.integers(0, 3, size=80000))[:1000000].encode()
    return needles, np.frombuffer(hay, dtype=np.uint8).copy()


def c2_needles(n: int = 1000):
    return synth.random_needles(n, C2_SEEDS[0])


def c4_pairs(n: int = 5000):
    rng = np.random.default_rng(C4_SEEDS[2])
    needles = synth.random_needles(n, C4_SEEDS[0], 4, 16)
    repls = [bytes(rng.integers(ord("A"), ord("Z") + 1, size=int(rng.integers(0, 25)), dtype=np.uint8)) for _ in needles]
    return needles, repls


def c5_needles(n: int = 100000):
    return synth.random_needles(n, C5_SEEDS[0], 6, 16)


def c3_needles(n: int = 10000):
    """Lower-case needles of 4..16 code points; 20 % draw from a pool that holds non-ASCII (already lower-case) code points."""
    rng = np.random.default_rng(52)
    ascii_l = "abcdefghijklmnopqrstuvwxyz"
    extra = "éößåяωǳⱥ"
    nset = set()
    while len(nset) < n:
        k = int(rng.integers(4, 17))
        pool = ascii_l + (extra * 3 if rng.random() < 0.2 else "")
        nset.add("".join(pool[int(i)] for i in rng.integers(0, len(pool), size=k)))
    return [s.encode("utf-8") for s in sorted(nset)]


def c3_unit(needles, size: int = C3_UNIT, seed: int = 53) -> np.ndarray:
    """`size` bytes of mixed-case UTF-8: 70 % ASCII letters of either case, 10 % space / punctuation,
    15 % two-byte code points (the letters of the Latin-1 supplement and of the Cyrillic block in either case, plus ω Ω and the
    DZ digraphs so that every needle letter occurs), 4 % three-byte (incl. K U+212A, Å U+212B, ẞ, whose lower case has another
    UTF-8 length; ⱥ; general punctuation), 1 % four-byte; about one randomly re-cased needle per 3 500 symbols.  Ends on a
    code point boundary (padded with spaces), so units can be laid back to back."""
    rng = np.random.default_rng(seed)
    ascii_l = "abcdefghijklmnopqrstuvwxyz"
    syms, wts = [], []

    def add(chars, total):
        for c in chars:
            syms.append(c.encode("utf-8")); wts.append(total / len(chars))
    latin1 = "".join(chr(c) for c in range(0xC0, 0x100) if c not in (0xD7, 0xF7))
    cyrillic = "".join(chr(c) for c in range(0x410, 0x450))
    add(ascii_l + ascii_l.upper(), 0.70); add(" .,;-", 0.10); add(latin1 + cyrillic + "\u03c9\u03a9\u01f3\u01f2\u01f1", 0.15)
    add("\u1e9e\u212a\u212b\u2c65\u20ac\u2013\u2026\u201c\u201d", 0.04); add("\U0001d11e\U0001f4a9", 0.01)
    plants = []
    for _ in range(512):
        nd = needles[int(rng.integers(0, len(needles)))].decode("utf-8")
        plants.append("".join((c.upper() if (rng.random() < 0.5 and len(c.upper()) == 1) else c) for c in nd).encode("utf-8"))
    for p in plants:
        syms.append(p); wts.append(1.0 / 3500 / len(plants))
    wts = np.array(wts); wts /= wts.sum()
    lens_tab = np.array([len(s) for s in syms], dtype=np.int64)
    maxlen = int(lens_tab.max())
    tab = np.zeros((len(syms), maxlen), dtype=np.uint8)
    for i, s in enumerate(syms):
        tab[i, : len(s)] = np.frombuffer(s, dtype=np.uint8)
    out = np.full(size, ord(" "), dtype=np.uint8)
    cdf = np.cumsum(wts)
    filled = 0
    while filled < size:
        want = (size - filled) // 1 + 16
        k = int(min(want, 16 << 20))
        idx = np.minimum(np.searchsorted(cdf, rng.random(k)), len(syms) - 1)
        lens = lens_tab[idx]
        ends = np.cumsum(lens)
        keep = ends <= size - filled
        idx, lens, ends = idx[keep], lens[keep], ends[keep]
        if idx.size == 0:
            break                                               # the rest stays spaces
        starts = ends - lens + filled
        for j in range(maxlen):
            msk = lens > j
            if not msk.any():
                break
            out[starts[msk] + j] = tab[idx[msk], j]
        filled += int(ends[-1])
        if idx.size < k:
            break
    return out
