"""Shared test helpers: the reference's generators (tests/Data/Text/TestInstances.hs) and naive oracles."""
import numpy as np

SIMPLE_ALPHABET = "abAB12"                       # TestInstances.hs:81
FANCY_ALPHABET = "яЯåÅÅ𝄞💩ßẞ"     # TestInstances.hs:83-90 (Å U+212B and Å U+00C5 both lower to å)


def random_alphabet(rng):                        # TestInstances.hs:92 (8 arbitrary chars)
    out = []
    while len(out) < 8:
        cp = int(rng.integers(1, 0x2FFF)) if rng.random() < 0.9 else int(rng.integers(0x10000, 0x1FFFF))
        if 0xD800 <= cp <= 0xDFFF:
            continue
        out.append(chr(cp))
    return "".join(out)


def needles_haystack(rng, max_needles=6, big=40):
    """arbitraryNeedlesHaystack (TestInstances.hs:60-71): needles and haystack from one fragment pool."""
    alphabet = [SIMPLE_ALPHABET, FANCY_ALPHABET, random_alphabet(rng)][int(rng.integers(0, 3))]
    frags = ["".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=int(rng.integers(1, 6))))
             for _ in range(int(rng.integers(1, 8)))]
    def cat(lo, hi):
        return "".join(frags[int(i)] for i in rng.integers(0, len(frags), size=int(rng.integers(lo, hi + 1))))
    needles = [cat(1, 3) for _ in range(int(rng.integers(1, max_needles + 1)))]
    haystack = cat(1, big)
    return needles, haystack


def naive_find_all(needles, haystack: bytes):
    """All (end_pos, needle_index) by brute force, in the reference's callback order."""
    out = []
    for i, n in enumerate(needles):
        nb = n if isinstance(n, bytes) else n.encode("utf-8")
        if not nb:
            continue
        start = 0
        while True:
            k = haystack.find(nb, start)
            if k < 0:
                break
            out.append((k + len(nb), -len(nb), -i, i))
            start = k + 1
    out.sort()
    return [(e, i) for e, _, _, i in out]


def as_pairs(structured):
    """structured (pos/end_pos, value/needle_id) array -> list of tuples"""
    names = structured.dtype.names
    return list(zip(structured[names[0]].tolist(), structured[names[1]].tolist()))
