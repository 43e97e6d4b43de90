import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "alfred-margaret_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import am_oracle_py as O
from alfred_margaret_b200 import automaton
from helpers import as_pairs
cases = [
  (["$", "£"], "$€£𐍈"), (["$"], "$"), (["$"], "a$b"), (["ab"], "xxabxx"), (["a"], "a"), (["a", "bcd"], "abcd a"),
  (["abc", "rst", "xyz"], "abcdefghijklmnopqrstuvwxyz"), (["£"], "$€£𐍈"), (["a","aa"], "aaaa"),
]
for needles, hay in cases:
    want = as_pairs(O.Machine(needles).find_all(hay))
    for kind in (0, 1):
        m = automaton.AcMachine([(n, i) for i, n in enumerate(needles)], force_kernel=kind)
        got = as_pairs(m.find_all(hay)); c = m.count_matches(hay)
        print(needles, repr(hay), "kind", kind, m.info(), "OK" if got == want and c == len(want) else "MISMATCH got=%s count=%d want=%s" % (got, c, want))
