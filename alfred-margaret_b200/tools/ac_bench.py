#!/usr/bin/env python
"""`ac-bench` for libam_b200: the reference's cross-language benchmark protocol.

Same contract as benchmark/haskell/app/Main.hs:42-76 so that benchmark/benchmark.py can drive it:
  * each argument is a file "needle\\nneedle\\n...\\n\\nhaystack" (UTF-8; split at the first empty line, :26-40);
  * 5 iterations per file; an iteration times  build automaton + count all matches  (:61-76);
  * the five durations in nanoseconds are printed tab-separated on one stdout line, the match count goes to stderr
    on the first iteration (:52-57).
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def read_needle_haystack_file(path):
    """`readNeedleHaystackFile` (benchmark/haskell/app/Main.hs:26-40)."""
    data = open(path, "rb").read()
    needles, pos = [], 0
    while pos < len(data):
        nl = data.find(b"\n", pos)
        if nl == pos:                       # empty line: the rest is the haystack
            return needles, data[pos + 1:]
        if nl < 0:                          # no newline left: a last needle without haystack
            needles.append(data[pos:])
            return needles, b""
        needles.append(data[pos:nl])
        pos = nl + 1
    return needles, b""


def count_matches(needles, haystack):
    """`countMatches` (:67-76): build + `runText 0 (\\n _ -> Step (n + 1))`."""
    from alfred_margaret_b200 import automaton
    if not needles:
        return 0
    return automaton.AcMachine([(n, ()) for n in needles]).count_matches(haystack)


def main(argv):
    for path in argv:
        needles, haystack = read_needle_haystack_file(path)
        out = []
        for i in range(5):
            t0 = time.perf_counter_ns()
            count = count_matches(needles, haystack)
            out.append(str(time.perf_counter_ns() - t0))
            if i == 0:
                print(count, file=sys.stderr)
        print("\t".join(out) + "\t")


if __name__ == "__main__":
    main(sys.argv[1:])
