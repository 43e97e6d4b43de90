#!/bin/bash
# A/B the filter-kernel build variants (development aid): parity smoke + stage-isolated timings.
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-_dbg}; do
  export AM_LIB=$PWD/alfred-margaret_b200/lib/libam_b200$v.so
  echo "=== variant '$v'"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  for f in 1 2 0; do AM_DEBUG_FLAGS=$f python scripts/quick_perf.py 4294967296 1000 --count-only 2>&1 | grep count | sed "s/^/flags=$f /"; done
done
