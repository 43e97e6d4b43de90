#!/bin/bash
# A/B of filter-kernel build variants on the headline workload (development aid): count + find_all at 4 GiB.
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-""}; do
  [ "$v" = "base" ] && v=""
  export AM_LIB=$PWD/alfred-margaret_b200/lib/libam_b200$v.so
  echo "=== variant '$v'"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all"
done
