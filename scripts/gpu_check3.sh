#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/check3.log 2>&1
( timeout 300 python scripts/perf_configs.py 2>&1 | tail -16 ) >> $O/check3.log 2>&1
for v in base _p2 _p8 _d16 _d48; do
  [ "$v" = "base" ] && v=""
  export AM_LIB=$PWD/alfred-margaret_b200/lib/libam_b200$v.so
  echo "=== variant '$v'" >> $O/check3.log
  ( timeout 200 python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all" ) >> $O/check3.log 2>&1
done
cat $O/check3.log
