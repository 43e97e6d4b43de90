#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/check6.log 2>&1
( timeout 300 python scripts/perf_configs.py 2>&1 | grep -E "IC|10k CS auto" ) >> $O/check6.log 2>&1
( AM_IC_ONE_PASS=0 timeout 300 python scripts/perf_configs.py 2>&1 | grep -E "IC auto" | sed 's/^/two-pass: /' ) >> $O/check6.log 2>&1
( timeout 200 python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all" ) >> $O/check6.log 2>&1
cat $O/check6.log
