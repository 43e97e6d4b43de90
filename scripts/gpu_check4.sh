#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/check4.log 2>&1
( timeout 200 python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all" ) >> $O/check4.log 2>&1
( timeout 300 python scripts/full_configs.py C1 C2 2>&1 | tail -2 ) >> $O/check4.log 2>&1
cat $O/check4.log
