#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/check2.log 2>&1
( timeout 300 python scripts/perf_configs.py 2>&1 | tail -16 ) >> $O/check2.log 2>&1
( timeout 200 python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all" ) >> $O/check2.log 2>&1
( timeout 900 python scripts/full_configs.py > $O/full_configs.json 2> $O/full_configs.err; tail -5 $O/full_configs.err ) >> $O/check2.log 2>&1
cat $O/check2.log; cat $O/full_configs.json
