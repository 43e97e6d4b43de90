#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/check5.log 2>&1
( timeout 300 python scripts/perf_host_paths.py 2>&1 | tail -6 ) >> $O/check5.log 2>&1
( timeout 300 python bench.py 2>&1 | tail -1 ) >> $O/check5.log 2>&1
cat $O/check5.log
