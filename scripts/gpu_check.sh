#!/bin/bash
# GPU call: parity tests + replacer timing (carried match list vs a full scan per pass) + headline kernel timing.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/check.log 2>&1
( timeout 300 python scripts/perf_replacer.py 2>&1 | tail -3 ) >> $O/check.log 2>&1
( AM_REPLACER_RESCAN=1 timeout 300 python scripts/perf_replacer.py 2>&1 | tail -3 ) >> $O/check.log 2>&1
( timeout 200 python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all" ) >> $O/check.log 2>&1
cat $O/check.log
