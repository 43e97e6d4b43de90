#!/bin/bash
# One GPU call for the record: parity tests, both bench arms, ncu launch list and ncu --set full of the scan kernel.
cd "$(dirname "$0")/.."
TAG=${TAG:-r1c}
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/pytest_$TAG.log 2>&1
timeout 400 python bench.py --impl reference > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
timeout 400 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
# one full capture of the scan kernel (EMIT and COUNT), 2 GiB launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o $O/${TAG}_filter_emit -f python scripts/prof_one.py 2147483648 find_all > $O/ncu_emit.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 1 -c 1 -o $O/${TAG}_filter_count -f python scripts/prof_one.py 2147483648 count > $O/ncu_count.log 2>&1
python scripts/perf_configs.py > $O/perf_configs_$TAG.log 2>&1
timeout 900 python tests/full_configs.py > $O/full_configs_$TAG.json 2> $O/full_configs_$TAG.err
cat $O/pytest_$TAG.log $O/bench_ref_$TAG.json $O/bench_$TAG.json; tail -3 $O/ncu_emit.log; cat $O/perf_configs_$TAG.log $O/full_configs_$TAG.json
