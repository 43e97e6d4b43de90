#!/bin/bash
# One GPU call for the record: parity tests, both bench arms, the ncu launch list of the bench command, ncu --set full
# captures of the kernels behind C2 / C3 / C5 (and the walk kernel C5 left), compute-sanitizer on small configurations.
#   TAG=r2x scripts/gpu_round.sh [quick]        (quick: tests + bench arms only)
cd "$(dirname "$0")/.."
TAG=${TAG:-r2}
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/pytest_$TAG.log 2>&1
timeout 600 python bench.py --impl reference > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
cat $O/pytest_$TAG.log; python scripts/show_bench.py $O/bench_$TAG.json; tail -c 400 $O/bench_ref_$TAG.json
[ "$1" = quick ] && exit 0
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --configs '' > $O/bench_under_ncu.log 2>&1
# full captures, one 2 GiB scan each (third scan of the process: warm)
for spec in c2:2:1 c3:4:2 c5:4:2 c5walk:1:1; do
  w=${spec%%:*}; r=${spec#*:}; skip=${r%%:*}; cnt=${r#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"filter_kernel|verify_kernel|walk_kernel" -s $skip -c $cnt -o $O/${TAG}_$w -f python scripts/prof_r2.py $w > $O/ncu_$w.log 2>&1
done
# compute-sanitizer on the small parity cases (C1 plumbing + the ABI driver): memcheck and racecheck
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or one_machine or edge_cases or long_qgram or stride2 or ignore_case_length or contains_all" > $O/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/${TAG}_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or one_machine or edge_cases" > $O/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/${TAG}_sanitizer_racecheck.log
tail -4 $O/${TAG}_sanitizer_memcheck.log $O/${TAG}_sanitizer_racecheck.log
ls -la $O | grep $TAG
