#!/bin/bash
# One GPU call: parity tests on the default build, then A/B of the variant builds (development aid).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) >> gpurun_out/ab.log 2>&1
for v in ${VARIANTS:-base _c2 _c1 _c2hi _c1hi _c2nt _c2ol}; do
  [ "$v" = "base" ] && v=""
  export AM_LIB=$PWD/alfred-margaret_b200/lib/libam_b200$v.so
  echo "=== variant '$v'" >> gpurun_out/ab.log
  ( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ) >> gpurun_out/ab.log 2>&1
  ( timeout 200 python scripts/quick_perf.py 4294967296 1000 2>&1 | grep -E "count|find_all" ) >> gpurun_out/ab.log 2>&1
done
export AM_LIB=$PWD/alfred-margaret_b200/lib/libam_b200_dbg.so
echo "=== stage isolation (_dbg)" >> gpurun_out/ab.log
for f in 1 2 0; do ( AM_DEBUG_FLAGS=$f timeout 200 python scripts/quick_perf.py 4294967296 1000 --count-only 2>&1 | grep count | sed "s/^/flags=$f /" ) >> gpurun_out/ab.log 2>&1; done
unset AM_LIB
( timeout 200 python scripts/perf_configs.py 2>&1 | tail -12 ) >> gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
