#!/bin/bash
# 2-GPU call: parity tests on GPU 0, then bench.py at N = 1 and N = 2 (torchrun, NCCL all-gather of the match counts).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $O/n2.log 2>&1
timeout 300 python bench.py > $O/bench_n1_r1e.json 2> $O/bench_n1_r1e.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_r1e.json 2> $O/bench_n2_r1e.err
cat $O/n2.log $O/bench_n1_r1e.json $O/bench_n2_r1e.json; tail -3 $O/bench_n2_r1e.err
