#!/bin/bash
# development helper: validation of the final tree on one GPU
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > $O/bench_a.json 2> $O/bench_a.err; tail -c 700 $O/bench_a.json; tail -3 $O/bench_a.err
MASTER_ADDR=127.0.0.1 MASTER_PORT=29511 RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 BPE_BENCH_FORCE_SHARDED=1 timeout 600 python bench.py --size-mib 256 > $O/bench_sharded_w1.json 2> $O/bench_sharded_w1.err; tail -c 900 $O/bench_sharded_w1.json; tail -3 $O/bench_sharded_w1.err
