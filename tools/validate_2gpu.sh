#!/bin/bash
# development helper: two-GPU validation (sharded parity tests + the N=2 bench line)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 1200 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
