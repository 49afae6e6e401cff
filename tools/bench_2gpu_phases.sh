#!/bin/bash
# development helper: two-GPU bench line with per-rank phase timing
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 > $O/bench_2gpu_b.json 2> $O/bench_2gpu_b.err
grep -o '"phases_ms": {.*}}, "roofline' $O/bench_2gpu_b.json; grep -o '"ms_per_step": [0-9.]*' $O/bench_2gpu_b.json; grep -o '"value": [0-9.]*' $O/bench_2gpu_b.json | head -1; grep -o '"clocks": {[^}]*}' $O/bench_2gpu_b.json; grep -o '"consistent": [a-z]*' $O/bench_2gpu_b.json; tail -2 $O/bench_2gpu_b.err
