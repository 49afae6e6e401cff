#!/bin/bash
# development helper: two-GPU bench with NCCL transport log and per-phase step timing
cd "$(dirname "$0")"
O=gpurun_out
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 > $O/bench_2gpu_diag.json 2> $O/bench_2gpu_diag.err
grep -E " via |NVLS|Channel 00|Connected|transport|P2P|SHM" $O/bench_2gpu_diag.err | head -12
grep -o '"phases_ms": {[^}]*}' $O/bench_2gpu_diag.json; grep -o '"ms_per_step": [0-9.]*' $O/bench_2gpu_diag.json; grep -o '"clocks": {[^}]*}' $O/bench_2gpu_diag.json
nvidia-smi topo -m 2>/dev/null | head -8
