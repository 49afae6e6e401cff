#!/bin/bash
# round-2 call 4 (1 GPU): piecewise split, sharded path on one rank, N=1 bench with the 16 GiB strong leg
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_split.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -8 > $O/r2_t4.log; cat $O/r2_t4.log
MASTER_ADDR=127.0.0.1 MASTER_PORT=29511 RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 BPE_BENCH_FORCE_SHARDED=1 timeout 600 python bench.py --size-mib 256 --strong-gib 2 --strong-sparse-at 300 > $O/r2_bench4_w1.json 2> $O/r2_bench4_w1.err; tail -c 2500 $O/r2_bench4_w1.json; tail -5 $O/r2_bench4_w1.err
timeout 1200 python bench.py > $O/r2_bench4.json 2> $O/r2_bench4.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench4.json'))
    print(json.dumps({k:d[k] for k in ('value','ms_per_step','merges_per_s','strong_cfg4','full_run')}, indent=None)[:3000])
except Exception as e: print("ERR", e)
PY
tail -5 $O/r2_bench4.err
