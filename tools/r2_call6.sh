#!/bin/bash
# round-2 call 6 (1 GPU): whole GPU suite + default bench (train legs, cfg4 strong leg, cfg5 encode leg)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > $O/r2_t6.log; cat $O/r2_t6.log
timeout 1500 python bench.py > $O/r2_bench6.json 2> $O/r2_bench6.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench6.json'))
    print(json.dumps({k:d[k] for k in ('value','ms_per_step','merges_per_s','encode_cfg5')}, indent=None)[:4000])
except Exception as e: print("ERR", e)
PY
tail -5 $O/r2_bench6.err
