#!/bin/bash
# round-2 call 2: cfg3 full run (test + bench leg)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_full.py -x -q -m gpu 2>&1 | tail -15 > $O/r2_t2.log; cat $O/r2_t2.log
timeout 900 python bench.py > $O/r2_bench2.json 2> $O/r2_bench2.err; tail -c 3000 $O/r2_bench2.json; tail -5 $O/r2_bench2.err
timeout 600 python bench.py --impl reference > $O/r2_ref2.json 2> $O/r2_ref2.err; tail -c 1500 $O/r2_ref2.json; tail -3 $O/r2_ref2.err
