#!/bin/bash
# round-2 call 1: validate the merged next/* branches on one GPU, record box facts
cd "$(dirname "$0")/.."
O=gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; ls gpurun_probe_late.txt 2>&1; python -c "import os;print(len(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; } > $O/r2_box.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/r2_t1.log; cat $O/r2_t1.log
timeout 600 python bench.py --extras > $O/r2_extras1.json 2> $O/r2_extras1.err; tail -c 1500 $O/r2_extras1.json; tail -3 $O/r2_extras1.err
python - <<'PY' > $O/r2_synth.txt 2>&1
import time, os
from minbpe_b200.synth import generate
for thr in (1, 8, 32, 64):
    t0=time.time(); a=generate(1338, 1<<30, threads=thr); print(thr, "threads 1GiB", round(time.time()-t0,2), "s")
PY
cat $O/r2_synth.txt $O/r2_box.txt
