#!/bin/bash
# development helper (round-end validation on the GPU box): pick the fastest launch configuration of the merge
# kernel among the prebuilt variants, then run the whole validation + measurement set with it.
cd "$(dirname "$0")/.."
O=gpurun_out
bash tools/bench_variants.sh > $O/variants4.txt 2>&1
cat $O/variants4.txt
python - <<'PY' > $O/chosen.txt
import re
best=None; base=None
for line in open('gpurun_out/variants4.txt'):
    m=re.match(r'(\S+) ms/step \S+ kernel ms (\S+)', line)
    if not m: continue
    p,ms=m.group(1),float(m.group(2))
    if p.endswith('lib_w8_b4_t8.so'): base=(ms,p)
    if best is None or ms<best[0]: best=(ms,p)
if base and best and best[0] > 0.98*base[0]: best=base
print(best[1] if best else '')
PY
CH=$(cat $O/chosen.txt); echo "chosen: $CH"
[ -n "$CH" ] && cp "$CH" minbpe_b200/csrc/libb200bpe.so
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; tail -c 600 $O/bench_final.json
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.json
timeout 600 python bench.py --extras > $O/extras_final.json 2> $O/extras_final.err; tail -c 300 $O/extras_final.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_merge_seg -s 8 -c 1 -o $O/prof_seg_final python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_final.log 2>&1; tail -1 $O/ncu_final.log | cut -c1-120
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --size-mib 256 > $O/b_launch.log 2>&1; tail -c 100 $O/b_launch.log
