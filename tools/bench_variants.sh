#!/bin/bash
# development helper: bench every kernel variant under minbpe_b200/csrc/variants on the 1 GiB workload
for so in minbpe_b200/csrc/variants/*.so; do
  BPE_LIB_PATH=$PWD/$so timeout 300 python bench.py --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('$so', 'ms/step %.3f' % d['ms_per_step'], 'kernel ms %.3f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'value %.1f' % d['value'], 'e2e %.1f (%.3fs load %.3fs)' % (d['e2e']['value'], d['e2e']['seconds'], d['e2e'].get('load_seconds', -1)))
"
done
