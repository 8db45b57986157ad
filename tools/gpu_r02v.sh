#!/bin/bash
set -u
mkdir -p gpurun_out
for sk in "" gen reset d2h gen,reset,d2h; do
  python bench.py --steps 5 --warmup 3 --no-cpu --no-modules --e2e-skip "$sk" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('skip=%-14s value %.4g  ms/step %.3f  e2e %.4g  e2e ms %.3f ratio %.3f' % ('$sk', d['value'], d['ms_per_step'], e.get('value'), e.get('ms_per_step'), e.get('value')/d['value']))"
done
for g in 0 1 4 8; do
  python bench.py --steps 5 --warmup 3 --no-cpu --no-modules --gen-ctas $g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('gen-ctas=%-3s value %.4g  ms/step %.3f  e2e %.4g  e2e ms %.3f ratio %.3f' % ('$g', d['value'], d['ms_per_step'], e.get('value'), e.get('ms_per_step'), e.get('value')/d['value']))"
done
