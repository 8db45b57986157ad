#!/bin/bash
for g in -148 -74 -37 -20; do
  python bench.py --steps 4 --warmup 2 --no-cpu --no-modules --gen-ctas=$g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('gen-ctas=%-5s value %.4g  e2e %.4g  e2e ms %.3f ratio %.3f' % ('$g', d['value'], e.get('value'), e.get('ms_per_step'), e.get('value')/d['value']))"
done
