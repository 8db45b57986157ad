#!/bin/bash
set -u
run() { ROBOTICK_LIB=$PWD/$1 python bench.py --steps 5 --warmup 3 --no-cpu --no-modules --gen-ctas $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('$1 gen-ctas=$2 value %.4g  e2e %.4g  e2e ms %.3f ratio %.3f' % (d['value'], e.get('value'), e.get('ms_per_step'), e.get('value')/d['value']))"; }
run tools/variants/lib_sb128.so 4
run tools/variants/lib_sb128.so 8
run tools/variants/lib_sb512.so 1
run tools/variants/lib_sb512.so 2
run tools/variants/lib_sb1024.so 1
