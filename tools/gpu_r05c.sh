#!/bin/bash
python -m pytest tests/test_tick_gpu.py tests/test_streams_gpu.py tests/test_imu_gpu.py -q -x -k "not two_gpu and not full_size" 2>&1 | tail -3
for f in "" "--e2e-tables"; do
python bench.py --steps 5 --warmup 3 --no-cpu --no-modules $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('$f value %.4g  ms/step %.3f  e2e %.4g e2e ms %.3f ratio %.3f' % (d['value'], d['ms_per_step'], e.get('value'), e.get('ms_per_step'), e.get('value')/d['value']))"
done
