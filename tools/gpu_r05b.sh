#!/bin/bash
python -m pytest tests/test_vehicle_gpu.py tests/test_tick_gpu.py -q -x -k "reset or tick or fast_kernel" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu --no-modules 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('value %.4g  ms/step %.3f  e2e %.4g e2e ms %.3f ratio %.3f' % (d['value'], d['ms_per_step'], e.get('value'), e.get('ms_per_step'), e.get('value')/d['value']))"
python bench.py --workload vehicle --steps 5 --warmup 3 --no-cpu --no-modules 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get('e2e') or {}; print('vehicle %.4g %.3f ms e2e %.4g' % (d['value'], d['ms_per_step'], e.get('value')))"
