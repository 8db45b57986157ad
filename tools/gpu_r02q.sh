#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02q
timeout 900 python -m pytest tests/test_vehicle_gpu.py -q -x -k "stream" 2>&1 | tail -2
python tools/bench_modules.py --only stream --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ('roofline','workload','kernel')}); print(d.get('roofline'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vdt_rollout_stream -s 4 -c 1 -f -o gpurun_out/prof_stream_$T python tools/bench_modules.py --only stream --reps 2 > gpurun_out/ncu_full_stream_$T.log 2>&1; echo "ncu rc=$?"
