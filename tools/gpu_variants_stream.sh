#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/variants_stream_${1:-x}.txt; : > $OUT
for f in tools/variants/*.so; do
  echo "### $f" | tee -a $OUT
  ROBOTICK_LIB=$PWD/$f timeout 900 python -m pytest tests/test_vehicle_gpu.py -q -x -k "stream" 2>&1 | tail -2 | tee -a $OUT
  ROBOTICK_LIB=$PWD/$f python tools/bench_modules.py --only stream --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ('roofline','workload','kernel')}); print(d.get('roofline'))" | tee -a $OUT
done
