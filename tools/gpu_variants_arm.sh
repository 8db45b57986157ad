#!/bin/bash
# On the GPU box: arm parity subset + arm-module bench (+ full tick) for every tuning build in tools/variants/
set -u
mkdir -p gpurun_out
OUT=gpurun_out/variants_arm_${1:-x}.txt; : > $OUT
for f in tools/variants/*.so; do
  echo "### $f" | tee -a $OUT
  if [ "${PARITY:-1}" = "1" ]; then
    ROBOTICK_LIB=$PWD/$f timeout 900 python -m pytest tests/test_arm_gpu.py tests/test_armhome_gpu.py tests/test_tick_gpu.py -q -x -k "not full_size and not two_gpu" 2>&1 | tail -2 | tee -a $OUT
  fi
  ROBOTICK_LIB=$PWD/$f python tools/bench_modules.py --only arm --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arm module %.4g ticks/s %.3f ms' % (d['arm_ticks_per_s'], d['ms_per_launch']))" | tee -a $OUT
  if [ "${FULL:-1}" = "1" ]; then
  ROBOTICK_LIB=$PWD/$f python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-modules 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full %.4g  %.3f ms/step' % (d['value'], d['ms_per_step']))" | tee -a $OUT
  fi
done
