#!/bin/bash
# On the GPU box: parity subset + vehicle-module bench for every tuning build in tools/variants/
set -u
mkdir -p gpurun_out
OUT=gpurun_out/variants_${1:-x}.txt; : > $OUT
for f in tools/variants/*.so; do
  echo "### $f" | tee -a $OUT
  if [ "${PARITY:-1}" = "1" ]; then
    ROBOTICK_LIB=$PWD/$f timeout 900 python -m pytest tests/test_vehicle_gpu.py tests/test_tick_gpu.py tests/test_vdt_task_gpu.py -q -x -k "not full_size and not two_gpu" 2>&1 | tail -2 | tee -a $OUT
  fi
  for r in 1 2; do
  ROBOTICK_LIB=$PWD/$f python bench.py --workload vehicle --steps 5 --warmup 3 --no-e2e --no-cpu --no-modules 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vehicle %.4g steps/s  %.3f ms' % (d['value'], d['ms_per_step']), d['clocks']['sm_mhz'])" | tee -a $OUT
  done
done
