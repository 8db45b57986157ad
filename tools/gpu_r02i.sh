#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02i
for f in tools/variants/lib_armocc*.so; do
  echo "### $f"
  ROBOTICK_LIB=$PWD/$f python tools/bench_modules.py --only arm --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arm module %.4g ticks/s %.3f ms' % (d['arm_ticks_per_s'], d['ms_per_launch']))"
  for SC in 1 2; do
    ROBOTICK_LIB=$PWD/$f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-modules --side-ctas $SC 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full side-ctas $SC: %.4g  %.3f ms/step' % (d['value'], d['ms_per_step']))"
  done
done 2>&1 | tee gpurun_out/armocc_$T.txt
for SC in 2 3; do python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-modules --side-ctas $SC 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('main lib full side-ctas $SC: %.4g  %.3f ms/step' % (d['value'], d['ms_per_step']))"; done 2>&1 | tee -a gpurun_out/armocc_$T.txt
