#!/bin/bash
# On the GPU box: bench every tuning build in tools/variants/ (kernel-only value).
for f in tools/variants/*.so; do
  ROBOTICK_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu ${BENCH_ARGS:-} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', '%.4g steps/s  %.3f ms' % (d['value'], d['ms_per_step']), d['clocks']['sm_mhz'])"
done
