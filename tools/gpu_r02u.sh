#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02u}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$T.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-modules > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"; tail -2 gpurun_out/bench_full_$T.err
python - gpurun_out/bench_full_$T.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f" % (d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"]))
PY
