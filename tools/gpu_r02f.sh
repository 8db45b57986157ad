#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02f
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
for G in 2 1 4 0; do
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-modules --gen-ctas $G > gpurun_out/bench_full_gen${G}_$T.json 2> gpurun_out/bench_full_gen${G}_$T.err; echo "full gen-ctas $G rc=$?"
done
timeout 300 python tools/bench_modules.py --only arm --reps 5 > gpurun_out/arm_$T.jsonl 2>&1; cat gpurun_out/arm_$T.jsonl | cut -c1-260
for f in gpurun_out/bench_full_gen*_$T.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f" % (d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"]))
PY
done
