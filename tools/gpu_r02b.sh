#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$T.log
for V in "--lanes 2 --side-ctas 1" "--lanes 1 --side-ctas 1" "--lanes 2 --side-ctas 0" "--lanes 1 --side-ctas 0" "--lanes 2 --side-ctas 2"; do
  echo "### $V"; timeout 300 python tools/tick_timeline.py $V 2>&1 | tail -12
done > gpurun_out/timeline_$T.txt 2>&1
cat gpurun_out/timeline_$T.txt
B="python bench.py --no-cpu"
timeout 300 $B --workload vehicle --steps 5 --warmup 3 > gpurun_out/bench_vehicle_$T.json 2> gpurun_out/bench_vehicle_$T.err; echo "vehicle rc=$?"
SHORT="python bench.py --workload vehicle --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vdt_rollout_fast -s 1 -c 1 -f -o gpurun_out/prof_vdt_$T $SHORT > gpurun_out/ncu_full_vdt_$T.log 2>&1; echo "ncu vdt rc=$?"
python - gpurun_out/bench_vehicle_$T.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %s  frac %.3f" % (d["value"], d["ms_per_step"], e.get("value"), d["roofline"]["frac"]))
PY
