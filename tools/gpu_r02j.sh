#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02j
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"
python - gpurun_out/bench_full_$T.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f" % (d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"]))
if d.get("modules"): print({k:(v["value"], v["ms_per_launch"]) for k,v in d["modules"].items()})
PY
SHORTA="python tools/bench_modules.py --only arm --reps 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adt_update -s 4 -c 1 -f -o gpurun_out/prof_arm_$T $SHORTA > gpurun_out/ncu_full_arm_$T.log 2>&1; echo "ncu arm rc=$?"
SHORTI="python tools/bench_modules.py --only imu --reps 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:imt_update -s 8 -c 1 -f -o gpurun_out/prof_imu_$T $SHORTI > gpurun_out/ncu_full_imu_$T.log 2>&1; echo "ncu imu rc=$?"
