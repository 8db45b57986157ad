#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02c
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$T.log
B="python bench.py --no-cpu"
for V in "--ffsat 1" "--ffsat 0"; do
  N=$(echo $V | tr -d ' -')
  timeout 300 $B --workload vehicle --steps 5 --warmup 3 --no-e2e $V > gpurun_out/bench_vehicle_${N}_$T.json 2> gpurun_out/bench_vehicle_${N}_$T.err; echo "vehicle $V rc=$?"
done
timeout 600 $B --steps 3 --warmup 2 > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"
timeout 600 $B --steps 3 --warmup 2 --side-ctas 0 --no-modules > gpurun_out/bench_full_sidectas0_$T.json 2> gpurun_out/bench_full_sidectas0_$T.err; echo "full sidectas0 rc=$?"
timeout 600 $B --steps 3 --warmup 2 --ffsat 0 --no-modules --no-e2e > gpurun_out/bench_full_ffsat0_$T.json 2> gpurun_out/bench_full_ffsat0_$T.err; echo "full ffsat0 rc=$?"
python tools/tick_timeline.py --lanes 2 --side-ctas 1 > gpurun_out/timeline_$T.txt 2>&1
python tools/tick_timeline.py --lanes 2 --side-ctas 0 >> gpurun_out/timeline_$T.txt 2>&1
cat gpurun_out/timeline_$T.txt
SHORT="python bench.py --workload vehicle --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vdt_rollout_fast -s 1 -c 1 -f -o gpurun_out/prof_vdt_$T $SHORT > gpurun_out/ncu_full_vdt_$T.log 2>&1; echo "ncu vdt rc=$?"
SHORTA="python tools/bench_modules.py --only arm --reps 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adt_update -s 4 -c 1 -f -o gpurun_out/prof_arm_$T $SHORTA > gpurun_out/ncu_full_arm_$T.log 2>&1; echo "ncu arm rc=$?"
for f in gpurun_out/bench_*_$T.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("value %.4g  ms/step %.3f  e2e %s  frac %.3f  alone %s  clocks %s" % (d["value"], d["ms_per_step"], e.get("value"), d["roofline"]["frac"], d["roofline"].get("launch_ms_alone"), d["clocks"]["sm_mhz"]))
    if d.get("modules"): print({k:(v["value"], v["ms_per_launch"]) for k,v in d["modules"].items()})
except Exception as ex:
    print("unreadable:", ex)
PY
done
