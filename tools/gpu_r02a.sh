#!/bin/bash
# Round 2, first GPU pass: tests, short benches of both workloads with scheduling variants, one ncu capture.
set -u
mkdir -p gpurun_out
T=r02a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi_$T.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$T.log
B="python bench.py --no-cpu"
timeout 300 $B --workload vehicle --steps 5 --warmup 3 > gpurun_out/bench_vehicle_$T.json 2> gpurun_out/bench_vehicle_$T.err; echo "vehicle rc=$?"
timeout 300 $B --workload vehicle --steps 5 --warmup 3 --occupancy 3 --no-e2e > gpurun_out/bench_vehicle_occ3_$T.json 2> gpurun_out/bench_vehicle_occ3_$T.err; echo "vehicle occ3 rc=$?"
timeout 600 $B --steps 3 --warmup 2 > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"
for V in "--side-ctas 0" "--side-ctas 2" "--lanes 1" "--lanes 3" "--occupancy 3"; do
  N=$(echo $V | tr -d ' -')
  timeout 600 $B --steps 3 --warmup 2 --no-e2e --no-modules $V > gpurun_out/bench_full_${N}_$T.json 2> gpurun_out/bench_full_${N}_$T.err; echo "full $V rc=$?"
done
SHORT="python bench.py --workload vehicle --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vdt_rollout_fast -s 1 -c 1 -f -o gpurun_out/prof_vdt_$T $SHORT > gpurun_out/ncu_full_vdt_$T.log 2>&1; echo "ncu vdt rc=$?"
SHORTF="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-modules --total 2097152"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_full_$T.csv $SHORTF > gpurun_out/ncu_list_full_$T.log 2>&1; echo "ncu list rc=$?"
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
ls -la gpurun_out | tail -30
