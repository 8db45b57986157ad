#!/bin/bash
# Final evidence of the round (1 GPU, ~5 GPU-minutes): tests, the driver's bench commands, launch list, ncu of the dominant kernel.
set -u
mkdir -p gpurun_out
T=${1:-r05}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$T.log
timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"
timeout 900 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
timeout 900 python3 bench.py --workload vehicle --steps 20 --warmup 5 --no-cpu --no-modules > gpurun_out/bench_vehicle_$T.json 2> gpurun_out/bench_vehicle_$T.err; echo "vehicle rc=$?"
for f in gpurun_out/bench_full_$T.json gpurun_out/bench_vehicle_$T.json gpurun_out/bench_ref_$T.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f" % (d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"]))
if d.get("modules"): print({k:(v["value"], v["ms_per_launch"]) for k,v in d["modules"].items()})
PY
done
timeout 600 python tools/bench_modules.py > gpurun_out/modules_$T.jsonl 2> gpurun_out/modules_$T.err; echo "modules rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-modules"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full_$T.csv $SHORT > gpurun_out/ncu_list_$T.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vdt_rollout_fast -s 20 -c 1 -f -o gpurun_out/prof_vdt_$T $SHORT > gpurun_out/ncu_full_vdt_$T.log 2>&1; echo "ncu vdt rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_imu_samples -s 2 -c 1 -f -o gpurun_out/prof_streamimu_$T $SHORT > gpurun_out/ncu_full_streamimu_$T.log 2>&1; echo "ncu stream rc=$?"
