#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02h
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"
timeout 300 python tools/bench_modules.py --reps 5 > gpurun_out/modules_$T.jsonl 2> gpurun_out/modules_$T.err; echo "modules rc=$?"
python tools/tick_timeline.py --lanes 2 --side-ctas 1 > gpurun_out/timeline_$T.txt 2>&1; tail -8 gpurun_out/timeline_$T.txt
for f in gpurun_out/bench_full_$T.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f" % (d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"]))
if d.get("modules"): print({k:(v["value"], v["ms_per_launch"]) for k,v in d["modules"].items()})
PY
done
python - <<'PY'
import json
for l in open("gpurun_out/modules_r02h.jsonl"):
    d=json.loads(l); k=[x for x in d if x.endswith("_per_s")][0]
    print(d["kernel"], d["workload"][:60], "%.4g" % d[k], "%.3f ms" % d["ms_per_launch"], d["roofline"].get("frac"))
PY
