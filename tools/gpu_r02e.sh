#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02e
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$T.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"; tail -3 gpurun_out/bench_full_$T.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
for f in gpurun_out/bench_full_$T.json gpurun_out/bench_ref_$T.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %s  cpu %s" % (d["value"], d["ms_per_step"], e.get("value"), (d.get("cpu_baseline") or {}).get("value")))
if d.get("modules"): print({k:(v["value"], v["ms_per_launch"]) for k,v in d["modules"].items()})
print(d["config"].get("parity_spot_check"))
PY
done
