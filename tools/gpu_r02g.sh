#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02g
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.log
BENCH_ARGS="--workload vehicle" bash tools/run_variants.sh > gpurun_out/variants_$T.txt 2>&1; cat gpurun_out/variants_$T.txt
python bench.py --workload vehicle --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('main lib t128_u2 %.4g steps/s  %.3f ms' % (d['value'], d['ms_per_step']))" | tee -a gpurun_out/variants_$T.txt
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; echo "full rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-modules --chunk 2097152 > gpurun_out/bench_full_chunk2m_$T.json 2> gpurun_out/bench_full_chunk2m_$T.err; echo "full chunk 2M rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-modules --lanes 3 > gpurun_out/bench_full_lanes3_$T.json 2> gpurun_out/bench_full_lanes3_$T.err; echo "full lanes3 rc=$?"
for f in gpurun_out/bench_full*_$T.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f" % (d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"]))
if d.get("modules"): print({k:(v["value"], v["ms_per_launch"]) for k,v in d["modules"].items()})
PY
done
