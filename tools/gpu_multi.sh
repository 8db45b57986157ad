#!/bin/bash
# Multi-GPU evidence: the driver's torchrun command at N GPUs (+ the >= 2-device slice test when N == 2)
set -u
N=${1:-2}; T=${2:-r02}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_tick_gpu.py -q -k "two_gpu" 2>&1 | tail -3
fi
P=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_full_n${N}_$T.json 2> gpurun_out/bench_full_n${N}_$T.err; echo "full N=$N rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps 20 --warmup 5 --workload vehicle --no-cpu --no-modules > gpurun_out/bench_vehicle_n${N}_$T.json 2> gpurun_out/bench_vehicle_n${N}_$T.err; echo "vehicle N=$N rc=$?"
for f in gpurun_out/bench_full_n${N}_$T.json gpurun_out/bench_vehicle_n${N}_$T.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d.get("e2e") or {}
print("n_gpus %d value %.4g  ms/step %.3f  e2e %.4g  ratio %.3f  scaling %s" % (d["n_gpus"], d["value"], d["ms_per_step"], e.get("value"), e.get("value")/d["value"], d["scaling"]))
print(d["config"].get("parity_spot_check"))
PY
done
tail -3 gpurun_out/bench_full_n${N}_$T.err
