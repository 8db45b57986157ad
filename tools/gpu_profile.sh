#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench, then the ncu launch list and one --set full
# capture of the rollout kernel for the same short command.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_${TAG}.err
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vdt_rollout -s 1 -c 1 \
    -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/
