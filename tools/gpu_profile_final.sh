#!/bin/bash
# Runs on the GPU box (via gpurun): plain runs first, then the ncu launch list of the same short bench command and one
# --set full capture per kernel of interest.  Outputs land in gpurun_out/ (summarised into profiles/ on the CPU box).
set -u
mkdir -p gpurun_out
TAG=${1:-r01f}
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python tools/bench_modules.py > gpurun_out/bench_modules_${TAG}.jsonl 2> gpurun_out/bench_modules_${TAG}.err; echo "modules rc=$?"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vdt_rollout -s 1 -c 1 \
    -f -o gpurun_out/prof_vdt_${TAG} $SHORT > gpurun_out/ncu_full_vdt_${TAG}.log 2>&1
echo "ncu vdt rc=$?"
for K in imu:imt_update_kernel wire:imt_feed_bytes guard:rmt_guard; do
  W=${K%%:*}; R=${K##*:}
  python tools/bench_modules.py --only $W --reps 2 > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$R -s 4 -c 1 \
      -f -o gpurun_out/prof_${W}_${TAG} python tools/bench_modules.py --only $W --reps 2 > gpurun_out/ncu_full_${W}_${TAG}.log 2>&1
  echo "ncu $W rc=$?"
done
ls -la gpurun_out/ | tail -20
