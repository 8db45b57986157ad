set -u
mkdir -p gpurun_out
TAG=r01g
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"
python bench.py --workload full > gpurun_out/bench_full_${TAG}.json 2> gpurun_out/bench_full_${TAG}.err; echo "full rc=$?"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vdt_rollout -s 1 -c 1 -f -o gpurun_out/prof_vdt_${TAG} $SHORT > gpurun_out/ncu_full_vdt_${TAG}.log 2>&1
echo "ncu vdt rc=$?"
FULL="python bench.py --workload full --steps 1 --warmup 1 --no-e2e --no-cpu --total 2097152"
$FULL > gpurun_out/plain_full_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_full_${TAG}.csv $FULL > gpurun_out/ncu_list_full_${TAG}.log 2>&1
echo "ncu list full rc=$?"
