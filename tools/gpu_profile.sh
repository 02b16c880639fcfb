#!/bin/bash
# Runs on the GPU box (under gpurun): bench line (+ reference arm), ncu launch list, ncu --set full of the solve kernel
# (default streaming form and, with "full", the record form), benches of the other BASELINE scenes.
# usage: tools/gpu_profile.sh <tag> [full]
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "${2:-}" = "full" ]; then
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --settle 30 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 30 -c 1 -f -o gpurun_out/k_solve_${TAG} \
    python bench.py --steps 2 --warmup 1 --settle 30 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
PHYX_SOLVE_PAIRS=2 ncu --set full --clock-control none --import-source on -k regex:k_solve_pairs2 -s 30 -c 1 -f -o gpurun_out/k_solve_pairs2_${TAG} \
    python bench.py --steps 2 --warmup 1 --settle 30 --no-cpu-baseline > gpurun_out/ncu_full_pairs2_${TAG}.log 2>&1; echo "ncu full (record form) rc=$?"
for sc in stack_100k islands_1m; do
python bench.py --scene $sc > gpurun_out/bench_${sc}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "bench $sc rc=$?"
done
fi
ls -la gpurun_out | tail -12
