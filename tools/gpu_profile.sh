#!/bin/bash
# Runs on the GPU box (under gpurun): bench line (+ reference arm), ncu launch list, ncu --set full of k_solve.
# usage: tools/gpu_profile.sh <tag> [full]
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "${2:-}" = "full" ]; then
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --settle 6 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 8 -c 1 -f -o gpurun_out/k_solve_${TAG} \
    python bench.py --steps 2 --warmup 1 --settle 6 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out | tail -8
