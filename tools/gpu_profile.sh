#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, ncu --set full of the dominant kernel.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --settle 1 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -f -o gpurun_out/k_solve_${TAG} \
    python bench.py --steps 2 --warmup 1 --settle 1 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
