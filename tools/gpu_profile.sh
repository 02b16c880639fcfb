#!/bin/bash
# Runs on the GPU box (under gpurun): bench line (+ reference arm), ncu launch list, ncu --set full of the iteration kernel,
# per-pass trace of the strip kernel, benches of the other BASELINE scenes.
# usage: tools/gpu_profile.sh <tag> [full]
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "${2:-}" = "full" ]; then
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --settle 30 --no-cpu-baseline --no-parity > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu list rc=$?"
python tools/summarise_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}_summary.txt
ncu --set full --clock-control none --import-source on -k regex:k_solve_strips -s 33 -c 1 -f -o gpurun_out/k_solve_strips_${TAG} \
    python bench.py --steps 2 --warmup 3 --settle 30 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
PYTHONPATH=. python tools/strip_trace.py pyramid_1m 40 > gpurun_out/strip_trace_pyramid_1m_${TAG}.txt 2>&1
PYTHONPATH=. python tools/strip_trace.py stack_100k 40 > gpurun_out/strip_trace_stack_100k_${TAG}.txt 2>&1
for sc in stack_100k islands_1m pyramid_100k islands_128k; do
python bench.py --scene $sc --no-parity > gpurun_out/bench_${sc}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "bench $sc rc=$?"
done
fi
ls -la gpurun_out | tail -14
