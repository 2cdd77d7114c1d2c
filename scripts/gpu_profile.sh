#!/bin/bash
# GPU pass: parity tests, smoke, bench (both miss paths), ncu launch list + full captures of the two hot kernels.
# usage: bash scripts/gpu_profile.sh <round-tag>   (outputs under gpurun_out/)
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 900 python bench.py --miss-path staged --no-cpu-baseline > gpurun_out/bench_${TAG}_staged.json 2> gpurun_out/bench_${TAG}_staged.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_gather|pull_misses" -s 10 -c 4 \
  -o gpurun_out/hot_${TAG} -f python bench.py --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
