#!/bin/bash
# GPU pass: parity tests, smoke, bench, host-side trace, ncu launch list + one full capture of the probe kernel.
# usage: bash scripts/gpu_profile.sh <round-tag>   (outputs under gpurun_out/)
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
HPSX_TRACE=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/trace_${TAG}.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_gather -s 3 -c 2 \
  -o gpurun_out/probe_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
grep hpsx gpurun_out/trace_${TAG}.log | head -12
