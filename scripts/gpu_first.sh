#!/bin/bash
# first GPU pass: parity tests, smoke, short bench for both probe variants, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --variant ldg > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err
timeout 900 python bench.py --steps 10 --warmup 3 --variant tma --no-cpu-baseline > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_ldg.json gpurun_out/bench_tma.json; tail -5 gpurun_out/bench_ldg.err
