#!/bin/bash
# quick GPU pass: tests + short traced bench in both miss-path modes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --miss-path direct > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err
HPSX_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --miss-path staged > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_direct.err; cat gpurun_out/bench_direct.json gpurun_out/bench_staged.json
