#!/bin/bash
# quick GPU pass: tests + link probe + short traced bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
./tools/pcie_probe > gpurun_out/pcie_probe.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nvidia-smi -q | grep -i -A3 "link\|pcie" | head -40 >> gpurun_out/lscpu.txt
HPSX_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --load-factor 0.5 > gpurun_out/bench_quick.json 2> gpurun_out/trace_quick.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/pcie_probe.txt; grep hpsx gpurun_out/trace_quick.log | tail -6; cat gpurun_out/bench_quick.json
