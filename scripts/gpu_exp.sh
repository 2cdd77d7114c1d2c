#!/bin/bash
# scratch GPU pass: full parity suite + the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
tail -3 gpurun_out/bench_exp.err
cut -c1-300 gpurun_out/bench_exp.json
