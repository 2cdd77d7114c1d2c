#!/bin/bash
# quick GPU pass: parity tests, default bench, single-GPU run of the model-parallel workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
timeout 600 python bench.py --workload c4 --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_c4_p2p_1.json 2> gpurun_out/bench_c4_p2p_1.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_c4_p2p_1.err; cat gpurun_out/bench_c4_p2p_1.json
