#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for M in 1 0; do
HPSX_BATCH_MERGE=$M timeout 600 python bench.py --no-cpu-baseline --skip-triton-arm --steps 10 > gpurun_out/bench_exp_$M.json 2> gpurun_out/bench_exp_$M.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_exp_$M.json').read().splitlines() if l.startswith('{')][-1])
sb=d['small_batch']; print('merge=$M batched %.3f ms/req (%.1f M/s) one-by-one %.3f ms/req' % (sb['batched_ms_per_request'], sb['batched_vectors_per_s']/1e6, sb['one_by_one_ms_per_request']))
PY
done
