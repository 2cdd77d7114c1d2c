#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dense_mlp_gpu.py -m gpu -q -x > gpurun_out/pytest_mlp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mlp.log
tail -5 gpurun_out/pytest_mlp.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/exp_mlp.json 2> gpurun_out/exp_mlp.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/exp_mlp.json').read().splitlines() if l.startswith('{')][-1])
print(d['dense_head'])
PY
tail -3 gpurun_out/exp_mlp.err
