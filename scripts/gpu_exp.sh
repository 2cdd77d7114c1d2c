#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/exp_mlp.json 2> gpurun_out/exp_mlp.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/exp_mlp.json').read().splitlines() if l.startswith('{')][-1])
print('value %.1f step %.3f probe %.3f (%.3f) e2e %.3f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step']))
print(json.dumps(d['dense_head'], indent=1))
PY
tail -3 gpurun_out/exp_mlp.err
