#!/bin/bash
# round 2, pass H: NVLink tier logic on one GPU (two ranks on one device) + the whole GPU suite + the full bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_tier_gpu.py -m gpu -x -q > gpurun_out/pytest_tier.log 2>&1
echo "tier pytest exit $?" >> gpurun_out/pytest_tier.log
tail -n 25 gpurun_out/pytest_tier.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
echo "bench exit $?"
tail -n 5 gpurun_out/bench_r02h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02h.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified_rows')})
print('e2e',d['e2e']['value'],'host_out',d['e2e_host_output']['value'])
for k in ('c1','c5','c3'):
    print(k, {kk:vv for kk,vv in (d.get(k) or {}).items() if kk in ('value','ms_per_step','hit_rate_measured','error','us_per_request')})
PY
