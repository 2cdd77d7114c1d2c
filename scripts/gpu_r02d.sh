#!/bin/bash
# round 2, pass D: fused binned pull; parity suite; bench; world-2 repeat
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
for i in $(seq 1 20); do timeout 300 python -m pytest tests/test_shard_group_gpu.py -q -x -k world2 > gpurun_out/w2_$i.log 2>&1 || { echo "run $i FAILED"; grep -E "HpsxError|Error|assert" gpurun_out/w2_$i.log | head -n 8; }; tail -n 1 gpurun_out/w2_$i.log; done > gpurun_out/world2_x20.log 2>&1
grep -c "2 passed" gpurun_out/world2_x20.log; grep -A4 FAILED gpurun_out/world2_x20.log | head -n 20
timeout 300 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench_r02d.err
timeout 300 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline --zipf 1.05 2>> gpurun_out/bench_r02d.err
timeout 600 python bench.py > gpurun_out/bench_r02d.json 2>> gpurun_out/bench_r02d.err
tail -n 5 gpurun_out/bench_r02d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02d.json'))
for k in ['value','ms_per_step','verified_rows']: print(k, d.get(k))
print('e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'))
print('e2e_session', d['e2e_session'].get('value'), d['e2e_session'].get('ms_per_step'))
print('roofline', {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_ms','share_of_step']})
print('host link', {k:d['roofline_host_link'][k] for k in ['achieved','frac','avg_ms_per_step']})
print('cache_hit', d['cache_hit'])
print('small_batch', d['small_batch'])
print('two', d['two_instances'])
PY
