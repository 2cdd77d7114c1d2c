#!/bin/bash
# round 2, pass A: stream-aliasing diagnosis, full GPU parity suite, smoke, default bench
mkdir -p gpurun_out
CUDA_MODULE_LOADING=EAGER ./tools/stream_alias_probe 24 > gpurun_out/stream_alias_eager.txt 2>&1
tail -n 2 gpurun_out/stream_alias_eager.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log
for i in $(seq 1 20); do timeout 300 python -m pytest tests/test_shard_group_gpu.py -q -x -k world2 2>&1 | tail -n 1; done > gpurun_out/world2_x20.log
sort gpurun_out/world2_x20.log | uniq -c
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
timeout 600 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
tail -n 5 gpurun_out/bench_r02a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02a.json'))
for k in ['value','ms_per_step','verified_rows']: print(k, d.get(k))
print('e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'))
print('roofline', {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_ms','share_of_step']})
print('host link', {k:d['roofline_host_link'][k] for k in ['achieved','frac','avg_ms_per_step']})
print('cache_hit', d['cache_hit'])
print('small_batch', d['small_batch'])
print('two', d['two_instances'])
print('miss_path', d['miss_path'])
PY
