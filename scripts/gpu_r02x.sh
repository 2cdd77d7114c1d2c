#!/bin/bash
# pass X (1 GPU): final tree — whole GPU suite (TF32 head, Triton-level tier tests), smoke, default bench line, Zipf line
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
echo "bench exit $?"; tail -n 3 gpurun_out/bench_r02.err
timeout 300 python bench.py --value-only --zipf 1.05 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_zipf105.json 2> gpurun_out/bench_r02_zipf105.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified_rows')}, 'e2e', d['e2e']['value'], 'host_out', (d.get('e2e_host_output') or {}).get('value'))
print('link', d['roofline_host_link']['achieved'], d['roofline_host_link']['frac'], 'probe', d['roofline']['frac'])
for k in ('c1','c5','c3'):
    print(k, {kk:vv for kk,vv in (d.get(k) or {}).items() if kk in ('value','ms_per_step','hit_rate_measured','error','us_per_request')})
print('dense', {k:v for k,v in d['dense_head'].items() if k in ('tflops','frac_of_peak')}, d['dense_head'].get('tf32'), d['dense_head']['fused_bf16']['head_tflops'])
z=json.loads(open('gpurun_out/bench_r02_zipf105.json').read().strip().splitlines()[-1])
print('zipf', {k:z[k] for k in ('ms_per_step','probe_frac','pull_ms','misses','miss_duplicates','hit_rate_measured','unique_over_keys') if k in z})
PY
