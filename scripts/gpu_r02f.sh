#!/bin/bash
# round 2, pass F: pull grid (fused), world-2 x20 after the GC fix, full default bench with the extra configurations
mkdir -p gpurun_out
: > gpurun_out/sweep_r02f.jsonl
for c in 0; do
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --pull-ctas $c --no-cpu-baseline >> gpurun_out/sweep_r02f.jsonl 2>> gpurun_out/sweep_r02f.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02f.jsonl'):
    d=json.loads(l)
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','pull_ms','link_gbs','pull_ctas')})
PY
for i in $(seq 1 20); do timeout 300 python -m pytest tests/test_shard_group_gpu.py -q -x -s -k world2 > gpurun_out/w2_$i.log 2>&1 || { echo "run $i FAILED"; grep -E "rank [01], request|Thread|File|line" gpurun_out/w2_$i.log | head -n 60; }; tail -n 1 gpurun_out/w2_$i.log; done > gpurun_out/world2_x20.log 2>&1
grep -c "2 passed" gpurun_out/world2_x20.log; grep -A60 FAILED gpurun_out/world2_x20.log | head -n 80
timeout 1500 python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
tail -n 5 gpurun_out/bench_r02f.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02f.json'))
for k in ['value','ms_per_step','verified_rows']: print(k, d.get(k))
print('e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'))
print('e2e_host_output', d['e2e_host_output'])
for k in ('c1','c5','c3'):
    print(k, json.dumps(d.get(k))[:1500])
PY
