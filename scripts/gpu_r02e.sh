#!/bin/bash
# round 2, pass E: fused pull variants (grid, claim order); world-2 x20 with eager loading
mkdir -p gpurun_out
: > gpurun_out/sweep_r02e.jsonl
for cfg in "148 0" "222 0" "296 0" "444 0" "296 8" "148 8" "592 8"; do
  set -- $cfg
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --pull-ctas $1 --debug-flags $2 --no-cpu-baseline >> gpurun_out/sweep_r02e.jsonl 2>> gpurun_out/sweep_r02e.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02e.jsonl'):
    d=json.loads(l)
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','pull_ms','link_gbs','pull_ctas')})
PY
for i in $(seq 1 20); do timeout 300 python -m pytest tests/test_shard_group_gpu.py -q -x -s -k world2 > gpurun_out/w2_$i.log 2>&1 || { echo "run $i FAILED"; grep -E "rank [01], request|HpsxError|assert" gpurun_out/w2_$i.log | head -n 8; }; tail -n 1 gpurun_out/w2_$i.log; done > gpurun_out/world2_x20.log 2>&1
grep -c "2 passed" gpurun_out/world2_x20.log; grep -A6 FAILED gpurun_out/world2_x20.log | head -n 30
