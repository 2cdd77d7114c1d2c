#!/bin/bash
# round 2, pass J (1 GPU): restructured slot claim (locks released before the row arrives) + 4 rows in flight per warp for
# tier pulls; probe-kernel experiments (L2 row prefetch, deeper unroll); local tier = pull from local HBM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_r02j.jsonl
for cfg in "--variant v8" "--variant v8p1" "--variant v8p2" "--variant v8u8" "--variant v8 --local-tier" "--variant v8 --local-tier --pull-ctas 1184"; do
  echo "{\"cfg\": \"$cfg\"}" >> gpurun_out/sweep_r02j.jsonl
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline $cfg >> gpurun_out/sweep_r02j.jsonl 2>> gpurun_out/sweep_r02j.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02j.jsonl'):
    d=json.loads(l)
    if 'cfg' in d: print(d['cfg']); continue
    print('   ', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','probe_frac','pull_ms','link_gbs','tier_gbs','all_hit_kernel_ms','all_hit_frac')})
PY
tail -n 3 gpurun_out/sweep_r02j.err
