#!/bin/bash
# round 2, pass K (1 GPU): per-warp insert counters; tables that live in the tier only (device-generated shards, tier
# gather kernel); what the LRU touch costs the probe kernel (static cache)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_tier_gpu.py -m gpu -x -q > gpurun_out/pytest_tier.log 2>&1
echo "tier pytest exit $?" >> gpurun_out/pytest_tier.log
tail -n 30 gpurun_out/pytest_tier.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_r02k.jsonl
for cfg in "--variant v8" "--static-cache" "--local-tier" "--local-tier --pull-ctas 296" "--local-tier --pull-ctas 1184"; do
  echo "{\"cfg\": \"$cfg\"}" >> gpurun_out/sweep_r02k.jsonl
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline $cfg >> gpurun_out/sweep_r02k.jsonl 2>> gpurun_out/sweep_r02k.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02k.jsonl'):
    d=json.loads(l)
    if 'cfg' in d: print(d['cfg']); continue
    print('   ', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','probe_frac','pull_ms','link_gbs','tier_gbs','all_hit_kernel_ms','all_hit_frac')})
PY
tail -n 3 gpurun_out/sweep_r02k.err
