#!/bin/bash
# round 2, pass C: timeline of the binned pipeline under the experiment knobs
mkdir -p gpurun_out
: > gpurun_out/sweep_r02c.jsonl
for cfg in "1 296 0" "4 296 0" "4 296 1" "4 296 2" "4 296 3" "4 74 0" "2 296 0"; do
  set -- $cfg
  echo "== chunks $1 pull_ctas $2 debug $3" >> gpurun_out/sweep_r02c.err
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --chunks $1 --pull-ctas $2 --debug-flags $3 --no-cpu-baseline >> gpurun_out/sweep_r02c.jsonl 2>> gpurun_out/sweep_r02c.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02c.jsonl'):
    d=json.loads(l)
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','pull_ms','miss_phase_ms','link_gbs','chunks','pull_ctas')})
PY
grep -E "==|timeline" gpurun_out/sweep_r02c.err
