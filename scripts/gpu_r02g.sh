#!/bin/bash
# round 2, pass G: pull window x pull grid on the C2 (10M rows) and C3 (100M rows, 50 % hit) shapes
mkdir -p gpurun_out
: > gpurun_out/sweep_r02g.jsonl
for cfg in "16 370" "16 148" "16 222"; do
  set -- $cfg
  echo "{\"cfg\": \"c2 window $1 ctas $2\"}" >> gpurun_out/sweep_r02g.jsonl
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --window-mb $1 --pull-ctas $2 --no-cpu-baseline >> gpurun_out/sweep_r02g.jsonl 2>> gpurun_out/sweep_r02g.err
done
for cfg in "16 148" "16 222" "16 370" "64 148" "8 148"; do
  set -- $cfg
  echo "{\"cfg\": \"c3 window $1 ctas $2\"}" >> gpurun_out/sweep_r02g.jsonl
  timeout 600 python bench.py --value-only --rows 100000000 --hit 0.42 --prefill 26 --steps 6 --warmup 2 --window-mb $1 --pull-ctas $2 --no-cpu-baseline >> gpurun_out/sweep_r02g.jsonl 2>> gpurun_out/sweep_r02g.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02g.jsonl'):
    d=json.loads(l)
    if 'cfg' in d: print(d['cfg']); continue
    print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','pull_ms','link_gbs','misses','all_hit_kernel_ms')})
PY
tail -n 3 gpurun_out/sweep_r02g.err
