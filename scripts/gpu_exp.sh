#!/bin/bash
# experiment: inbox kernel launch shape / store policy at world = 1
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload c4 --gpus 1 --steps 10 --warmup 3 --load-factor 0.4 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/exp_$tag.json')); r=d['roofline']
    print('$tag: step %.3f ms, kernel %.3f ms, %.0f GB/s (%.3f), misses %s' % (d['ms_per_step'], r['avg_launch_ms'], r['achieved'], r['frac'], d['exchange_stats_rank0']['misses']))
except Exception as e:
    print('$tag failed', e); print(open('gpurun_out/exp_$tag.err').read()[-800:])
PY
}
run c4 HPSX_INBOX_CTAS=4
run c0 HPSX_INBOX_CTAS=0
run c8 HPSX_INBOX_CTAS=8
run c4s HPSX_INBOX_CTAS=4 HPSX_INBOX_ST=1
run c0s HPSX_INBOX_CTAS=0 HPSX_INBOX_ST=1
