#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/exp_$tag.json').read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['cache_hit']; sb=d['small_batch']
    print('$tag: step %.3f ms | e2e %.3f ms | small: batched %.1f M/s (%.3f ms/req) one-by-one %.1f M/s (%.3f ms/req)' % (d['ms_per_step'], d['e2e']['ms_per_step'], sb['batched_vectors_per_s']/1e6, sb['batched_ms_per_request'], sb['one_by_one_vectors_per_s']/1e6, sb['one_by_one_ms_per_request']))
except Exception as e:
    print('$tag failed', e); print(open('gpurun_out/exp_$tag.err').read()[-900:])
PY
}
run min16k HPSX_PULL_SORT_MIN=16384
run min0 HPSX_PULL_SORT_MIN=0
run min4k HPSX_PULL_SORT_MIN=4096
