#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/exp_$tag.json').read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['cache_hit']; sb=d['small_batch']
    print('$tag: step %.3f ms probe %.3f (%.3f) | e2e %.3f ms session %.3f | small: batched %.1f M/s (%.3f ms/req) one-by-one %.1f M/s (%.3f ms/req)' % (d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['e2e']['ms_per_step'], d['e2e_session']['ms_per_step'], sb['batched_vectors_per_s']/1e6, sb['batched_ms_per_request'], sb['one_by_one_vectors_per_s']/1e6, sb['one_by_one_ms_per_request']))
except Exception as e:
    print('$tag failed', e); print(open('gpurun_out/exp_$tag.err').read()[-900:])
PY
}
run c4 HPSX_COPY_CHUNKS=4
run c0 HPSX_COPY_CHUNKS=0
run c8 HPSX_COPY_CHUNKS=8
run c2 HPSX_COPY_CHUNKS=2
