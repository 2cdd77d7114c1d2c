#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "v8 or resident" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/exp_$tag.json').read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['cache_hit']
    print('$tag: step %.3f ms, probe %.3f ms %.0f GB/s (%.3f) | all-hit %.3f ms %.0f GB/s (%.3f) | e2e %.3f ms' % (d['ms_per_step'], r['avg_launch_ms'], r['achieved'], r['frac'], c['kernel_ms'], c['hbm_gbs'], c['frac_of_peak'], d['e2e']['ms_per_step']))
except Exception as e:
    print('$tag failed', e); print(open('gpurun_out/exp_$tag.err').read()[-600:])
PY
}
run ldg HPSX_PROBE=ldg
run v8u4 HPSX_PROBE=v8 HPSX_V8_UNROLL=4
run v8u2 HPSX_PROBE=v8 HPSX_V8_UNROLL=2
run v8u8 HPSX_PROBE=v8 HPSX_V8_UNROLL=8
