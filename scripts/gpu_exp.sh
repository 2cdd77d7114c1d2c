#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --core-arms-only --skip-triton-arm --steps 20 > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  grep "L2 persist" gpurun_out/exp_$tag.err | head -1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/exp_$tag.json').read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['cache_hit']
print('$tag: step %.3f ms probe %.3f ms (%.3f) | all-hit %.3f ms (%.3f)' % (d['ms_per_step'], r['avg_launch_ms'], r['frac'], c['kernel_ms'], c['frac_of_peak']))
PY
}
run base HPSX_L2_PERSIST=0
run persist HPSX_L2_PERSIST=1
run ldg_persist HPSX_L2_PERSIST=1 HPSX_PROBE=ldg
