#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json'))
l=d['roofline_host_link']
print('$tag: value %.1fM ms %.3f hit %.4f' % (d['value']/1e6, d['ms_per_step'], d['config']['hit_rate_measured']), 'link %.1f GB/s of %.1f, pull %.3f ms' % (l['achieved'], l['peak'], l['avg_ms_per_step']), 'e2e %.1fM' % (d['e2e']['value']/1e6))" || tail -5 gpurun_out/bench_$tag.err
}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run sort1 HPSX_PULL_SORT=1
run sort0 HPSX_PULL_SORT=0
run sort1c4 HPSX_PULL_SORT=1 HPSX_PULL_CTAS=4
run sort1c16 HPSX_PULL_SORT=1 HPSX_PULL_CTAS=16
