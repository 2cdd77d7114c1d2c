#!/bin/bash
mkdir -p gpurun_out
for c in 4 8 16; do
  HPSX_PULL_CTAS=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pull$c.json 2> gpurun_out/bench_pull$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_pull$c.json'))
print('pull ctas $c: value %.1fM ms %.3f hit %.4f' % (d['value']/1e6, d['ms_per_step'], d['config']['hit_rate_measured']), 'probe', round(d['roofline']['frac'],3), 'ceiling', round(d['roofline']['random_gather_ceiling_gbs'],1), 'link', d['roofline_host_link'], 'e2e %.1fM' % (d['e2e']['value']/1e6))"
done
