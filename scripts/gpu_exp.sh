#!/bin/bash
# configs[2] shape: 100M-row table, 50 % measured hit rate (host-miss path stressed)
mkdir -p gpurun_out
free -g | tee gpurun_out/free.txt
AVAIL=$(free -g | awk '/^Mem:/ {print $7}')
if [ "$AVAIL" -lt 120 ]; then echo "not enough host memory ($AVAIL GB): skipping"; exit 0; fi
timeout 900 python bench.py --rows 100000000 --hit 0.42 --prefill 80 --steps 8 --warmup 3 --no-cpu-baseline --core-arms-only --skip-triton-arm \
  > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "exit $?"
tail -3 gpurun_out/bench_c3.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_c3.json').read().splitlines() if l.startswith('{')][-1])
print('C3: value %.1f M/s step %.3f ms hit %.4f | e2e(session) %.1f M/s | probe frac %.3f | link %.1f GB/s frac %.3f | setup %.1f s' % (d['value']/1e6, d['ms_per_step'], d['config']['hit_rate_measured'], d['e2e_session']['value']/1e6, d['roofline']['frac'], d['roofline_host_link']['achieved'], d['roofline_host_link']['frac'], d['config']['setup_s']))
PY
