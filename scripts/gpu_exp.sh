#!/bin/bash
mkdir -p gpurun_out
for SPLIT in 1 0; do
HPSX_SPLIT_LOCK=$SPLIT timeout 600 python bench.py --no-cpu-baseline --steps 20 --skip-triton-arm > gpurun_out/bench_exp_$SPLIT.json 2> gpurun_out/bench_exp_$SPLIT.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_exp_$SPLIT.json').read().splitlines() if l.startswith('{')][-1])
t=d['two_instances']; print('split=$SPLIT', 'agg %.1f M/s (x%.3f)' % (t['vectors_per_s']/1e6, t['vs_one_instance']), t['small_request_ms_beside_large_stream'])
print('   fused', d['dense_head']['fused_bf16']['gather_with_mirror_kernel_ms'], d['dense_head']['fused_bf16']['head_frac_of_peak'])
PY
tail -2 gpurun_out/bench_exp_$SPLIT.err
done
