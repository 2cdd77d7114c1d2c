#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_group_gpu.py -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_sharded.sh 2 > gpurun_out/sharded2.log 2>&1
tail -3 gpurun_out/pytest_sharded_2.log
python - <<PY
import json
for ex in ["p2p","nccl"]:
    try:
        txt=open(f'gpurun_out/bench_c4_{ex}_2.json').read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=d['roofline']
        print('%s: value %.2f G/s step %.3f ms, kernel %.3f ms, nvlink out %.0f GB/s, e2e %.3f ms' % (ex, d['value']/1e9, d['ms_per_step'], r['avg_launch_ms'], r['nvlink_out_gbs_per_gpu'], d['e2e']['ms_per_step']))
    except Exception as e:
        print(ex, "failed", e); print(open(f'gpurun_out/bench_c4_{ex}_2.err').read()[-1500:])
PY
