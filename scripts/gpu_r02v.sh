#!/bin/bash
# pass V (2 GPUs): tier tests after the gather-kernel restructure, then the N=2 line (c4 with two instances per GPU)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_peer_tier_gpu.py -m gpu -x -q --timeout 90 > gpurun_out/pytest_tier.log 2>&1
echo "tier pytest exit $?"; tail -n 4 gpurun_out/pytest_tier.log
timeout 300 python scripts/c4_repro.py 2>&1 | tail -n 3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err
echo "bench exit $?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_r02v.err | tail -n 8
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02v.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus','wall_ms_per_step')})
    print('  e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','verified_rows','note','setup_s')})
    c4=d.get('c4') or {}
    print('  c4',{k:v for k,v in c4.items() if k in ('value','ms_per_step','verified_rows','error','setup_s','arm_wall_s','rows','instances_per_gpu')}, (c4.get('roofline_nvlink') or {}).get('achieved'))
except Exception as e:
    print('no line', e)
PY
