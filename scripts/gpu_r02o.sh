#!/bin/bash
# pass O (2 GPUs): quad claims before the row loads; the default bench line at N=2 incl. the one-server arm and c4
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_peer_tier_gpu.py -m gpu -x -q --timeout 90 > gpurun_out/pytest_tier.log 2>&1
echo "tier pytest exit $?"; tail -n 4 gpurun_out/pytest_tier.log
timeout 120 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline --local-tier 2>> gpurun_out/sweep_r02o.err | cut -c1-330
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err
echo "bench exit $?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_r02o.err | tail -n 8
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02o.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus')})
    print('  e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','verified_rows','note','setup_s')})
    print('  nvlink',{k:v for k,v in (d.get('roofline_nvlink_tier') or {}).items() if k in ('achieved','avg_ms_per_step','frac')})
    c4=d.get('c4') or {}
    print('  c4',{k:v for k,v in c4.items() if k in ('value','ms_per_step','verified_rows','error','setup_s','arm_wall_s','rows')}, (c4.get('roofline_nvlink') or {}).get('achieved'))
except Exception as e:
    print('no line', e)
PY
