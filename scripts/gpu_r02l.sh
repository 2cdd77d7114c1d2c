#!/bin/bash
# round 2, pass L (2 GPUs): the default bench line at N=2 — replicas + NVLink tier, the one-server Triton arm (C++ instance
# threads) and the model-parallel configuration c4 (125 M rows per GPU, generated on the devices)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k peer_tier > gpurun_out/pytest_sharded_2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sharded_2.log
tail -n 5 gpurun_out/pytest_sharded_2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err
echo "bench exit $?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_r02l.err | tail -n 12
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02l.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus')})
    print('  e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','verified_rows','note','setup_s')})
    print('  nvlink',{k:v for k,v in (d.get('roofline_nvlink_tier') or {}).items() if k in ('achieved','avg_ms_per_step','frac')})
    c4=d.get('c4') or {}
    print('  c4',{k:v for k,v in c4.items() if k in ('value','ms_per_step','verified_rows','error','setup_s','arm_wall_s','rows')}, (c4.get('roofline_nvlink') or {}).get('achieved'))
except Exception as e:
    print('no line', e)
PY
