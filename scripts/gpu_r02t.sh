#!/bin/bash
# pass T (2 GPUs): the N > 1 bench line with ONE server process driving all GPUs (value, session e2e, Triton one-server arm, c4)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 > gpurun_out/bench_r02t.json 2> gpurun_out/bench_r02t.err
echo "bench exit $?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_r02t.err | tail -n 8
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02t.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus','wall_ms_per_step')})
    print('  e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','verified_rows','note','setup_s')})
    print('  e2e_session',{k:v for k,v in d['e2e_session'].items() if k in ('value','ms_per_step')})
    print('  nvlink',{k:v for k,v in (d.get('roofline_nvlink_tier') or {}).items() if k in ('achieved','avg_ms_per_step','frac')})
    print('  roofline',{k:v for k,v in d['roofline'].items() if k in ('achieved','frac','avg_launch_ms')}, d['miss_path'])
    c4=d.get('c4') or {}
    print('  c4',{k:v for k,v in c4.items() if k in ('value','ms_per_step','verified_rows','error','setup_s','arm_wall_s','rows')}, (c4.get('roofline_nvlink') or {}).get('achieved'))
except Exception as e:
    print('no line', e)
PY
