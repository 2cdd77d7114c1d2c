#!/bin/bash
# pass U (8 GPUs): the default bench line at N=8 — one server process, 8 caches + NVLink tier, Triton one-server arm, c4 at
# 1 B rows (125 M per GPU)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err
echo "bench exit $?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_r02w.err | tail -n 8
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02w.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus','wall_ms_per_step')})
    print('  e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','verified_rows','note','setup_s')})
    print('  e2e_session',{k:v for k,v in d['e2e_session'].items() if k in ('value','ms_per_step')})
    print('  nvlink',{k:v for k,v in (d.get('roofline_nvlink_tier') or {}).items() if k in ('achieved','avg_ms_per_step','frac')})
    print('  roofline',{k:v for k,v in d['roofline'].items() if k in ('achieved','frac','avg_launch_ms')}, d['miss_path'])
    c4=d.get('c4') or {}
    print('  c4',{k:v for k,v in c4.items() if k in ('value','ms_per_step','verified_rows','error','setup_s','arm_wall_s','rows')}, (c4.get('roofline_nvlink') or {}).get('achieved'))
    print('  clocks', d['clocks'].get('sm_mhz'), [c.get('sm_mhz') for c in d['clocks'].get('per_gpu', [])])
except Exception as e:
    print('no line', e)
PY
