#!/bin/bash
# round 2, pass I (2 GPUs): NVLink tier over CUDA IPC — parity test, then the default bench line at N=2 and, for
# comparison, the same without the tier (every miss over PCIe, the reference's replica behaviour)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2.txt 2>&1
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/pytest_sharded_2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sharded_2.log
tail -n 15 gpurun_out/pytest_sharded_2.log
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 $2 > gpurun_out/bench_r02i_$1.json 2> gpurun_out/bench_r02i_$1.err
  echo "bench $1 exit $?"
  tail -n 4 gpurun_out/bench_r02i_$1.err
}
run tier ""
run notier "--no-peer-tier --core-arms-only"
python - <<'PY'
import json
for nm in ('tier','notier'):
    try:
        d=json.loads(open(f'gpurun_out/bench_r02i_{nm}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(nm,'no line',e); continue
    print(nm,{k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus')})
    print('  e2e',{k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','verified_rows','note')})
    print('  tier',d.get('peer_tier'),'\n  nvlink',d.get('roofline_nvlink_tier'))
    print('  miss_path',d['miss_path'])
PY
