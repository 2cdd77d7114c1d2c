#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --workload c4 --gpus 8 --exchange p2p --steps 30 --warmup 5 > gpurun_out/bench_c4_p2p_8.json 2> gpurun_out/bench_c4_p2p_8.err
python - <<PY
import json
txt=open('gpurun_out/bench_c4_p2p_8.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=d['roofline']
print('p2p N=8: value %.2f G/s step %.3f ms, kernel %.3f ms, nvlink out %.0f GB/s, e2e %.3f ms (%.2f G/s)' % (d['value']/1e9, d['ms_per_step'], r['avg_launch_ms'], r['nvlink_out_gbs_per_gpu'], d['e2e']['ms_per_step'], d['e2e']['value']/1e9))
PY
tail -2 gpurun_out/bench_c4_p2p_8.err
