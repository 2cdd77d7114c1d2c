#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_dcn_2.json 2> gpurun_out/bench_dcn_2.err
echo "exit $?"
python - <<PY
import json
txt=open('gpurun_out/bench_dcn_2.json').read()
print(repr(txt[:80]))
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('N=2 replicas: value %.1f M/s, step %.3f ms, e2e %.1f M/s, dense %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['dense_head'] and d['dense_head']['tflops']))
PY
tail -3 gpurun_out/bench_dcn_2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2.json 2> gpurun_out/bench_ref_2.err
echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_2.json
