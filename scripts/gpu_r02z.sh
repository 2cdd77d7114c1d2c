#!/bin/bash
# pass Z (2 GPUs): last check of the final bench.py — the N=2 line (server path) and the replica-process path
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 > gpurun_out/bench_r02z.json 2> gpurun_out/bench_r02z.err
echo "bench exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02z_ref.json 2> gpurun_out/bench_r02z_ref.err
echo "reference exit $?"; cut -c1-200 gpurun_out/bench_r02z_ref.json
HPSX_BENCH_REPLICA_PROCESSES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 \
    bench.py --gpus 2 --value-only --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep "\[bench\]\|value" | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02z.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified_rows','n_gpus')}, 'e2e', d['e2e']['value'], 'c4', (d.get('c4') or {}).get('value'))
PY
