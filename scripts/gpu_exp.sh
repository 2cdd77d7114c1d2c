#!/bin/bash
# 2-GPU experiment: NVLink peer-store methods + phase trace of the fused exchange
mkdir -p gpurun_out
timeout 120 ./tools/nvlink_probe > gpurun_out/nvlink_probe.txt 2>&1
cat gpurun_out/nvlink_probe.txt
HPSX_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --workload c4 --gpus 2 --exchange p2p --steps 6 --warmup 2 > gpurun_out/trace_c4.json 2> gpurun_out/trace_c4.err
grep "shard lookup rank 0" gpurun_out/trace_c4.err | tail -8
