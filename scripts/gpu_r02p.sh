#!/bin/bash
# pass P (2 GPUs): is the IPC-mapped tier slow per byte or per call?  value-only arm at 90 % and ~20 % hit rate
mkdir -p gpurun_out
: > gpurun_out/sweep_r02p.jsonl
for cfg in "" "--hit 0.0 --prefill 2"; do
  echo "{\"cfg\": \"$cfg\"}" >> gpurun_out/sweep_r02p.jsonl
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --value-only --steps 10 --warmup 3 --no-cpu-baseline $cfg >> gpurun_out/sweep_r02p.jsonl 2>> gpurun_out/sweep_r02p.err
done
cut -c1-400 gpurun_out/sweep_r02p.jsonl
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version" gpurun_out/sweep_r02p.err | tail -n 5
