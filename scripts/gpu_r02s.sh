#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_r02s.jsonl; : > gpurun_out/sweep_r02s.err
run() {
  echo "== $1 | $2" >> gpurun_out/sweep_r02s.err
  env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --value-only --steps 10 --warmup 3 --no-cpu-baseline $2 >> gpurun_out/sweep_r02s.jsonl 2>> gpurun_out/sweep_r02s.err
}
run "A=1" ""
run "A=1" "--local-tier"
grep "^==\|\[bench\] rank" gpurun_out/sweep_r02s.err
