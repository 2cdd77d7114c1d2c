#!/bin/bash
# pass Y (2 GPUs): one replica process per GPU over CUDA IPC — does the ~0.7 ms the pull kernel pays per launch there sit in
# the slot claims (atomics / MEMBAR beside IPC-mapped memory)?  Split form = pull without insert + separate insert pass.
mkdir -p gpurun_out
: > gpurun_out/sweep_r02y.err
run() {
  echo "== $1" >> gpurun_out/sweep_r02y.err
  env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --value-only --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>> gpurun_out/sweep_r02y.err
}
run "HPSX_BENCH_REPLICA_PROCESSES=1"
run "HPSX_BENCH_REPLICA_PROCESSES=1 HPSX_BENCH_SPLIT_FORM=1 HPSX_TRACE=0"
grep "^==\|\[bench\] rank" gpurun_out/sweep_r02y.err
