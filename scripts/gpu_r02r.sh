#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_r02r.jsonl; : > gpurun_out/sweep_r02r.err
run() {
  echo "{\"cfg\": \"$1 | $2\"}" >> gpurun_out/sweep_r02r.jsonl
  echo "== $1 | $2" >> gpurun_out/sweep_r02r.err
  env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --value-only --steps 10 --warmup 3 --no-cpu-baseline $2 >> gpurun_out/sweep_r02r.jsonl 2>> gpurun_out/sweep_r02r.err
}
run "HPSX_BENCH_IDLE_RANKS=1" ""
run "A=1" ""
run "A=1" "--prefill 0"
run "A=1" "--hit 0.0 --prefill 0"
grep "^==\|\[bench\] rank" gpurun_out/sweep_r02r.err
