#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"shard_|probe_gather_inbox|pull_misses|insert_merge_kernel" -s 489 -c 200 --csv \
  --log-file gpurun_out/launches_r01d_c4.csv python bench.py --workload c4 --gpus 1 --steps 2 --warmup 3 > gpurun_out/ncu_bench_c4.log 2>&1
tail -3 gpurun_out/ncu_bench_c4.log | cut -c1-200
grep -c "gpu__time_duration" gpurun_out/launches_r01d_c4.csv
