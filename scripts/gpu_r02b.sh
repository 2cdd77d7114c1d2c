#!/bin/bash
# round 2, pass B: shape of the binned pipeline (chunks x pull grid), world-2 repeat with failure details
mkdir -p gpurun_out
: > gpurun_out/sweep_r02b.jsonl
for cfg in "1 296" "1 1184" "4 296" "4 148" "4 74" "4 37" "2 148" "8 148" "4 592"; do
  set -- $cfg
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --chunks $1 --pull-ctas $2 --no-cpu-baseline >> gpurun_out/sweep_r02b.jsonl 2>> gpurun_out/sweep_r02b.err
done
cat gpurun_out/sweep_r02b.jsonl
for i in $(seq 1 20); do timeout 300 python -m pytest tests/test_shard_group_gpu.py -q -x -k world2 > gpurun_out/w2_$i.log 2>&1 || { echo "run $i FAILED"; grep -E "HpsxError|Error|assert" gpurun_out/w2_$i.log | head -n 8; }; tail -n 1 gpurun_out/w2_$i.log; done > gpurun_out/world2_x20.log 2>&1
sort gpurun_out/world2_x20.log | uniq -c | sort -rn | head -n 12
