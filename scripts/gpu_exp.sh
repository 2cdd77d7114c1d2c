#!/bin/bash
mkdir -p gpurun_out
HPSX_TRACE=1 HPS_TRACE=1 HPSX_PIPE_CHUNKS=0 BENCH_NO_SAMPLER=1 timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err
grep -n "\[hps\] execute\|host keys" gpurun_out/ab_new.err | sed -n '1,400p' | awk 'NR%1==0' | tail -40
