#!/bin/bash
# pass Q (2 GPUs): what is the fixed ~0.7 ms of the tier pull in the one-process-per-GPU bench?  static cache (no claims),
# gloo instead of NCCL
mkdir -p gpurun_out
: > gpurun_out/sweep_r02q.jsonl
run() {
  echo "{\"cfg\": \"$1 | $2\"}" >> gpurun_out/sweep_r02q.jsonl
  env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --value-only --steps 10 --warmup 3 --no-cpu-baseline $2 >> gpurun_out/sweep_r02q.jsonl 2>> gpurun_out/sweep_r02q.err
}
run "HPSX_BENCH_BACKEND=gloo" ""
run "HPSX_BENCH_BACKEND=nccl" "--static-cache"
run "HPSX_BENCH_BACKEND=nccl HPSX_TRACE=1" "--steps 3"
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r02q.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'cfg' in d: print(d['cfg']); continue
    print('   ', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','probe_ms','pull_ms','misses','tier_gbs')})
PY
grep "hpsx\]" gpurun_out/sweep_r02q.err | tail -n 12
