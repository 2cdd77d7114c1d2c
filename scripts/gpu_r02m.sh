#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/c4_repro.py > gpurun_out/c4_repro.log 2>&1
echo "repro exit $?"
cat gpurun_out/c4_repro.log | tail -n 20
timeout 600 python -m pytest tests/test_peer_tier_gpu.py tests/test_engine_gpu.py -m gpu -x -q > gpurun_out/pytest_tier.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_tier.log
: > gpurun_out/sweep_r02m.jsonl
for cfg in "--local-tier"; do
  echo "{\"cfg\": \"$cfg\"}" >> gpurun_out/sweep_r02m.jsonl
  timeout 300 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline $cfg >> gpurun_out/sweep_r02m.jsonl 2>> gpurun_out/sweep_r02m.err
done
cat gpurun_out/sweep_r02m.jsonl | cut -c1-400
