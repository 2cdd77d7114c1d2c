#!/bin/bash
# pass N (1 GPU): repro of the tier-only table at scale (fill sync fixed), try-lock quad claims, duplication stress test
mkdir -p gpurun_out
timeout 300 python scripts/c4_repro.py > gpurun_out/c4_repro.log 2>&1
echo "repro exit $?"
tail -n 12 gpurun_out/c4_repro.log
timeout 300 python -m pytest tests/test_peer_tier_gpu.py -m gpu -x -q --timeout 90 > gpurun_out/pytest_tier.log 2>&1
echo "tier pytest exit $?"; tail -n 8 gpurun_out/pytest_tier.log
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_r02n.jsonl
for cfg in "--local-tier" "--variant v8"; do
  echo "{\"cfg\": \"$cfg\"}" >> gpurun_out/sweep_r02n.jsonl
  timeout 120 python bench.py --value-only --steps 20 --warmup 3 --no-cpu-baseline $cfg >> gpurun_out/sweep_r02n.jsonl 2>> gpurun_out/sweep_r02n.err
done
cut -c1-420 gpurun_out/sweep_r02n.jsonl
