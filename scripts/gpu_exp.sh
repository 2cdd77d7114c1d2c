#!/bin/bash
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_engine_gpu.py tests/test_shard_group_gpu.py -m gpu -q -x \
  -k "(all_resident and tma and 4001) or route_and_scatter or (world2 and True) or (world1 and 128 and True)" \
  > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|racecheck exit|hazard" gpurun_out/sanitizer_racecheck.log | head -20
