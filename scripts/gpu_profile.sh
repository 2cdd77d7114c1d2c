#!/bin/bash
# GPU pass: parity tests, smoke, bench (both miss paths + reference arm), ncu launch list + full captures of the hot
# kernels, and the single-GPU form of the model-parallel workload.  usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 900 python bench.py --miss-path staged --no-cpu-baseline > gpurun_out/bench_${TAG}_staged.json 2> gpurun_out/bench_${TAG}_staged.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
timeout 600 python bench.py --workload c4 --gpus 1 > gpurun_out/bench_${TAG}_c4_1gpu.json 2> gpurun_out/bench_${TAG}_c4_1gpu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline --core-arms-only > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"shard_|probe_gather_inbox|pull_misses|insert_merge_kernel" -s 489 -c 200 --csv \
  --log-file gpurun_out/launches_${TAG}_c4.csv python bench.py --workload c4 --gpus 1 --steps 2 --warmup 3 > gpurun_out/ncu_bench_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_gather_v8|pull_misses" -s 10 -c 4 \
  -o gpurun_out/hot_${TAG} -f python bench.py --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline --core-arms-only > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_gather_inbox|shard_dispatch" -s 6 -c 2 \
  -o gpurun_out/hot_${TAG}_c4 -f python bench.py --workload c4 --gpus 1 --steps 2 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv \
  --log-file gpurun_out/launches_${TAG}_mlp.csv python scripts/mlp_profile_driver.py > gpurun_out/ncu_mlp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_gemm_tcgen05" -s 3 -c 1 \
  -o gpurun_out/hot_${TAG}_mlp -f python scripts/mlp_profile_driver.py > gpurun_out/ncu_full_mlp.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
