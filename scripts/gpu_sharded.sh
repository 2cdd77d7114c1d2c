#!/bin/bash
# Model-parallel (configs[3]-shaped) pass on N GPUs of one box: parity tests of both exchanges, then the c4 bench
# with the fused peer-memory exchange and with NCCL all-to-all-v.  usage: bash scripts/gpu_sharded.sh <N> [rows_per_gpu]
NG=${1:-2}
RPG=${2:-16000000}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${NG}.txt 2>&1
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/pytest_sharded_${NG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sharded_${NG}.log
for EX in p2p nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --workload c4 --gpus $NG --exchange $EX --rows-per-gpu $RPG --steps 20 --warmup 3 \
    > gpurun_out/bench_c4_${EX}_${NG}.json 2> gpurun_out/bench_c4_${EX}_${NG}.err
  echo "bench $EX exit $?"
done
tail -5 gpurun_out/pytest_sharded_${NG}.log
for EX in p2p nccl; do cat gpurun_out/bench_c4_${EX}_${NG}.json; tail -5 gpurun_out/bench_c4_${EX}_${NG}.err; done
