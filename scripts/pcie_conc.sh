#!/bin/bash
# per-GPU host-link rates with N GPUs busy at once (N = 1, 2, 4, 8): profiles/pcie_conc_r02.txt
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/host.txt; free -g | head -n 2 >> gpurun_out/host.txt
: > gpurun_out/pcie_conc.txt
for N in 1 2 4 8; do
  echo "== $N GPU(s) at once" >> gpurun_out/pcie_conc.txt
  START=$(python3 -c "import time; print(time.time() + 12)")
  for ((i = 0; i < N; i++)); do ./tools/pcie_probe_conc $i $START >> gpurun_out/pcie_conc.txt 2>&1 & done
  wait
done
cat gpurun_out/host.txt gpurun_out/pcie_conc.txt
