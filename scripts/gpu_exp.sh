#!/bin/bash
# experiment pass: kernel variants + host-link probe
mkdir -p gpurun_out
{
python scripts/probe_variants.py ldg 1.0
python scripts/probe_variants.py ldg 0.9
for cfg in 4x4 4x3 8x2 4x2 8x1; do
  HPSX_PIPE_CFG=$cfg python scripts/probe_variants.py pipe 1.0
done
HPSX_PIPE_CFG=4x4 python scripts/probe_variants.py pipe 0.9
HPSX_PIPE_CFG=8x2 python scripts/probe_variants.py pipe 0.9
} > gpurun_out/variants.txt 2>&1
./tools/pcie_probe > gpurun_out/pcie_probe3.txt 2>&1
cat gpurun_out/variants.txt; grep "4 GiB" gpurun_out/pcie_probe3.txt
