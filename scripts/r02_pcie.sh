#!/bin/bash
# e2e anatomy: which legs of the host-buffer call go through the copy engines?
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 300 python scripts/exp_pcie.py > gpurun_out/r02k_pcie.txt 2>&1
cat gpurun_out/r02k_pcie.txt
