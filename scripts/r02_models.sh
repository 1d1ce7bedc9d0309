#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_gpu_tests.log
tail -30 gpurun_out/r02c_gpu_tests.log
SWEEP_ONLY="tab=1 g=1 st=3 bulk=0,tab=1 g=2 st=2 bulk=0,tab=1 g=2 st=4 bulk=0,tab=1 g=3 st=3 bulk=0,tab=1 g=1 st=2 bulk=0,tab=1 g=1 st=4 bulk=0,tab=1 g=1 st=6 bulk=0,loop tab=1 ctas=3,tab=0 g=2 st=2 bulk=0" timeout 300 python scripts/sweep_persist.py > gpurun_out/r02c_sweep.txt 2>&1; cat gpurun_out/r02c_sweep.txt
