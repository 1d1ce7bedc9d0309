#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_gpu_tests.log
tail -8 gpurun_out/r02j_gpu_tests.log
timeout 300 python scripts/build_split.py > gpurun_out/r02j_split.txt 2>&1; cat gpurun_out/r02j_split.txt
timeout 300 python scripts/tick_cost.py > gpurun_out/r02j_tick.txt 2>&1; cat gpurun_out/r02j_tick.txt
LIG_BUILD_DEBUG=1 timeout 300 python scripts/tick_cost.py 2>&1 | grep stamps | sed -n "3,5p" > gpurun_out/r02j_stamps.txt; cat gpurun_out/r02j_stamps.txt
