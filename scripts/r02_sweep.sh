#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "queue_modes or device_pointer or config_parity" > gpurun_out/r02b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_gpu_tests.log
tail -5 gpurun_out/r02b_gpu_tests.log
timeout 1500 python scripts/sweep_persist.py > gpurun_out/r02b_sweep.txt 2>&1; cat gpurun_out/r02b_sweep.txt
