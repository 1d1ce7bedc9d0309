#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_gpu_tests.log
tail -30 gpurun_out/r02f_gpu_tests.log
