#!/bin/bash
# the whole -m gpu suite + smoke on one GPU (no bench)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_bench_contract.py::test_gpu_arm_contract > gpurun_out/r02x_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_gpu_tests.log
tail -4 gpurun_out/r02x_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-160
