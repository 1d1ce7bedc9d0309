#!/bin/bash
# compute-sanitizer over the GPU test suite itself (every kernel family, every queue mode)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
out=gpurun_out/r02_compute_sanitizer_tests.txt
: > $out
echo "===== compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q --deselect tests/test_bench_contract.py =====" >> $out
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q --deselect tests/test_bench_contract.py::test_gpu_arm_contract 2>&1 | grep -v "^=========$" | tail -12 | cut -c1-300 >> $out
echo "===== compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_delta.py tests/test_feedback.py tests/test_gpu_models.py -m gpu -x -q =====" >> $out
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_delta.py tests/test_feedback.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | grep -v "^=========$" | tail -8 | cut -c1-300 >> $out
cat $out
