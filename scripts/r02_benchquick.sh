#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -3 gpurun_out/r02i_bench.err
timeout 600 python -m pytest tests/test_bench_contract.py tests/test_hermetic_replay.py -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1; tail -5 gpurun_out/r02i_tests.log
