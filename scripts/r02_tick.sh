#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_delta.py tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_tests.log
tail -20 gpurun_out/r02g_tests.log
timeout 300 python scripts/tick_cost.py > gpurun_out/r02g_tick.txt 2>&1; cat gpurun_out/r02g_tick.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lig_class_build -s 4 -c 1 -o gpurun_out/r02g_build python scripts/tick_cost.py > gpurun_out/r02g_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['snapshot_build_us'], json.dumps(d['streaming'])[:1200])"
