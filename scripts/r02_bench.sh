#!/bin/bash
# tests + smoke + both bench arms + ncu captures of the default queue kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_gpu_tests.log
tail -25 gpurun_out/r02d_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02d_smoke.log 2>&1; tail -3 gpurun_out/r02d_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; tail -c 600 gpurun_out/r02d_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 3000 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lig_pick_persistent -s 3 -c 1 -o gpurun_out/r02d_persist python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/r02d_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/r02d_ncu_launches.log 2>&1
ls -la gpurun_out | tail -12
