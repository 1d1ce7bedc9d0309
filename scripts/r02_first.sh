#!/bin/bash
# First GPU trip of round 2: parity of the new default path, bench, sweep, ncu.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.log
tail -5 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-streaming --no-cpu-baseline > gpurun_out/r02_bench_k20.json 2> gpurun_out/r02_bench_k20.err; tail -c 1500 gpurun_out/r02_bench_k20.json
timeout 900 python scripts/sweep_persist.py > gpurun_out/r02_sweep_persist.txt 2>&1; cat gpurun_out/r02_sweep_persist.txt
LIG_HOST_TMA=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-streaming --no-cpu-baseline > gpurun_out/r02_bench_k20_hosttma.json 2> gpurun_out/r02_bench_k20_hosttma.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lig_pick_persistent -s 2 -c 2 -o gpurun_out/r02_persist python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/r02_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out | tail -20
