#!/bin/bash
# N GPUs: the C-ABI-only multi-GPU tests, then the bench through torchrun (one rank per GPU)

N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_host_runtime.py -m gpu -x -q > gpurun_out/r02n_multi_tests_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_multi_tests_n$N.log
tail -15 gpurun_out/r02n_multi_tests_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err
tail -c 1500 gpurun_out/r02n_bench_n$N.json; tail -5 gpurun_out/r02n_bench_n$N.err
