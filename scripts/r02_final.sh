#!/bin/bash
# round-2 closing evidence on ONE GPU: tests + smoke + both bench arms + launch list + ncu of the
# queue kernel, the fast class build and the model-request kernel
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
T=r02m
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -4 gpurun_out/${T}_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lig_pick_persistent -s 3 -c 1 -o gpurun_out/${T}_persist -f python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lig_class_build_fast -s 2 -c 1 -o gpurun_out/${T}_build -f python scripts/tick_cost.py > gpurun_out/${T}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lig_pick_models_kernel -s 2 -c 1 -o gpurun_out/${T}_models -f python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_ncu3.log 2>&1
ls -la gpurun_out | grep ${T}
