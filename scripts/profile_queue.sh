#!/bin/bash
tag=${1:-r01d}
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0 > gpurun_out/launches_${tag}.stdout 2>gpurun_out/launches_${tag}.err
ncu --set full --clock-control none --import-source on -k regex:lig_pick_queue -s 11 -c 1 -o gpurun_out/prof_queue_${tag} -f \
  python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0 > /dev/null 2>>gpurun_out/launches_${tag}.err
ls -la gpurun_out/*${tag}*
