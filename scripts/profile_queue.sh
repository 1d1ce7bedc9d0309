#!/bin/bash
# Re-capture the evidence of the default queue kernel (run under gpurun, ONE GPU):
#   launch list of a --timed-only bench run  -> gpurun_out/launches_<tag>.csv
#   ncu --set full of one K=20 queue launch  -> gpurun_out/prof_queue_<tag>.ncu-rep
# then, back in the build container:
#   python scripts/make_traffic_json.py gpurun_out/prof_queue_<tag>.ncu-rep C4 20 > profiles/traffic.json
#   python scripts/summarize_ncu.py gpurun_out/prof_queue_<tag>.ncu-rep > profiles/<tag>_queue_kernel_ncu.txt
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > gpurun_out/launches_${tag}.stdout 2>gpurun_out/launches_${tag}.err
ncu --set full --clock-control none --import-source on -k regex:lig_pick_persistent -s 3 -c 1 -o gpurun_out/prof_queue_${tag} -f \
  python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0.001 > /dev/null 2>>gpurun_out/launches_${tag}.err
ls -la gpurun_out/*${tag}*
