#!/bin/bash
# ncu evidence for profiles/: launch list of the timed region + full captures of the kernels.
tag=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0 > gpurun_out/launches_${tag}.stdout 2>gpurun_out/launches_${tag}.err
ncu --set full --clock-control none --import-source on -k regex:lig_pick -s 12 -c 3 -o gpurun_out/prof_pick_${tag} -f \
  python bench.py --steps 20 --warmup 3 --timed-only --min-seconds 0 > /dev/null 2>>gpurun_out/launches_${tag}.err
ncu --set full --clock-control none --import-source on -k regex:lig_class_build -c 1 -o gpurun_out/prof_build_${tag} -f \
  python bench.py --steps 5 --warmup 3 --timed-only --min-seconds 0 > /dev/null 2>>gpurun_out/launches_${tag}.err
ncu --set full --clock-control none --import-source on -k regex:lig_scan -c 1 -o gpurun_out/prof_scan_${tag} -f \
  python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>>gpurun_out/launches_${tag}.err
ls -la gpurun_out/*${tag}*
