#!/bin/bash
fmt='import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], "us/step=%.2f value=%.3e frac=%.3f parity=%d reps=%d minmedmax=%s" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["frac"], d["parity_checked"], d["config"]["timed_region_repeats"], d["config"]["region_ms_min_med_max"]))'
for g in 1; do for ns in 4; do
  LIG_GRAPH=$g LIG_QUEUE_STREAMS=$ns timeout 60 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-streaming --min-seconds 0.3 2>>gpurun_out/exp6.err | python -c "$fmt" "graph=$g ns=$ns R=1M"
done; done | tee gpurun_out/exp6.txt
for g in 0 1; do for R in 131072 65536 1024; do
  LIG_GRAPH=$g timeout 60 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-streaming --min-seconds 0.3 --requests-per-gpu $R 2>>gpurun_out/exp6.err | python -c "$fmt" "graph=$g R=$R"
done; done | tee -a gpurun_out/exp6.txt
