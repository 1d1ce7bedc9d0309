#!/bin/bash
# BASELINE.json configs C2, C3, C4 on one GPU (resident throughput, e2e), one summary line each.
fmt='import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], "value=%.3e us/step=%.3f frac=%.3f e2e=%.3e launches/region=%d build_us=%.1f rebuild=%.3e scan=%.3e" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["snapshot_build_us"], d["with_snapshot_rebuild"]["value"], d["direct_scan"]["value"]))'
for w in C2 C3 C4; do
  timeout 120 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --no-streaming --min-seconds 0.3 2>>gpurun_out/configs.err | tee gpurun_out/bench_$w.json | python -c "$fmt" $w
done | tee gpurun_out/configs.txt
