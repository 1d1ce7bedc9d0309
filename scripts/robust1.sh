#!/bin/bash
fmt='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "OK value=%.3e us/step=%.3f frac=%.3f launches=%s reps=%s" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d.get("gpu_launches"), d["config"]["timed_region_repeats"]))'
run1() { timeout 120 python bench.py "$@" --no-cpu-baseline --no-streaming 2>>gpurun_out/robust.err | python -c "$fmt" "N=1 $*" || echo "N=1 $* FAILED"; }
run1 --steps 1 --warmup 0 --min-seconds 0.1
run1 --steps 3 --warmup 3 --min-seconds 0.1
run1 --steps 10 --warmup 3 --min-seconds 0.1
run1 --steps 20 --warmup 5 --min-seconds 0.1
run1 --steps 50 --warmup 10 --min-seconds 0.1
run1 --steps 200 --warmup 20
run1 --steps 1000 --warmup 10 --min-seconds 0.1
tail -3 gpurun_out/robust.err | cut -c1-200
