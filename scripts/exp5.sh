#!/bin/bash
fmt='import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], "us/step=%.2f value=%.3e frac=%.3f parity=%d" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["frac"], d["parity_checked"]))'
for pf in 0 1; do for ns in 2 4 6; do for ppt in 2 4; do
  LIG_PREFETCH=$pf LIG_QUEUE_STREAMS=$ns LIG_PICK_PER_THREAD=$ppt timeout 60 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --min-seconds 0.3 2>>gpurun_out/exp5.err | python -c "$fmt" "prefetch=$pf ns=$ns ppt=$ppt"
done; done; done | tee gpurun_out/exp5.txt
