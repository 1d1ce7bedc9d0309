#!/bin/bash
fmt='import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], "us/step=%.2f value=%.3e frac=%.3f parity=%d" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["frac"], d["parity_checked"]))'
for ppt in 4 8 16; do for ns in 2 4; do for pf in 0 1; do for g in 0 1; do
  LIG_GRAPH=$g LIG_PREFETCH=$pf LIG_QUEUE_STREAMS=$ns LIG_PICK_PER_THREAD=$ppt timeout 60 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-streaming --min-seconds 0.3 --timed-only 2>>gpurun_out/exp7.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], 'us/step=%.2f value=%.3e frac=%.3f' % (d['ms_per_step']*1e3, d['value'], 25755648/(d['ms_per_step']*1e-3)/6585.1e9))" "ppt=$ppt ns=$ns prefetch=$pf graph=$g"
done; done; done; done | tee gpurun_out/exp7.txt
