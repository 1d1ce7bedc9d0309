#!/bin/bash
for R in 1048576 2097152 4194304 16777216; do
  timeout 100 python bench.py --steps 50 --warmup 5 --requests-per-gpu $R --no-cpu-baseline --no-streaming --min-seconds 0.3 --timed-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
R=$R
print('R=%d us/step=%.2f value=%.3e frac=%.3f' % (R, d['ms_per_step']*1e3, d['value'], (24*R+589824)/(d['ms_per_step']*1e-3)/6585.1e9))"
done | tee gpurun_out/exp10.txt
