#!/bin/bash
for ring in 6 8 11 16 24; do
  timeout 100 python bench.py --steps 200 --warmup 20 --ring $ring --no-cpu-baseline --no-streaming --min-seconds 0.3 --timed-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ring=$ring (%d MiB in+out) us/step=%.3f frac=%.3f' % ($ring*24, d['ms_per_step']*1e3, 25755648/(d['ms_per_step']*1e-3)/6585.1e9))"
done
