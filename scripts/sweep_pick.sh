#!/bin/bash
# Sweep the pick-kernel tuning knobs on one GPU; one JSON summary line per variant.
out=gpurun_out/sweep_pick.txt
: > $out
for ppt in 1 2 4; do
  for ns in 1 2 3; do
    LIG_PICK_PER_THREAD=$ppt LIG_QUEUE_STREAMS=$ns timeout 200 python bench.py --steps 100 --warmup 10 \
      --no-cpu-baseline --min-seconds 0.3 "$@" 2>>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ppt=$ppt ns=$ns us/step=%.2f value=%.3e frac=%.3f e2e=%.3e rebuild=%.3e build_us=%.1f scan=%.3e' % (d['ms_per_step']*1e3, d['value'], d['roofline']['frac'], d['e2e']['value'], d['with_snapshot_rebuild']['value'], d['snapshot_build_us'], d['direct_scan']['value']))" >> $out
  done
done
cat $out
