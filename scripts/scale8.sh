#!/bin/bash
out=gpurun_out/scale8.txt
: > $out
for g in 0 1; do for n in 8; do
  LIG_GRAPH=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n+g)) \
    bench.py --gpus $n --steps 200 --warmup 20 --timed-only --min-seconds 0.5 2>> gpurun_out/scale8.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('graph=$g n=$n', 'value=%.3e us/step=%.2f' % (d['value'], d['ms_per_step']*1e3))" >> $out
done; done
# is it the host? same 8 ranks, but each rank's launch thread pinned away from the others
for g in 0 1; do
  LIG_GRAPH=$g OMP_NUM_THREADS=4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29820+g)) \
    bench.py --gpus 8 --steps 200 --warmup 20 --timed-only --min-seconds 0.5 --requests-per-gpu 2097152 2>> gpurun_out/scale8.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('graph=$g n=8 R=2M', 'value=%.3e us/step=%.2f' % (d['value'], d['ms_per_step']*1e3))" >> $out
done
cat $out; nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader | head -8
