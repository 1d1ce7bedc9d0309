#!/bin/bash
# N = 1, 2, 4, 8 back to back on one box (what the driver does at round end).
out=gpurun_out/scale_r01.jsonl
: > $out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline --no-streaming >> $out 2>> gpurun_out/scale_r01.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 200 --warmup 20 2>> gpurun_out/scale_r01.err | grep '^{' >> $out
  fi
done
# the strong-scaling reading of configs[3]: ONE 2^20-request batch sharded over 8 GPUs
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
  bench.py --gpus 8 --steps 200 --warmup 20 --scaling strong 2>> gpurun_out/scale_r01.err | grep '^{' >> $out
python - <<'PY'
import json
for line in open('gpurun_out/scale_r01.jsonl'):
    d=json.loads(line)
    print(d['n_gpus'], d['scaling'], 'value=%.3e us/step=%.2f frac=%.3f e2e=%.3e' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))
PY
