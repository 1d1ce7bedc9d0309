#!/bin/bash
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pct in 0 25 33; do
  LIG_HOST_DMA_PERCENT=$pct timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-streaming --min-seconds 0.2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('dma_percent=$pct e2e=%.3e value=%.3e frac=%.3f' % (d['e2e']['value'], d['value'], d['roofline']['frac']))"
done
