#!/bin/bash
# model-request kernel rework: parity tests, then the bench's model_requests leg
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_hermetic_replay.py tests/test_host_runtime.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --stream-seconds 2 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'models', d['model_requests']['value'], d['model_requests']['us_per_step'], 'e2e', d['e2e']['value'], d['e2e']['value_snapshot_resident'])
print('strong', d['strong']['us_per_step'], 'stream', d['streaming']['latency_us'], d['streaming'].get('attempts'))
PY
