#!/bin/bash
# async uploads: parity tests, then the bench's e2e leg
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_delta.py tests/test_host_runtime.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --stream-seconds 2 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value', d['value'], 'e2e', e['value'], 'blocking', e['value_blocking_uploads'], 'resident', e['value_snapshot_resident'], 'descr', e['descriptor_call']['value'], e['descriptor_call']['value_snapshot_resident'])
PY
