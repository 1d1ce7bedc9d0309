#!/bin/bash
# model-id host path with 16-byte-per-lane transactions: parity tests, PCIe anatomy, bench e2e
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_hermetic_replay.py tests/test_host_runtime.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
{ timeout 300 python scripts/exp_pcie.py | head -4; echo "LIG_MODELS_VEC4=0:"; LIG_MODELS_VEC4=0 timeout 300 python scripts/exp_pcie.py | head -2; } > gpurun_out/r02t_pcie.txt 2>&1; cat gpurun_out/r02t_pcie.txt
timeout 600 python bench.py --steps 20 --warmup 5 --stream-seconds 2 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value', d['value'], 'e2e', e['value'], 'blocking', e['value_blocking_uploads'], 'resident', e['value_snapshot_resident'], 'descr', e['descriptor_call']['value'], e['descriptor_call']['value_snapshot_resident'])
PY
