#!/bin/bash
# closing check on ONE GPU: the whole -m gpu suite, smoke(), both bench arms (no profiler)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
T=r02v
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gpu_tests.log
tail -4 gpurun_out/${T}_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['value_snapshot_resident'], 'tick', d['snapshot_tick']['us'], 'models', d['model_requests']['value'])
s=d['streaming']; print('stream', s['latency_us'], s['service_latency_us']['p99.9'], s['service_latency_us']['max'], s['slowest_device_call_us'], s['slowest_call_cpu_us'], s.get('attempts'))
PY
