#!/bin/bash
# odd-argument robustness of bench.py (what a driver might pass)
fmt='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "OK value=%.3e us/step=%.3f launches=%s" % (d["value"], d["ms_per_step"]*1e3, d.get("gpu_launches")))'
run1() { timeout 120 python bench.py "$@" --no-cpu-baseline --no-streaming --min-seconds 0.1 2>>gpurun_out/robust.err | python -c "$fmt" "N=1 $*" || echo "N=1 $* FAILED"; }
run2() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29900 + RANDOM % 50)) bench.py --gpus 2 "$@" --min-seconds 0.1 2>>gpurun_out/robust.err | python -c "$fmt" "N=2 $*" || echo "N=2 $* FAILED"; }
run1 --steps 1 --warmup 0
run1 --steps 2 --warmup 1
run1 --steps 3 --warmup 3
run1 --steps 10 --warmup 3
run1 --steps 1000 --warmup 10
run1 --steps 7 --warmup 2 --workload C2
run1 --steps 7 --warmup 2 --workload C3
run1 --steps 5 --warmup 2 --workload C5
run2 --steps 10 --warmup 3
run2 --steps 10 --warmup 3 --workload C3
run2 --steps 10 --warmup 3 --scaling strong
tail -5 gpurun_out/robust.err | cut -c1-200
