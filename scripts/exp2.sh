#!/bin/bash
out=gpurun_out/exp2.txt
: > $out
fmt='import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], "us/step=%.2f value=%.3e frac=%.3f e2e=%.3e" % (d["ms_per_step"]*1e3, d["value"], d["roofline"]["frac"], d["e2e"]["value"]))'
for ppt in 2 4 8; do
  for ns in 3 4 6 8; do
    LIG_PICK_PER_THREAD=$ppt LIG_QUEUE_STREAMS=$ns timeout 200 python bench.py --steps 100 --warmup 10 \
      --no-cpu-baseline --min-seconds 0.3 2>>gpurun_out/exp2.err | python -c "$fmt" "R=1M ppt=$ppt ns=$ns" >> $out
  done
done
for R in 4194304 16777216; do
  for ns in 1 3; do
    LIG_PICK_PER_THREAD=4 LIG_QUEUE_STREAMS=$ns timeout 300 python bench.py --steps 40 --warmup 5 --requests-per-gpu $R \
      --no-cpu-baseline --min-seconds 0.3 2>>gpurun_out/exp2.err | python -c "$fmt" "R=$R ppt=4 ns=$ns" >> $out
  done
done
python - >> $out <<'PY'
import torch, time
for mb in (1, 16, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(20): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); h2d = 20*n/(time.perf_counter()-t)/1e9
    t=time.perf_counter()
    for _ in range(20): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); d2h = 20*n/(time.perf_counter()-t)/1e9
    print(f"pcie {mb} MiB: h2d {h2d:.1f} GB/s d2h {d2h:.1f} GB/s")
PY
cat $out
# launch list + full capture of the pick kernel
LIG_PICK_PER_THREAD=4 LIG_QUEUE_STREAMS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 0 > /dev/null 2>>gpurun_out/exp2.err
LIG_PICK_PER_THREAD=4 LIG_QUEUE_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:lig_pick -s 10 -c 3 \
  -o gpurun_out/prof_pick_r01 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 0 > /dev/null 2>>gpurun_out/exp2.err
LIG_PICK_PER_THREAD=4 LIG_QUEUE_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:lig_class_build -s 2 -c 1 \
  -o gpurun_out/prof_build_r01 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --min-seconds 0 > /dev/null 2>>gpurun_out/exp2.err
ls -la gpurun_out/
