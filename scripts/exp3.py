"""GPU experiment: host enqueue cost per launch, zero-copy vs DMA host path, small-batch latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.engine import Engine
from llm_instance_gateway_b200.packer import PICK_DTYPE
P, A, R = 4096, 1024, 1 << 20
snap = WL.make_snapshot(P, A)
eng = Engine(0, P, A, R)
eng.upload_snapshot(1, snap.packed)
hb = [WL.make_requests(R, A, seed=7 + b) for b in range(4)]
nb = 12
d_reqs = [torch.from_numpy(hb[b % 4].view(np.uint8).reshape(-1)).cuda() for b in range(nb)]
d_out = [torch.zeros(R * 8, dtype=torch.uint8, device='cuda') for b in range(nb)]
st = torch.cuda.Stream()
K = 100
rp = [d_reqs[i % nb].data_ptr() for i in range(K)]; op = [d_out[i % nb].data_ptr() for i in range(K)]
for _ in range(3):
    eng.schedule_batches_device(1, 1, rp, R, op, st.cuda_stream)
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.schedule_batches_device(1, 1, rp, R, op, st.cuda_stream)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"K={K}: host enqueue {1e6*(t1-t0)/K:.2f} us/launch, total {1e6*(t2-t0)/K:.2f} us/step (streams={os.environ.get('LIG_QUEUE_STREAMS')})")

# zero-copy: kernel reads pinned host memory directly and writes picks to pinned host memory
pin_in = [torch.from_numpy(hb[b].view(np.uint8).reshape(-1)).pin_memory() for b in range(4)]
pin_out = [torch.zeros(R * 8, dtype=torch.uint8).pin_memory() for b in range(4)]
want = eng.schedule_batch(1, 5, hb[0])
for mode in ("zero_copy_both", "zero_copy_in", "zero_copy_out", "dma_pipeline"):
    def step(i):
        b = i % 4
        if mode == "zero_copy_both":
            eng.schedule_batch_device(1, 5, pin_in[b].data_ptr(), R, pin_out[b].data_ptr(), st.cuda_stream)
        elif mode == "zero_copy_in":
            eng.schedule_batch_device(1, 5, pin_in[b].data_ptr(), R, d_out[b].data_ptr(), st.cuda_stream)
            with torch.cuda.stream(st): pin_out[b].copy_(d_out[b], non_blocking=True)
        elif mode == "zero_copy_out":
            with torch.cuda.stream(st): d_reqs[b].copy_(pin_in[b], non_blocking=True)
            eng.schedule_batch_device(1, 5, d_reqs[b].data_ptr(), R, pin_out[b].data_ptr(), st.cuda_stream)
        else:
            eng.schedule_batch_ptr(1, 5, pin_in[b].data_ptr(), R, pin_out[b].data_ptr())
    for i in range(3): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for i in range(n): step(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    ok = np.array_equal(pin_out[0].numpy().view(PICK_DTYPE), want)
    print(f"{mode}: {dt*1e6:.0f} us/step -> {R/dt:.3e} dec/s, in {16*R/dt/1e9:.1f} GB/s out {8*R/dt/1e9:.1f} GB/s parity={ok}")

# small batches: latency of one host-buffer call (C5-like sizes)
for r in (64, 1024, 16384):
    for mode in ("zero_copy", "dma"):
        ts = []
        for i in range(220):
            t0 = time.perf_counter()
            if mode == "zero_copy":
                eng.schedule_batch_device(1, i, pin_in[0].data_ptr(), r, pin_out[0].data_ptr(), st.cuda_stream)
                st.synchronize()
            else:
                eng.schedule_batch_ptr(1, i, pin_in[0].data_ptr(), r, pin_out[0].data_ptr())
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts[20:]) * 1e6
        print(f"R={r} {mode}: p50 {np.percentile(ts,50):.1f} us p99 {np.percentile(ts,99):.1f} us")

# concurrent H2D on two streams
n = 16 << 20
h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = [torch.empty(n, dtype=torch.uint8, device='cuda') for _ in range(2)]
ss = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(10):
    for k in range(2):
        with torch.cuda.stream(ss[k]): d[k].copy_(h[k], non_blocking=True)
torch.cuda.synchronize()
print(f"2-stream concurrent h2d: {20*n/(time.perf_counter()-t0)/1e9:.1f} GB/s")
eng.close()
