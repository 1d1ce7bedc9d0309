"""PCIe experiments: can zero-copy SM reads and copy-engine DMA add up? cost of lig_upload_snapshot."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.engine import Engine
P, A, R = 4096, 1024, 1 << 20
snap = WL.make_snapshot(P, A)
eng = Engine(0, P, A, R)
eng.upload_snapshot(1, snap.packed)
hb = [WL.make_requests(R, A, seed=7 + b) for b in range(2)]
pin_in = [torch.from_numpy(hb[b].view(np.uint8).reshape(-1)).pin_memory() for b in range(2)]
pin_out = [torch.zeros(R * 8, dtype=torch.uint8).pin_memory() for b in range(2)]
d_in = torch.empty(R * 16, dtype=torch.uint8, device='cuda'); d_out = torch.empty(R * 8, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
zc = lambda: eng.schedule_batch_device(1, 5, pin_in[0].data_ptr(), R, pin_out[0].data_ptr(), s1.cuda_stream)
def dma():
    with torch.cuda.stream(s2): d_in.copy_(pin_in[1], non_blocking=True)
def both():
    zc(); dma()
a, b, c = t(zc), t(dma), t(both)
print(f"zero-copy batch alone {a*1e6:.0f} us ({16*R/a/1e9:.1f} GB/s in); DMA H2D alone {b*1e6:.0f} us ({16*R/b/1e9:.1f} GB/s); both concurrently {c*1e6:.0f} us ({32*R/c/1e9:.1f} GB/s aggregate in)")
# split one batch: fraction f by DMA+device kernel, rest zero-copy
for f in (0.0, 0.25, 0.4, 0.5):
    n_dma = int(R * f) // 1024 * 1024
    def split():
        if n_dma:
            with torch.cuda.stream(s2):
                d_in[: n_dma * 16].copy_(pin_in[0][: n_dma * 16], non_blocking=True)
            eng.schedule_batch_device(1, 5, d_in.data_ptr(), n_dma, d_out.data_ptr(), s2.cuda_stream)
            with torch.cuda.stream(s2):
                pin_out[0][: n_dma * 8].copy_(d_out[: n_dma * 8], non_blocking=True)
        eng.schedule_batch_device(1, 5, pin_in[0].data_ptr() + n_dma * 16, R - n_dma, pin_out[0].data_ptr() + n_dma * 8, s1.cuda_stream)
    dt = t(split)
    print(f"split f={f}: {dt*1e6:.0f} us/batch -> {R/dt:.3e} dec/s")
up = t(lambda: eng.upload_snapshot(2, snap.packed), 50)
print(f"lig_upload_snapshot (pack+H2D+build+sync): {up*1e6:.0f} us")
t0 = time.perf_counter()
for _ in range(50): snap.packed.blob()
print(f"host pack only (lig_pack_snapshot into numpy): {(time.perf_counter()-t0)/50*1e6:.0f} us")
hs = t(lambda: eng.schedule_batch_ptr(2, 5, pin_in[0].data_ptr(), R, pin_out[0].data_ptr()))
print(f"lig_schedule_batch (pinned host, blocking): {hs*1e6:.0f} us -> {R/hs:.3e} dec/s")
eng.close()
