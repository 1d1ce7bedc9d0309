"""Where does the end-to-end (host-buffer) step go?  One B200, C4.

For the model-id call (4 B in, 4 B out per decision, buffers page-locked + device-mapped) and the
descriptor call (16 B in, 8 B out) prints: blocking call time, async-pipelined time per batch
(several tickets in flight), and for scale the cudaMemcpyAsync H2D / D2H rate of the same buffers.
Run twice: LIG_NUMA=1 (default: buffers on the GPU's NUMA node) and LIG_NUMA=0."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.engine import Engine
from llm_instance_gateway_b200.packer import pack_models

R, P, A = 1 << 20, 4096, 1024
lib = N.load()


class Pinned:
    def __init__(self, nbytes):
        self.ptr = lib.lig_host_alloc(nbytes)
        assert self.ptr
        self.n = nbytes
        self.np = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))


def numa_of(ptr):
    """NUMA node holding the page at ptr (move_pages query), or None."""
    libc = ctypes.CDLL(None, use_errno=True)
    pages = (ctypes.c_void_p * 1)(ptr)
    status = (ctypes.c_int * 1)(-1)
    rc = libc.syscall(279, 0, 1, pages, None, status, 0)   # SYS_move_pages on x86-64
    return status[0] if rc == 0 else None


def bench(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    torch.cuda.init()
    snap = WL.make_snapshot(P, A, seed=WL.SNAPSHOT_SEED)
    eng = Engine(0, P, A, R)
    eng.upload_snapshot(1, snap.packed)
    eng.upload_models(1, pack_models(WL.make_models(A), snap.packed))
    ids = [Pinned(R * 4) for _ in range(4)]
    for b, buf in enumerate(ids):
        buf.np[:] = WL.make_model_requests(R, A, seed=5 + b).view(np.uint8)
    mout = [Pinned(R * 4) for _ in range(4)]
    reqs = [Pinned(R * 16) for _ in range(4)]
    for b, buf in enumerate(reqs):
        buf.np[:] = WL.make_requests(R, A, seed=9 + b).view(np.uint8).reshape(-1)
    out = [Pinned(R * 8) for _ in range(4)]
    print(f"LIG_NUMA={os.environ.get('LIG_NUMA', '(unset)')} LIG_HOST_TMA={os.environ.get('LIG_HOST_TMA', '(unset)')} "
          f"buffer page on NUMA node {numa_of(ids[0].ptr)}; cpus usable {len(os.sched_getaffinity(0))}")
    i = [0]

    def models_blocking():
        k = i[0] % 4
        i[0] += 1
        eng.schedule_models_batch_ptr(1, 7, 0, ids[k].ptr, R, mout[k].ptr)

    def desc_blocking():
        k = i[0] % 4
        i[0] += 1
        eng.schedule_batch_ptr(1, 7, reqs[k].ptr, R, out[k].ptr)

    def desc_async4():
        ts = [eng.schedule_batch_async(1, 7, reqs[k].ptr, R, out[k].ptr) for k in range(4)]
        for t in ts:
            eng.schedule_wait(t)

    t = bench(models_blocking)
    print(f"model-id call, blocking:      {t:7.1f} us/batch  -> {R / t * 1e6:.3g} decisions/s, {R * 4 / t / 1e3:.1f} GB/s each way")
    t = bench(desc_blocking)
    print(f"descriptor call, blocking:    {t:7.1f} us/batch  -> {R / t * 1e6:.3g} decisions/s, in {R * 16 / t / 1e3:.1f} GB/s out {R * 8 / t / 1e3:.1f} GB/s")
    t = bench(desc_async4, n=10) / 4
    print(f"descriptor call, 4 in flight: {t:7.1f} us/batch  -> {R / t * 1e6:.3g} decisions/s, in {R * 16 / t / 1e3:.1f} GB/s out {R * 8 / t / 1e3:.1f} GB/s")
    # the DMA engines over the same buffers, for scale
    dev = torch.empty(R * 16, dtype=torch.uint8, device="cuda")
    rt = None
    for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            pass
    if rt is not None:
        rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        for name, nbytes, src in (("4 MiB", R * 4, ids[0]), ("16 MiB", R * 16, reqs[0])):
            h2d = bench(lambda: rt.cudaMemcpyAsync(dev.data_ptr(), src.ptr, nbytes, 1, None))
            d2h = bench(lambda: rt.cudaMemcpyAsync(src.ptr, dev.data_ptr(), nbytes, 2, None))
            print(f"cudaMemcpyAsync {name}: H2D {nbytes / h2d / 1e3:.1f} GB/s  D2H {nbytes / d2h / 1e3:.1f} GB/s (back to back on the null stream)")
        # hybrids for the model-id call: which legs go through the copy engines?
        d_ids = torch.empty(R, dtype=torch.uint32, device="cuda")
        d_out = torch.empty(R, dtype=torch.uint32, device="cuda")
        streams = [torch.cuda.Stream() for _ in range(2)]

        def hybrid(n_chunks, dma_in, dma_out):
            k = i[0] % 4
            i[0] += 1
            per = R // n_chunks
            for c in range(n_chunks):
                st = streams[c % 2]
                lo = c * per
                src = ids[k].ptr + lo * 4
                dst = mout[k].ptr + lo * 4
                if dma_in:
                    rt.cudaMemcpyAsync(d_ids.data_ptr() + lo * 4, src, per * 4, 1, st.cuda_stream)
                    src = d_ids.data_ptr() + lo * 4
                kout = d_out.data_ptr() + lo * 4 if dma_out else dst
                eng.schedule_models_batches_device(1, 7, lo, [src], per, [kout], stream=st.cuda_stream)
                if dma_out:
                    rt.cudaMemcpyAsync(dst, kout, per * 4, 2, st.cuda_stream)
            for st in streams:
                st.synchronize()

        for dma_in, dma_out, label in ((True, False, "DMA in, kernel writes host"), (False, True, "kernel reads host, DMA out"),
                                       (True, True, "DMA both ways")):
            for n_chunks in (1, 2, 4, 8):
                t = bench(lambda: hybrid(n_chunks, dma_in, dma_out))
                print(f"model-id hybrid [{label}] {n_chunks} chunk(s): {t:7.1f} us/batch -> {R / t * 1e6:.3g} decisions/s")
    eng.close()


if __name__ == "__main__":
    main()
