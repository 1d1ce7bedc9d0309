#!/usr/bin/env python3
"""Cost of one metrics-refresh tick on the device (C4: P=4096, A=1024): full upload vs 1 % delta."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from llm_instance_gateway_b200 import workload as WL  # noqa: E402
from llm_instance_gateway_b200.engine import Engine  # noqa: E402
from test_gpu_delta import mutate  # noqa: E402

for cfg in ("C4", "C5"):
    c = WL.CONFIGS[cfg]
    snap = WL.make_snapshot(c["P"], c["A"])
    packed2, delta, _ = mutate(snap, 0.01, seed=1)
    e = Engine(0, max_pods=c["P"], max_adapters=c["A"], max_batch=1024)
    e.upload_snapshot(1, snap.packed)
    blob = torch.from_numpy(snap.packed.blob()).cuda()
    st = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e.upload_snapshot_device(2, c["P"], c["A"], blob.data_ptr(), st.cuda_stream)
    st.synchronize()
    with torch.cuda.stream(st):
        torch.cuda._sleep(4_000_000)     # ~2 ms: all 20 ticks are enqueued before the device starts
        ev0.record(st)
        for _ in range(20):
            e.upload_snapshot_device(2, c["P"], c["A"], blob.data_ptr(), st.cuda_stream)
        ev1.record(st)
    st.synchronize()
    dev_us = ev0.elapsed_time(ev1) / 20 * 1e3

    def wall(fn, n=50):
        for _ in range(5):
            fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e6
    k = [2]

    def full():
        e.upload_snapshot(2, packed2)

    def dlt():
        e.update_snapshot(2, 1, **delta)
    print(f"{cfg}: device D2D+build {dev_us:.1f} us | blocking full upload {wall(full):.1f} us | blocking 1% delta "
          f"({len(delta['pod_idx'])} pods) {wall(dlt):.1f} us", flush=True)
    e.close()
