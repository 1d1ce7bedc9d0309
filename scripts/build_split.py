#!/usr/bin/env python3
"""Where does the class build spend its time?  P = 4096 pods with A = 0 adapters (2 classes: staging +
the shared stages only) vs A = 1024 (2050 classes)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llm_instance_gateway_b200 import workload as WL  # noqa: E402
from llm_instance_gateway_b200.engine import Engine  # noqa: E402
from llm_instance_gateway_b200.packer import pack_columns  # noqa: E402

snap = WL.make_snapshot(4096, 1024)
p = snap.packed
for A in (0, 32, 256, 1024):
    pk = pack_columns(p.kv, snap.q64, np.minimum(p.n_active, A if A else 0) if A == 0 else p.n_active, snap.max_active64, p.bitmap[:A])
    e = Engine(0, max_pods=4096, max_adapters=max(A, 1), max_batch=1024)
    e.upload_snapshot(1, pk)
    blob = torch.from_numpy(pk.blob()).cuda()
    st = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e.upload_snapshot_device(2, 4096, A, blob.data_ptr(), st.cuda_stream)
    st.synchronize()
    with torch.cuda.stream(st):
        torch.cuda._sleep(200_000)
        ev0.record(st)
        for _ in range(20):
            e.upload_snapshot_device(2, 4096, A, blob.data_ptr(), st.cuda_stream)
        ev1.record(st)
    st.synchronize()
    print(f"P=4096 A={A}: D2D copy + counters memset + build + header D2H = {ev0.elapsed_time(ev1) / 20 * 1e3:.1f} us per tick", flush=True)
    e.close()
