"""Cost of one snapshot refresh through the C++ host runtime at C4 scale (f1: the packer)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from llm_instance_gateway_b200 import host as H, workload as WL
from test_host_runtime import snapshot_to_podmetrics
for P, A in [(256, 64), (4096, 1024)]:
    snap = WL.make_snapshot(P, A)
    prov = H.HostProvider(snapshot_to_podmetrics(snap))
    s = H.HostScheduler(prov, max_pods=P, max_adapters=A, max_batch=4096)
    ts = []
    for _ in range(20):
        s.Refresh(); ts.append(s.refresh_timing())
    print(f"P={P} A={A}: pack {np.median([t['pack_us'] for t in ts]):.0f} us, upload {np.median([t['upload_us'] for t in ts]):.0f} us per refresh")
    s.close(); prov.close()
