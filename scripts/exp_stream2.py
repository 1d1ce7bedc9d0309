"""Streaming latency with the host-runtime latency knobs (busy_poll, caller spin)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from llm_instance_gateway_b200 import host as H, workload as WL
from test_host_runtime import snapshot_to_podmetrics
P, A = 256, 64
snap = WL.make_snapshot(P, A)
prov = H.HostProvider(snapshot_to_podmetrics(snap))
models = [WL.adapter_name(a) for a in range(A)] + [WL.UNKNOWN_MODEL]
models = models + models
critical = [False] * (A + 1) + [True] * (A + 1)
for window, busy, spin, threads in [(5, 0, 0, 32), (5, 1, 0, 32), (5, 0, 100, 32), (5, 1, 100, 32), (0, 1, 100, 32), (2, 1, 100, 16), (5, 1, 100, 64)]:
    s = H.HostScheduler(prov, max_pods=256, max_adapters=64, max_batch=1 << 14, flush_size=4096,
                        batch_window_us=window, refresh_interval_ms=50, busy_poll=bool(busy), caller_spin_us=spin)
    lat, nerr = s.stream_bench(1e5, 2.0, threads, models, critical, seed=5)
    st = s.stats()
    print(f"window={window}us busy_poll={busy} spin={spin}us threads={threads}: n={len(lat)} p50={np.percentile(lat,50):.1f} "
          f"p90={np.percentile(lat,90):.1f} p99={np.percentile(lat,99):.1f} p99.9={np.percentile(lat,99.9):.1f} max={lat.max():.0f} us "
          f"errors={nerr} avg_batch={st['scheduled']/max(st['batches'],1):.1f}")
    s.close()
prov.close()
