"""Streaming config (BASELINE.json configs[4]): 100K req/s Poisson into a 256-pod pool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from llm_instance_gateway_b200 import host as H, workload as WL
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_runtime import snapshot_to_podmetrics
P, A = 256, 64
snap = WL.make_snapshot(P, A)
prov = H.HostProvider(snapshot_to_podmetrics(snap))
models = [WL.adapter_name(a) for a in range(A)] + [WL.UNKNOWN_MODEL]
models = models + models
critical = [False] * (A + 1) + [True] * (A + 1)
for window, flush, threads, rate in [(50, 4096, 32, 1e5), (20, 4096, 32, 1e5), (5, 4096, 32, 1e5), (0, 4096, 32, 1e5),
                                     (20, 4096, 64, 1e5), (20, 64, 64, 1e5), (20, 4096, 64, 3e5), (20, 4096, 96, 1e6)]:
    s = H.HostScheduler(prov, max_pods=256, max_adapters=64, max_batch=1 << 14, flush_size=flush,
                        batch_window_us=window, refresh_interval_ms=50)
    lat, nerr = s.stream_bench(rate, 2.0, threads, models, critical, seed=5)
    st = s.stats()
    print(f"window={window}us flush={flush} threads={threads} rate={rate:.0e}: n={len(lat)} achieved={len(lat)/2.0:.0f}/s "
          f"p50={np.percentile(lat,50):.1f} p90={np.percentile(lat,90):.1f} p99={np.percentile(lat,99):.1f} "
          f"p99.9={np.percentile(lat,99.9):.1f} max={lat.max():.0f} us errors={nerr} batches={st['batches']} "
          f"avg_batch={st['scheduled']/max(st['batches'],1):.1f} max_batch={st['max_batch']} refreshes={st['refreshes']}")
    s.close()
prov.close()
