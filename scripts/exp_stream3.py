"""Streaming latency: launch-per-flush vs persistent doorbell kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from llm_instance_gateway_b200 import host as H, workload as WL
from llm_instance_gateway_b200.engine import Engine
from test_host_runtime import snapshot_to_podmetrics
P, A = 256, 64
snap = WL.make_snapshot(P, A)
# raw round trip of one call
e = Engine(0, 256, 64, 8192)
e.upload_snapshot(1, snap.packed)
reqs = WL.make_requests(4096, A)
e.stream_open()
for n in (1, 8, 64, 1024):
    r = np.ascontiguousarray(reqs[:n]); ts = []
    for i in range(600):
        t0 = time.perf_counter(); e.stream_submit(1, i, r); ts.append(time.perf_counter() - t0)
    ts = np.array(ts[100:]) * 1e6
    print(f"doorbell n={n}: p50 {np.percentile(ts,50):.1f} us p99 {np.percentile(ts,99):.1f} us")
e.stream_close()
for n in (1, 8, 64, 1024):
    r = np.ascontiguousarray(reqs[:n]); ts = []
    for i in range(600):
        t0 = time.perf_counter(); e.schedule_batch(1, i, r); ts.append(time.perf_counter() - t0)
    ts = np.array(ts[100:]) * 1e6
    print(f"launch   n={n}: p50 {np.percentile(ts,50):.1f} us p99 {np.percentile(ts,99):.1f} us")
e.close()
prov = H.HostProvider(snapshot_to_podmetrics(snap))
models = [WL.adapter_name(a) for a in range(A)] + [WL.UNKNOWN_MODEL]
models = models + models
critical = [False] * (A + 1) + [True] * (A + 1)
for door, window, rate, threads in [(0, 2, 1e5, 32), (1, 2, 1e5, 32), (1, 0, 1e5, 32), (1, 2, 3e5, 64), (0, 2, 3e5, 64)]:
    s = H.HostScheduler(prov, max_pods=256, max_adapters=64, max_batch=1 << 14, flush_size=4096, batch_window_us=window,
                        refresh_interval_ms=50, busy_poll=True, caller_spin_us=100, use_doorbell=bool(door))
    lat, nerr = s.stream_bench(rate, 2.0, threads, models, critical, seed=5)
    st = s.stats()
    print(f"doorbell={door} window={window}us rate={rate:.0e} threads={threads}: n={len(lat)} p50={np.percentile(lat,50):.1f} p90={np.percentile(lat,90):.1f} "
          f"p99={np.percentile(lat,99):.1f} p99.9={np.percentile(lat,99.9):.1f} max={lat.max():.0f} us errors={nerr} avg_batch={st['scheduled']/max(st['batches'],1):.1f} refreshes={st['refreshes']}")
    s.close()
prov.close()
