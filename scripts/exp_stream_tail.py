"""Streaming tail diagnosis: 3 x 10 s of C5; max latency vs the slowest device call / Flush."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
for i in range(3):
    r = bench.streaming_leg(0, seconds=10.0)
    print(json.dumps({k: r[k] for k in ("latency_us", "service_latency_us", "over_200us", "slowest_device_call_us",
                                        "slowest_flush_us", "slowest_call_batch", "slowest_call_cpu_us", "batches", "avg_batch")}), flush=True)
