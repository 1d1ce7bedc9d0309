#!/usr/bin/env python3
"""Sweep the knobs of the persistent TMA-pipelined pick kernel on the bench workload (C4, one GPU).
Every configuration is a fresh lig_ctx (the knobs are read at lig_create); timing as in bench.py
(spin-kernel gate, CUDA events on the launching stream, >L2 ring of distinct batches)."""
import itertools
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llm_instance_gateway_b200 import workload as WL  # noqa: E402
from llm_instance_gateway_b200.engine import Engine  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
cfg = WL.CONFIGS["C4"]
R, P, A = cfg["R"], cfg["P"], cfg["A"]
if len(sys.argv) > 1:
    R = int(sys.argv[1])
snap = WL.make_snapshot(P, A)
dev = torch.device("cuda", 0)
nb = 11
host = [WL.make_requests(R, A, seed=WL.REQUEST_SEED + b) for b in range(4)]
d_reqs, d_out = [], []
for b in range(nb):
    base = torch.from_numpy(host[b % 4].view(np.uint8).reshape(-1)).to(dev)
    if b >= 4:
        v = base.view(torch.int64).clone()
        v[1::2] ^= (0x9E3779B97F4A7C15 * (b + 1)) & 0x7FFFFFFFFFFFFFFF
        base = v.view(torch.uint8).view(-1, 16).roll(shifts=b * 7919, dims=0).contiguous().view(-1)
    d_reqs.append(base)
    d_out.append(torch.zeros(R * 8, dtype=torch.uint8, device=dev))
rp = [t.data_ptr() for t in d_reqs]
op = [t.data_ptr() for t in d_out]
stream = torch.cuda.Stream()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def measure(env, Ks=(20, 200), reps=15):
    for k in list(os.environ):
        if k.startswith("LIG_"):
            del os.environ[k]
    os.environ.update(env)
    eng = Engine(0, max_pods=P, max_adapters=A, max_batch=R)
    eng.upload_snapshot(1, snap.packed)
    out = {}
    for K in Ks:
        def run(first, seed):
            idx = [(first + i) % nb for i in range(K)]
            eng.schedule_batches_device(1, seed, [rp[i] for i in idx], R, [op[i] for i in idx], stream.cuda_stream)
        with torch.cuda.stream(stream):
            for w in range(3):
                run(w, w)
        torch.cuda.synchronize()
        ts = []
        for r in range(reps):
            with torch.cuda.stream(stream):
                torch.cuda._sleep(100_000)
                ev0.record(stream)
                run(r * K, 100 + r)
                ev1.record(stream)
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        ms = float(np.median(ts))
        us_step = ms * 1e3 / K
        out[K] = (us_step, 24 * R / (us_step * 1e-6) / 1e9 / PEAK, min(ts) * 1e3 / K)
    eng.close()
    return out


configs = [("merged(r1)", {"LIG_PICK_KERNEL": "merged"})]
for tab in (1, 0):
    for ct in (2, 3):
        configs.append((f"loop tab={tab} ctas={ct}", {"LIG_PICK_KERNEL": "loop", "LIG_PERSIST_CTAS": str(ct), "LIG_TAB_SMEM": str(tab)}))
for tab, g, st, bulk in itertools.product((1, 0), (1, 2, 3), (2, 3, 4, 6), (0, 1)):
    if st % g:
        continue
    configs.append((f"tma tab={tab} g={g} st={st} bulk={bulk}", {"LIG_TMA_GROUPS": str(g), "LIG_TMA_STAGES": str(st),
                                                                   "LIG_TMA_BULK_STORE": str(bulk), "LIG_TAB_SMEM": str(tab)}))
if os.environ.get("SWEEP_ONLY"):
    keep = os.environ["SWEEP_ONLY"].split(",")
    configs = [c for c in configs if any(k in c[0] for k in keep)]
_last = [time.time()]


def _watchdog():          # a deadlocked kernel must not burn the GPU lease
    while True:
        time.sleep(2)
        if time.time() - _last[0] > 45:
            print("WATCHDOG: configuration stalled, aborting", flush=True)
            os._exit(3)


threading.Thread(target=_watchdog, daemon=True).start()
print(f"# R={R} P={P} A={A}; peak {PEAK} GB/s; us/step (frac of peak, 24 B/decision) median of 15, [min]")
for name, env in configs:
    _last[0] = time.time()
    try:
        m = measure(env)
        print(f"{name:28s} " + "  ".join(f"K={K}: {v[0]:7.3f} us ({v[1]:.3f}) [min {v[2]:.3f}]" for K, v in m.items()), flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"{name:28s} FAILED {ex!r}", flush=True)
