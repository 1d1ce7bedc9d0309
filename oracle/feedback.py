"""Sequential definition of the opt-in in-batch load feedback.  TEST INFRASTRUCTURE ONLY.

The reference never writes pod metrics while scheduling (pkg/ext-proc/scheduling/scheduler.go:
113-122): inside one scrape window every request of a class sees the same survivors.  The
simulator's load balancer does account for every pick at once — the chosen pod's prefill queue grows
(simulations/llm_ig_simulation/src/loadbalancer.py:608-625).  include/lig.h's
lig_schedule_batch_feedback_device does so at window granularity; this is its definition:

    for each window w (sub_batch requests of every shard):
        schedule the window's requests with the reference tree against the CURRENT queue sizes
        WaitingQueueSize[p] += number of requests of this window (all shards) that picked p

Two interchangeable schedulers are provided for the inner step: the class-table oracle (fast, used
for large cases) and the structure-preserving port (tests check they agree).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import binding as B

INT32_MAX = (1 << 31) - 1


def _windows(R: int, sub_batch: int, shards: Sequence[Tuple[int, int]]) -> int:
    return max((hi - lo + sub_batch - 1) // sub_batch for lo, hi in shards) if shards else 0


def schedule_batch_feedback(P: int, A: int, kv, q, n_active, max_active, bitmap, reqs: np.ndarray, seed: int,
                            sub_batch: int, shards: Optional[List[Tuple[int, int]]] = None,
                            thresholds=(0.8, 5, 50)):
    """Returns (picks[R], total picks per pod[P], list of per-window histograms)."""
    R = len(reqs)
    shards = shards or [(0, R)]
    q = np.ascontiguousarray(q, dtype=np.int64).copy()
    out = np.zeros(R, dtype=B.PICK_DTYPE)
    total = np.zeros(P, dtype=np.int64)
    per_window = []
    for w in range(_windows(R, sub_batch, shards)):
        tab = B.ClassTable(P, A, kv, q.astype(np.int32), n_active, max_active, bitmap, thresholds)
        hist = np.zeros(P, dtype=np.int64)
        for lo, hi in shards:
            a, b = lo + w * sub_batch, min(lo + (w + 1) * sub_batch, hi)
            if a >= b:
                continue
            picks = tab.schedule_batch(np.ascontiguousarray(reqs[a:b]), seed)
            out[a:b] = picks
            pods = picks["pod_idx"][picks["pod_idx"] >= 0]
            if P:
                hist += np.bincount(pods, minlength=P)
        q = np.minimum(q + hist, INT32_MAX)
        total += hist
        per_window.append(hist)
    return out, total, per_window


def schedule_batch_feedback_port(pods: List[dict], adapter_names: List[str], unknown: str, reqs: np.ndarray, seed: int,
                                 sub_batch: int, shards: Optional[List[Tuple[int, int]]] = None):
    """The same with the structure-preserving port of Scheduler.Schedule as the inner scheduler."""
    R = len(reqs)
    shards = shards or [(0, R)]
    pods = [dict(p) for p in pods]
    out = np.zeros(R, dtype=B.PICK_DTYPE)
    P = len(pods)
    total = np.zeros(P, dtype=np.int64)
    for w in range(_windows(R, sub_batch, shards)):
        pool = B.Pool(pods)
        hist = np.zeros(P, dtype=np.int64)
        for lo, hi in shards:
            a, b = lo + w * sub_batch, min(lo + (w + 1) * sub_batch, hi)
            if a >= b:
                continue
            picks, _ = pool.schedule_batch(adapter_names, unknown, np.ascontiguousarray(reqs[a:b]), seed)
            out[a:b] = picks
            sel = picks["pod_idx"][picks["pod_idx"] >= 0]
            if P:
                hist += np.bincount(sel, minlength=P)
        for p in range(P):
            pods[p]["waiting_queue_size"] = min(pods[p]["waiting_queue_size"] + int(hist[p]), INT32_MAX)
        total += hist
    return out, total
