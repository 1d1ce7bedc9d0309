"""C oracle == Python oracle on random and adversarial pools; known answers for the pick RNG."""
import math
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from helpers import np_int31n
from oracle import lig_oracle_py as PY

NAMES = ["a0", "a1", "a2", "a3", "a4", "a5"]
SPECIAL_KV = [0.0, -0.0, 0.8, 0.8000000000000002, 0.7999999999999999, 1.0, float("inf"),
              float("-inf"), float("nan"), 1e-300, 5e-324, -0.25, 0.3, 1.0 / 3.0]

pod_st = st.fixed_dictionaries({
    "waiting_queue_size": st.one_of(st.integers(0, 8), st.integers(40, 60), st.integers(-5, 300),
                                    st.sampled_from([5, 6, 49, 50, 2**31 - 1, -(2**31)])),
    "kv_cache_usage_percent": st.one_of(st.floats(0, 1), st.sampled_from(SPECIAL_KV),
                                        st.floats(allow_nan=True, allow_infinity=True)),
    "max_active_models": st.one_of(st.integers(0, 4), st.sampled_from([-1, 70000, 2**40])),
    "active_models": st.lists(st.sampled_from(NAMES), max_size=4, unique=True),
})


def to_py(pods):
    return [PY.PodMetrics(PY.Pod(f"pod-{i}", f"address-{i}"),
                          PY.Metrics(active_models={m: 1 for m in p["active_models"]},
                                     max_active_models=p["max_active_models"],
                                     waiting_queue_size=p["waiting_queue_size"],
                                     kv_cache_usage_percent=p["kv_cache_usage_percent"]))
            for i, p in enumerate(pods)]


@settings(max_examples=400, deadline=None)
@given(pods=st.lists(pod_st, min_size=0, max_size=40),
       model=st.sampled_from(NAMES + ["unknown"]), critical=st.booleans(),
       seed=st.integers(0, 2**64 - 1), key=st.integers(0, 2**64 - 1))
def test_c_equals_python(oracle, pods, model, critical, seed, key):
    named = [dict(p, name=f"pod-{i}", address=f"address-{i}") for i, p in enumerate(pods)]
    pool = oracle.Pool(named)
    rc, idx = pool.filter(model, critical)
    pypods = to_py(pods)
    sched = PY.Scheduler(PY.StaticProvider(pypods))
    req = PY.LLMRequest(model=model, resolved_target_model=model, critical=critical)
    st_py, surv = sched.filter_only(req)
    assert rc == st_py
    assert idx == [int(p.pod.name.split("-")[1]) for p in surv]
    rc2, pod, n = pool.schedule(model, critical, seed, key)
    s_py, pod_py, n_py, _ = sched.Schedule(req, seed, key)
    assert (rc2, pod, n) == (s_py, pod_py, n_py)


def test_splitmix64_known_answers(oracle):
    # SplitMix64 (Steele, Lea, Flood 2014; Vigna's splitmix64.c), seed 1234567: widely published
    # first outputs.
    want = [6457827717110365317, 3203168211198807973, 9817491932198370423,
            4593380528125082431, 16408922859458223821]
    assert oracle.splitmix64_stream(1234567, 5) == want
    src = PY.SplitMix64Source(1234567)
    assert [src.next() for _ in range(5)] == want


def test_int31n_matches_go_algorithm(oracle):
    rng = np.random.default_rng(1)
    for n in [1, 2, 3, 4, 5, 7, 8, 63, 64, 65, 1000, 4096, 32768, 2**31 - 1, 2**30 + 1]:
        for state in rng.integers(0, 1 << 63, 50).tolist():
            k = oracle.int31n(state, n)
            assert 0 <= k < n
            assert k == PY.SplitMix64Source(state).int31n(n)
    # vectorised numpy restatement used by the full-size GPU tests
    keys = rng.integers(0, 1 << 64, 5000, dtype=np.uint64)
    ns = rng.integers(1, 5000, 5000)
    got = np_int31n(99, keys, ns)
    for i in range(0, 5000, 37):
        assert got[i] == oracle.int31n(99 ^ int(keys[i]), int(ns[i]))


def test_int31n_rejection_path(oracle):
    # n = 2^30 + 1 rejects ~50 % of draws: exercises the resample loop in all three restatements
    n = 2**30 + 1
    mx = (1 << 31) - 1 - (1 << 31) % n
    hit = 0
    for state in range(200):
        src = PY.SplitMix64Source(state)
        first = src.int31()
        if first > mx:
            hit += 1
        assert oracle.int31n(state, n) == PY.SplitMix64Source(state).int31n(n)
    assert hit > 20


def test_range_filter_corner_cases(oracle):
    # leastQueuing: max starts at 0 (filter.go:104) => with all-negative queues max stays 0
    pods = [dict(name=f"p{i}", address="", waiting_queue_size=q, kv_cache_usage_percent=0.0,
                 max_active_models=0, active_models=[]) for i, q in enumerate([-10, -7, -1])]
    rc, idx = oracle.Pool(pods).filter_func("leastQueuingFilterFunc")
    assert rc == 0 and idx == [0, 1]          # thr = -10 + (0 - -10)/3 = -7
    # Go truncated division: (max-min)/len with 7/3 = 2
    pods = [dict(pods[0], waiting_queue_size=q) for q in [1, 3, 8]]
    assert oracle.Pool(pods).filter_func("leastQueuingFilterFunc")[1] == [0, 1]
    # leastKV: all NaN => nobody passes, nil error (filter.go:148-152)
    pods = [dict(pods[0], kv_cache_usage_percent=float("nan")) for _ in range(3)]
    assert oracle.Pool(pods).filter_func("leastKVCacheFilterFunc") == (0, [])
    # and through the whole tree a critical request then yields ([], nil) => EMPTY
    assert oracle.Pool(pods).filter("x", True) == (oracle.LIGO_EMPTY, [])
    # empty pool: critical falls to the sheddable branch and is dropped
    assert oracle.Pool([]).filter("x", True) == (oracle.LIGO_DROP, [])
    assert oracle.Pool([]).filter("x", False) == (oracle.LIGO_DROP, [])


def test_no_fma_in_kv_threshold(oracle):
    # a case where min + (max-min)/n differs between fused and separately rounded evaluation
    found = 0
    rng = np.random.default_rng(5)
    for _ in range(2000):
        kv = rng.random(3).tolist()
        mn, mx = min(kv), max(kv)
        thr = mn + (mx - mn) / 3.0
        pods = [dict(name=f"p{i}", address="", waiting_queue_size=0, kv_cache_usage_percent=v,
                     max_active_models=0, active_models=[]) for i, v in enumerate(kv + [thr])]
        # add a pod exactly at the 4-pod threshold to make the boundary matter
        mn4, mx4 = min(kv + [thr]), max(kv + [thr])
        thr4 = mn4 + (mx4 - mn4) / 4.0
        rc, idx = oracle.Pool(pods).filter_func("leastKVCacheFilterFunc")
        want = [i for i, v in enumerate(kv + [thr]) if v >= mn4 and v <= thr4]
        assert idx == want
        found += 1
    assert found == 2000


def test_soa_variant_equals_structure_preserving_port(oracle):
    """The optimised-CPU datapoint (columns + bitmaps + mask words) against the reference-shaped
    port, on the synthetic configs and on adversarial pools (status, n, pick and survivor mask)."""
    import math
    from llm_instance_gateway_b200 import workload as WL
    from llm_instance_gateway_b200.packer import REQ_DTYPE, pack_pod_metrics
    from llm_instance_gateway_b200.backend import Metrics, Pod, PodMetrics
    for P, A, R in [(8, 4, 2000), (64, 32, 1024), (300, 24, 3000), (1, 0, 64), (65, 3, 500)]:
        snap = WL.make_snapshot(P, A, seed=P + A)
        reqs = WL.make_requests(R, A, seed=R)
        reqs["adapter_id"][::17] = -3
        p = snap.packed
        want, wm = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 11, True)
        got, gm = oracle.soa_schedule_batch(P, A, p.kv, p.q, p.n_active, p.max_active, p.bitmap, reqs, 11, True, 3)
        assert np.array_equal(got, want) and np.array_equal(gm, wm), (P, A)
    rng = np.random.default_rng(8)
    special = [0.0, -0.0, 0.8, 0.8000000000000002, 1.0, math.inf, -math.inf, math.nan, 5e-324, -0.25]
    for it in range(60):
        P = int(rng.integers(0, 80))
        pods = [PodMetrics(Pod(f"pod-{i}", f"address-{i}"),
                           Metrics(WaitingQueueSize=int(rng.choice([0, 5, 6, 49, 50, -1, 2**31 - 1, -(2**31), int(rng.integers(0, 70))])),
                                   KVCacheUsagePercent=float(rng.choice(special)) if rng.random() < 0.3 else float(np.round(rng.random(), 2)),
                                   MaxActiveModels=int(rng.integers(0, 4)),
                                   ActiveModels={a: 1 for a in rng.choice(["a0", "a1", "a2"], size=int(rng.integers(0, 3)), replace=False)}))
                for i in range(P)]
        p = pack_pod_metrics(pods)
        names = [None] * p.A
        for k, v in p.adapter_ids.items():
            names[v] = k
        reqs = np.zeros(2 * (p.A + 2), dtype=REQ_DTYPE)
        reqs["adapter_id"] = list(range(-1, p.A + 1)) * 2
        reqs["flags"] = [0] * (p.A + 2) + [1] * (p.A + 2)
        reqs["rand_key"] = rng.integers(0, 1 << 64, len(reqs), dtype=np.uint64)
        records = [dict(name=x.Pod.Name, address=x.Pod.Address, waiting_queue_size=x.Metrics.WaitingQueueSize,
                        kv_cache_usage_percent=x.Metrics.KVCacheUsagePercent, max_active_models=x.Metrics.MaxActiveModels,
                        active_models=list(x.Metrics.ActiveModels)) for x in pods]
        want, wm = oracle.Pool(records).schedule_batch(names, "zz-unknown", reqs, it, True)
        got, gm = oracle.soa_schedule_batch(p.P, p.A, p.kv, p.q, p.n_active, p.max_active, p.bitmap, reqs, it, True, 1)
        assert np.array_equal(got, want) and np.array_equal(gm, wm), it
