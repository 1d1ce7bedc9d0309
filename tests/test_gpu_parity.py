"""GPU parity tests proper: every result of the CUDA path (through the C ABI) is compared
bit-for-bit with the CPU oracle — status, survivor mask, n_survivors and picked pod index.

Bar: integer / index work, so the tolerance is zero everywhere (the one floating-point step,
min + (max-min)/n in leastKVCacheFilterFunc, only feeds comparisons; it must round exactly like
Go's three separate binary64 operations, and the survivor masks prove it).
"""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import golden_to_podmetrics, np_int31n
from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.backend import Metrics, Pod, PodMetrics
from llm_instance_gateway_b200.engine import Engine
from llm_instance_gateway_b200.packer import PICK_DTYPE, REQ_DTYPE, pack_pod_metrics
from llm_instance_gateway_b200.scheduling import LLMRequest, NewScheduler, StatusError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    e = Engine(0, max_pods=4096, max_adapters=1024, max_batch=1 << 20)
    yield e
    e.close()


class StaticProvider:
    def __init__(self, pods):
        self.pods = list(pods)

    def AllPodMetrics(self):
        return list(self.pods)


_epoch = [1000]


def next_epoch():
    _epoch[0] += 1
    return _epoch[0]


def all_class_requests(A, extra_ids=(), seed=3):
    """One request per (critical, adapter) class plus out-of-range adapter ids."""
    ids = list(range(A + 1)) + list(extra_ids)
    rng = np.random.default_rng(seed)
    reqs = np.zeros(2 * len(ids), dtype=REQ_DTYPE)
    reqs["adapter_id"] = ids + ids
    reqs["flags"] = [0] * len(ids) + [1] * len(ids)
    reqs["rand_key"] = rng.integers(0, 1 << 64, size=len(reqs), dtype=np.uint64)
    return reqs


def oracle_batch(oracle, records, adapter_names, reqs, seed, nthreads=1):
    pool = oracle.Pool(records)
    return pool.schedule_batch(adapter_names, WL.UNKNOWN_MODEL, reqs, seed, True, nthreads)


def compare(engine, oracle, packed, records, adapter_names, reqs, seed=0x5EED, scan=True, label=""):
    ep = next_epoch()
    engine.upload_snapshot(ep, packed)
    want, want_masks = oracle_batch(oracle, records, adapter_names, reqs, seed,
                                    nthreads=oracle.hardware_threads())
    got = engine.schedule_batch(ep, seed, reqs)
    for f in ("status", "n_survivors", "pod_idx"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"{label} fast path {f} differs at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]} req {reqs[bad[:5]]}"
    if scan:
        got2, masks = engine.schedule_scan(ep, seed, reqs, True, packed.W)
        assert np.array_equal(masks, want_masks), f"{label} survivor masks differ"
        assert np.array_equal(got2, want), f"{label} scan picks differ"
    return got


def pm_records(pods):
    """PodMetrics -> oracle records (names as given)."""
    return [dict(name=p.Pod.Name, address=p.Pod.Address, waiting_queue_size=p.Metrics.WaitingQueueSize,
                 kv_cache_usage_percent=p.Metrics.KVCacheUsagePercent,
                 max_active_models=p.Metrics.MaxActiveModels,
                 active_models=list(p.Metrics.ActiveModels)) for p in pods]


# ---- the reference's golden vectors through the GPU --------------------------------------------
def test_golden_TestFilter_through_the_scheduler_mirror(golden):
    for case in golden["TestFilter"]:
        if case["filter"]["name"] != "defaultFilter":
            continue            # filter_test.go:21-27 is a property of the Go node engine, not of the path
        sched = NewScheduler(StaticProvider([golden_to_podmetrics(p) for p in case["input"]]),
                             max_pods=64, max_adapters=64, max_batch=64, seed=7)
        req = LLMRequest(Model=case["req"]["model"], ResolvedTargetModel=case["req"]["resolved_target_model"],
                         Critical=case["req"]["critical"])
        survivors, err = sched.Filter(req)
        assert (err is not None) == case["err"], case["name"]
        assert [p.Name for p in survivors] == [p["name"] for p in case["output"]], case["name"]
        if case["err"]:
            with pytest.raises(StatusError) as ei:
                sched.Schedule(req)
            assert ei.value.code == "ResourceExhausted"          # -> 429, handlers/server.go:97-109
            assert "dropping request due to limited backend resources" in str(ei.value)
            assert str(ei.value).startswith("failed to apply filter, resulted 0 pods")
        else:
            for _ in range(5):
                assert sched.Schedule(req).Name == case["output"][0]["name"]
        sched.close()


def test_golden_TestFilterFunc_stage_vectors_through_the_tree(golden, engine, oracle):
    """filter_test.go:217-409 — each stage function is isolated inside the full tree by neutral
    settings of the other fields, and compared with the reference's expected survivors."""
    by_name = {c["name"]: c for c in golden["TestFilterFunc"]}

    def run(pods, model, critical):
        packed = pack_pod_metrics(pods)
        ep = next_epoch()
        engine.upload_snapshot(ep, packed)
        reqs = np.zeros(1, dtype=REQ_DTYPE)
        reqs[0] = (packed.adapter_id(model), int(critical), 5)
        picks, masks = engine.schedule_scan(ep, 1, reqs, True, packed.W)
        return [p for p in range(packed.P) if (int(masks[0, p >> 5]) >> (p & 31)) & 1], int(picks[0]["status"])

    # least queuing 0,3,10 -> [0,3]: critical, nobody has the adapter or room, equal KV
    c = by_name["least queuing"]
    pods = [PodMetrics(Pod(f"p{i}"), Metrics(WaitingQueueSize=p["waiting_queue_size"])) for i, p in enumerate(c["input"])]
    got, st = run(pods, "m", True)
    assert [c["input"][i]["waiting_queue_size"] for i in got] == [p["waiting_queue_size"] for p in c["output"]]
    # least kv 0,0.3,1.0 -> [0,0.3]: equal queues
    c = by_name["least kv cache"]
    pods = [PodMetrics(Pod(f"p{i}"), Metrics(KVCacheUsagePercent=p["kv_cache_usage_percent"])) for i, p in enumerate(c["input"])]
    got, st = run(pods, "m", True)
    assert [c["input"][i]["kv_cache_usage_percent"] for i in got] == [p["kv_cache_usage_percent"] for p in c["output"]]
    # noQueueAndLessThanKVCacheThresholdPredicate(0, 0.8): thresholds are parameters of the ABI
    c = by_name["noQueueAndLessThanKVCacheThresholdPredicate"]
    engine.set_thresholds(0.8, 0, 50)
    try:
        pods = [golden_to_podmetrics(p) for p in c["input"]]
        got, st = run(pods, "m", False)
        assert len(got) == 1 and st == N.LIG_OK
        assert (c["input"][got[0]]["waiting_queue_size"], c["input"][got[0]]["kv_cache_usage_percent"]) == (0, 0.0)
    finally:
        engine.set_thresholds()
    # low LoRA cost: first two pods survive
    c = by_name["low LoRA cost"]
    pods = [golden_to_podmetrics(p) for p in c["input"]]
    got, st = run(pods, c["req"]["resolved_target_model"], False)
    assert got == [0, 1]
    # empty inputs: no division by zero anywhere, request is shed
    ep = next_epoch()
    engine.upload_snapshot(ep, pack_pod_metrics([]))
    reqs = all_class_requests(0)
    for path in (engine.schedule_batch(ep, 1, reqs), engine.schedule_scan(ep, 1, reqs, False)[0]):
        assert (path["status"] == N.LIG_DROP).all() and (path["pod_idx"] == -1).all() and (path["n_survivors"] == 0).all()


def test_golden_hermetic_target_pod(golden):
    case = golden["TestHandleRequestBody"][0]
    resolved = case["models"][case["request_model"]]["target_models"][0]["name"]
    sched = NewScheduler(StaticProvider([golden_to_podmetrics(p) for p in case["pods"]]),
                         max_pods=64, max_adapters=64, max_batch=64)
    pod = sched.Schedule(LLMRequest(Model=case["request_model"], ResolvedTargetModel=resolved, Critical=False))
    assert pod.Address == next(h["raw_value"] for h in case["want_headers"] if h["key"] == "target-pod")
    sched.close()


# ---- seeded synthetic configs of BASELINE.json ----------------------------------------------------
@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C5"])
def test_config_parity_full(cfg, engine, oracle):
    c = WL.CONFIGS[cfg]
    R = min(c["R"], 65536)
    snap = WL.make_snapshot(c["P"], c["A"])
    reqs = WL.make_requests(R, c["A"])
    got = compare(engine, oracle, snap.packed, snap.pod_records(), snap.adapter_names(), reqs, label=cfg)
    assert (got["status"] == N.LIG_OK).any()
    # every class too (covers adapters the Zipf draw never hits)
    compare(engine, oracle, snap.packed, snap.pod_records(), snap.adapter_names(),
            all_class_requests(c["A"], extra_ids=(-1, c["A"] + 7, 2**31 - 1, -(2**31))), label=cfg + "/classes")


def test_config_C4_full_size(engine, oracle):
    """R = 2^20, P = 4096, A = 1024.  Every one of the 2050 class tables is compared with the
    oracle's survivor list; all 2^20 picks are then checked through a size-independent property
    (pick == class_list[Int31n(n)] recomputed in numpy) and a 32768-request sample directly."""
    c = WL.CONFIGS["C4"]
    snap = WL.make_snapshot(c["P"], c["A"])
    reqs = WL.make_requests(c["R"], c["A"])
    ep = next_epoch()
    engine.upload_snapshot(ep, snap.packed)
    seed = 0xC4
    got = engine.schedule_batch(ep, seed, reqs)
    pool = oracle.Pool(snap.pod_records())
    names = snap.adapter_names() + [WL.UNKNOWN_MODEL]
    A, P = c["A"], c["P"]
    lists, ns, sts = {}, np.zeros(2 * (A + 1), dtype=np.int64), np.zeros(2 * (A + 1), dtype=np.int64)
    table = np.full((2 * (A + 1), 64), -1, dtype=np.int64)
    for crit in (0, 1):
        for a in range(A + 1):
            rc, idx = pool.filter(names[a], bool(crit))
            st, n, lst = engine.read_class(ep, bool(crit), a, P)
            assert (st, n) == (rc, len(idx)) and lst.tolist() == idx, (crit, a)
            cls = crit * (A + 1) + a
            ns[cls], sts[cls] = n, rc
            if n > table.shape[1]:
                table = np.pad(table, ((0, 0), (0, n - table.shape[1])), constant_values=-1)
            table[cls, :n] = idx
    cls = (reqs["flags"] & 1).astype(np.int64) * (A + 1) + np.minimum(reqs["adapter_id"].astype(np.int64) % (1 << 32), A)
    n_req = ns[cls]
    assert np.array_equal(got["n_survivors"], n_req)
    assert np.array_equal(got["status"], sts[cls])
    ok = n_req > 0
    k = np_int31n(seed, reqs["rand_key"][ok], n_req[ok])
    assert np.array_equal(got["pod_idx"][ok], table[cls[ok], k])
    assert (got["pod_idx"][~ok] == -1).all()
    # direct oracle comparison on a sample, and the direct-scan kernel on a sub-sample
    sel = np.random.default_rng(4).choice(c["R"], 32768, replace=False)
    sub = np.ascontiguousarray(reqs[sel])
    want, want_masks = pool.schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, sub, seed, True,
                                           oracle.hardware_threads())
    assert np.array_equal(got[sel], want)
    got_scan, masks = engine.schedule_scan(ep, seed, np.ascontiguousarray(sub[:8192]), True, snap.packed.W)
    assert np.array_equal(got_scan, want[:8192]) and np.array_equal(masks, want_masks[:8192])
    # pick uniformity sanity: among requests of the most popular class, every survivor gets picked
    top = np.bincount(cls).argmax()
    if ns[top] > 1:
        assert set(got["pod_idx"][cls == top].tolist()) == set(table[top, :ns[top]].tolist())


# ---- edge cases -----------------------------------------------------------------------------------
def mkpods(qs, kvs, maxs=None, acts=None):
    n = len(qs)
    maxs = maxs or [0] * n
    acts = acts or [[] for _ in range(n)]
    return [PodMetrics(Pod(f"pod-{i}", f"address-{i}"),
                       Metrics(WaitingQueueSize=qs[i], KVCacheUsagePercent=kvs[i], MaxActiveModels=maxs[i],
                               ActiveModels={a: 1 for a in acts[i]})) for i in range(n)]


EDGE_POOLS = {
    "single_pod": mkpods([0], [0.1]),
    "single_pod_busy": mkpods([60], [0.95]),
    "all_nan_kv": mkpods([1, 2, 3], [math.nan] * 3),                       # -> EMPTY for critical
    "some_nan_kv": mkpods([0, 0, 0, 0], [math.nan, 0.2, math.nan, 0.1]),
    "inf_kv": mkpods([0, 0, 0], [math.inf, math.inf, math.inf]),
    "neg_inf_kv": mkpods([0, 0, 0], [-math.inf, 0.5, 0.2]),               # thr = NaN -> EMPTY
    "signed_zero_kv": mkpods([0, 0, 0, 0], [0.0, -0.0, -0.0, 0.0]),
    "negative_q": mkpods([-10, -7, -1, 4], [0.1, 0.1, 0.1, 0.1]),
    "int32_extremes": mkpods([2**31 - 1, -(2**31), 0, 7], [0.1, 0.2, 0.3, 0.4]),
    "q_all_high": mkpods([50, 51, 99, 50, 200], [0.5, 0.4, 0.3, 0.2, 0.1],
                         [2, 2, 2, 2, 2], [["x"], [], ["x", "y"], ["y"], []]),   # low-queue filter fails
    "q_boundaries": mkpods([49, 50, 5, 6, 0], [0.8, 0.8, 0.8000000000000002, 0.1, 0.7999999999999999]),
    "ties": mkpods([3] * 40, [0.25] * 40),
    "max_active_zero": mkpods([0, 0, 0], [0.1, 0.2, 0.3], [0, 0, 0], [["a"], ["b"], []]),
    "overfull": mkpods([1, 1, 1, 1], [0.3, 0.2, 0.1, 0.4], [1, 1, 2, 70000], [["a", "b"], ["a"], ["b"], []]),
    "affinity_vs_room": mkpods([0, 1, 2, 3, 4, 5], [0.5, 0.1, 0.9, 0.3, 0.2, 0.6], [2, 1, 2, 2, 0, 3],
                               [["a"], ["b"], ["a", "b"], [], ["a"], ["c"]]),
    "drop_all": mkpods([10, 3, 10], [0.9, 0.85, 0.85]),
}


@pytest.mark.parametrize("name", sorted(EDGE_POOLS))
def test_edge_pools(name, engine, oracle):
    pods = EDGE_POOLS[name]
    packed = pack_pod_metrics(pods)
    names = [None] * packed.A
    for k, v in packed.adapter_ids.items():
        names[v] = k
    reqs = all_class_requests(packed.A, extra_ids=(-1, packed.A + 1))
    reqs = np.concatenate([reqs] * 4)
    reqs["rand_key"] = np.random.default_rng(11).integers(0, 1 << 64, len(reqs), dtype=np.uint64)
    compare(engine, oracle, packed, pm_records(pods), names, np.ascontiguousarray(reqs), label=name)


def test_edge_statuses(engine):
    for name, crit, want in [("all_nan_kv", 1, N.LIG_EMPTY), ("neg_inf_kv", 1, N.LIG_EMPTY),
                             ("drop_all", 0, N.LIG_DROP), ("drop_all", 1, N.LIG_OK),
                             ("single_pod_busy", 0, N.LIG_DROP), ("single_pod_busy", 1, N.LIG_OK)]:
        packed = pack_pod_metrics(EDGE_POOLS[name])
        ep = next_epoch()
        engine.upload_snapshot(ep, packed)
        reqs = np.zeros(1, dtype=REQ_DTYPE)
        reqs[0] = (packed.A, crit, 1)
        assert int(engine.schedule_batch(ep, 0, reqs)[0]["status"]) == want, (name, crit)


@pytest.mark.parametrize("P", [1, 2, 31, 32, 33, 63, 64, 65, 100, 255, 257, 1000])
def test_ragged_pool_sizes(P, engine, oracle):
    A = 7
    snap = WL.make_snapshot(P, A, seed=P)
    reqs = WL.make_requests(512, A, seed=P + 1)
    compare(engine, oracle, snap.packed, snap.pod_records(), snap.adapter_names(), reqs, label=f"P={P}")
    compare(engine, oracle, snap.packed, snap.pod_records(), snap.adapter_names(),
            all_class_requests(A), label=f"P={P}/classes")


def test_random_small_pools_against_oracle(engine, oracle):
    """200 random pools with adversarial values (NaN, inf, ties, threshold boundaries)."""
    rng = np.random.default_rng(2024)
    kv_special = [0.0, -0.0, 0.8, 0.8000000000000002, 0.7999999999999999, 1.0, math.inf, -math.inf,
                  math.nan, 1e-300, 5e-324, -0.25, 1.0 / 3.0]
    q_special = [0, 5, 6, 49, 50, -1, 2**31 - 1, -(2**31)]
    adapters = ["a0", "a1", "a2", "a3"]
    for it in range(200):
        P = int(rng.integers(1, 70))
        qs = [int(rng.choice(q_special)) if rng.random() < 0.2 else int(rng.integers(0, 70)) for _ in range(P)]
        kvs = [float(rng.choice(kv_special)) if rng.random() < 0.25 else float(np.round(rng.random(), 2)) for _ in range(P)]
        maxs = [int(rng.integers(0, 4)) for _ in range(P)]
        acts = [list(rng.choice(adapters, size=int(rng.integers(0, 4)), replace=False)) for _ in range(P)]
        pods = mkpods(qs, kvs, maxs, acts)
        packed = pack_pod_metrics(pods)
        names = [None] * packed.A
        for k, v in packed.adapter_ids.items():
            names[v] = k
        compare(engine, oracle, packed, pm_records(pods), names, all_class_requests(packed.A, seed=it),
                seed=it, label=f"random#{it}")


def test_large_pool_unstaged_path(oracle):
    """P large enough that the pod columns do not fit in shared memory: the tree walk reads them
    from global memory instead (same results)."""
    P, A = 20000, 64
    e = Engine(0, max_pods=N.LIG_MAX_PODS, max_adapters=A, max_batch=4096)
    try:
        snap = WL.make_snapshot(P, A, seed=5)
        reqs = WL.make_requests(2048, A, seed=6)
        compare(e, oracle, snap.packed, snap.pod_records(), snap.adapter_names(), reqs, label="P=20000")
        # identical pods: every pod survives -> n_survivors = P needs all 15 bits
        pods = mkpods([1] * N.LIG_MAX_PODS, [0.5] * N.LIG_MAX_PODS)
        packed = pack_pod_metrics(pods)
        ep = next_epoch()
        e.upload_snapshot(ep, packed)
        r = all_class_requests(0)
        got = e.schedule_batch(ep, 9, r)
        assert (got["n_survivors"] == N.LIG_MAX_PODS).all() and (got["status"] == 0).all()
        want = np_int31n(9, r["rand_key"], np.full(len(r), N.LIG_MAX_PODS))
        assert np.array_equal(got["pod_idx"], want)
        got2, _ = e.schedule_scan(ep, 9, r, False)
        assert np.array_equal(got2, got)
    finally:
        e.close()


# ---- ABI behaviour --------------------------------------------------------------------------------
def test_epochs_and_errors(engine):
    lib = N.load()
    snap = WL.make_snapshot(64, 8, seed=1)
    snap2 = WL.make_snapshot(64, 8, seed=2)
    reqs = WL.make_requests(1000, 8)
    e = Engine(0, 64, 8, 1024)
    try:
        with pytest.raises(N.LigError) as ei:
            e.schedule_batch(1, 0, reqs)
        assert ei.value.code == N.LIG_ERR_NO_SNAPSHOT
        e.upload_snapshot(1, snap.packed)
        e.upload_snapshot(2, snap2.packed)
        a1 = e.schedule_batch(1, 0, reqs)           # both epochs resident
        a2 = e.schedule_batch(2, 0, reqs)
        assert not np.array_equal(a1, a2)
        e.upload_snapshot(3, snap.packed)           # evicts epoch 1
        with pytest.raises(N.LigError) as ei:
            e.schedule_batch(1, 0, reqs)
        assert ei.value.code == N.LIG_ERR_STALE_EPOCH and "epoch 1" in str(ei.value)
        assert np.array_equal(e.schedule_batch(3, 0, reqs), a1)
        assert np.array_equal(e.schedule_batch(2, 0, reqs), a2)
        with pytest.raises(N.LigError) as ei:       # over max_batch
            e.schedule_batch(3, 0, WL.make_requests(1025, 8))
        assert ei.value.code == N.LIG_ERR_INVALID
        with pytest.raises(N.LigError):             # over max_pods
            e.upload_snapshot(4, WL.make_snapshot(65, 8).packed)
        assert len(e.schedule_batch(3, 0, reqs[:0])) == 0
        # different seed or rand_key changes picks only where n_survivors > 1
        b = e.schedule_batch(3, 1, reqs)
        assert np.array_equal(b["status"], a1["status"]) and np.array_equal(b["n_survivors"], a1["n_survivors"])
        assert (b["pod_idx"][a1["n_survivors"] <= 1] == a1["pod_idx"][a1["n_survivors"] <= 1]).all()
        assert e.kernel_launches > 0 and e.sm_count >= 100
    finally:
        e.close()


def test_device_pointer_api_matches_host_api(engine):
    import torch
    c = WL.CONFIGS["C3"]
    snap = WL.make_snapshot(c["P"], c["A"])
    R = 100_003                                      # ragged tail for the 1024-request CTAs
    reqs = WL.make_requests(R, c["A"])
    ep = next_epoch()
    engine.upload_snapshot(ep, snap.packed)
    want = engine.schedule_batch(ep, 77, reqs)
    d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.zeros(R * 8, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()   # fills / copies above ran on torch's default stream
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        engine.schedule_batch_device(ep, 77, d_reqs.data_ptr(), R, d_out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(PICK_DTYPE), want)
    # pinned host buffers (the DMA-direct path) and pageable ones give the same bytes
    pin_in = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).pin_memory()
    pin_out = torch.zeros(R * 8, dtype=torch.uint8).pin_memory()
    engine.schedule_batch_ptr(ep, 77, pin_in.data_ptr(), R, pin_out.data_ptr())
    assert np.array_equal(pin_out.numpy().view(PICK_DTYPE), want)
    # a big pinned batch (zero-copy), the pageable bounce path and the device path agree
    R2 = 300_001
    reqs2 = WL.make_requests(R2, c["A"], seed=77)
    want2 = engine.schedule_batch(ep, 78, reqs2)                      # pageable: bounce path
    pin_in2 = torch.from_numpy(reqs2.view(np.uint8).reshape(-1)).pin_memory()
    pin_out2 = torch.zeros(R2 * 8, dtype=torch.uint8).pin_memory()
    engine.schedule_batch_ptr(ep, 78, pin_in2.data_ptr(), R2, pin_out2.data_ptr())
    assert np.array_equal(pin_out2.numpy().view(PICK_DTYPE), want2)
    d2 = torch.from_numpy(reqs2.view(np.uint8).reshape(-1)).cuda()
    o2 = torch.zeros(R2 * 8, dtype=torch.uint8, device="cuda")
    engine.schedule_batch_device(ep, 78, d2.data_ptr(), R2, o2.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(o2.cpu().numpy().view(PICK_DTYPE), want2)
    # snapshot upload from a blob already in HBM (the NCCL-broadcast path)
    blob = torch.from_numpy(snap.packed.blob()).cuda()
    ep2 = next_epoch()
    engine.upload_snapshot_device(ep2, snap.packed.P, snap.packed.A, blob.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
    with torch.cuda.stream(stream):
        engine.schedule_batch_device(ep2, 77, d_reqs.data_ptr(), R, d_out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(PICK_DTYPE), want)
    # direct scan through device pointers
    d_masks = torch.zeros(4096 * snap.packed.W, dtype=torch.int32, device="cuda")
    engine.schedule_scan_device(ep2, 77, d_reqs.data_ptr(), 4096, d_out.data_ptr(), d_masks.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(PICK_DTYPE)[:4096], want[:4096])


def test_thresholds_are_parameters(engine, oracle):
    snap = WL.make_snapshot(200, 16, seed=8)
    reqs = WL.make_requests(4096, 16, seed=9)
    try:
        for kv, qc, ql in [(0.5, 2, 10), (0.95, 0, 3), (0.0, 100, 1000)]:
            engine.set_thresholds(kv, qc, ql)
            oracle.set_thresholds(kv, qc, ql)
            compare(engine, oracle, snap.packed, snap.pod_records(), snap.adapter_names(), reqs,
                    label=f"thr={kv},{qc},{ql}")
    finally:
        engine.set_thresholds()
        oracle.set_thresholds()


QUEUE_MODES = {
    "tma": {},                                                             # default: resident CTAs fed by a TMA ring
    "tma_1group_2stage_bulk": {"LIG_TMA_GROUPS": "1", "LIG_TMA_STAGES": "2", "LIG_TMA_BULK_STORE": "1",
                               "LIG_PERSIST_CTAS": "1"},
    "tma_3group_3stage_bulk": {"LIG_TMA_GROUPS": "3", "LIG_TMA_STAGES": "3", "LIG_TMA_BULK_STORE": "1"},
    "tma_3group_2stage": {"LIG_TMA_GROUPS": "3", "LIG_TMA_STAGES": "2"},     # more groups than stages
    "tma_tables_in_global": {"LIG_TAB_SMEM": "0"},                         # strided tables through L1
    "loop": {"LIG_PICK_KERNEL": "loop"},                                   # resident CTAs, LDG + register prefetch
    "loop_tables_in_global": {"LIG_PICK_KERNEL": "loop", "LIG_TAB_SMEM": "0"},
    "merged": {"LIG_PICK_KERNEL": "merged"},                               # round-1 default: blockIdx.y = batch
    "streams": {"LIG_PICK_KERNEL": "merged", "LIG_MERGE_MAX": "1024"},     # one kernel per batch, forked streams
    "streams_pipelined": {"LIG_PICK_KERNEL": "merged", "LIG_MERGE_MAX": "1024", "LIG_PICK_PER_THREAD": "8",
                          "LIG_QUEUE_STREAMS": "2"},
}


@pytest.mark.parametrize("mode", sorted(QUEUE_MODES))
def test_batch_queue_modes_match_single_batches(mode, monkeypatch):
    """lig_schedule_batches_device: a queue of resident batches gives exactly the per-batch results
    in every execution mode (persistent TMA-ring kernel in several shapes, persistent LDG loop,
    merged launch, forked streams, software-pipelined kernel), for the first call, for a replay with another
    seed, and after the snapshot in the slot changed."""
    import torch
    for k, v in QUEUE_MODES[mode].items():
        monkeypatch.setenv(k, v)                    # the knobs are read at lig_create
    engine = Engine(0, max_pods=512, max_adapters=256, max_batch=1 << 18)
    try:
        c = WL.CONFIGS["C3"]
        snap = WL.make_snapshot(c["P"], c["A"], seed=41)
        R, nb = 150_000, 7
        host = [WL.make_requests(R, c["A"], seed=100 + b) for b in range(nb)]
        d_reqs = [torch.from_numpy(h.view(np.uint8).reshape(-1)).cuda() for h in host]
        d_out = [torch.zeros(R * 8, dtype=torch.uint8, device="cuda") for _ in range(nb)]
        torch.cuda.synchronize()   # the fills / copies ran on torch's default stream
        stream = torch.cuda.Stream()
        ep = next_epoch()
        engine.upload_snapshot(ep, snap.packed)

        def run_queue(epoch, seed, Rq=R):
            with torch.cuda.stream(stream):
                engine.schedule_batches_device(epoch, seed, [t.data_ptr() for t in d_reqs], Rq,
                                               [t.data_ptr() for t in d_out], stream.cuda_stream)
            stream.synchronize()
            return [t[: Rq * 8].cpu().numpy().view(PICK_DTYPE).copy() for t in d_out]

        for seed in (5, 5, 9000):                   # first call, replay, replay with a new seed
            got = run_queue(ep, seed)
            for b in range(nb):
                assert np.array_equal(got[b], engine.schedule_batch(ep, seed + b, host[b])), (mode, seed, b)
        # a different snapshot uploaded over the older slot: a cached graph must see the new tables
        snap2 = WL.make_snapshot(c["P"], c["A"], seed=42)
        ep2, ep3 = next_epoch(), next_epoch()
        engine.upload_snapshot(ep2, snap2.packed)
        engine.upload_snapshot(ep3, snap.packed)    # evicts `ep`
        for epoch in (ep2, ep3):
            got = run_queue(epoch, 77)
            for b in range(nb):
                assert np.array_equal(got[b], engine.schedule_batch(epoch, 77 + b, host[b])), (mode, epoch, b)
        # ragged and tiny batch sizes
        for Rs in (1, 1000, 1024, 4097):
            got = run_queue(ep3, 500, Rs)
            for b in range(nb):
                assert np.array_equal(got[b], engine.schedule_batch(ep3, 500 + b, host[b][:Rs])), (mode, Rs, b)
        # a queue of two, and of one
        with torch.cuda.stream(stream):
            engine.schedule_batches_device(ep3, 3, [d_reqs[0].data_ptr(), d_reqs[1].data_ptr()], R,
                                           [d_out[0].data_ptr(), d_out[1].data_ptr()], stream.cuda_stream)
            engine.schedule_batches_device(ep3, 9, [d_reqs[2].data_ptr()], R, [d_out[2].data_ptr()], stream.cuda_stream)
        stream.synchronize()
        assert np.array_equal(d_out[1].cpu().numpy().view(PICK_DTYPE), engine.schedule_batch(ep3, 4, host[1]))
        assert np.array_equal(d_out[2].cpu().numpy().view(PICK_DTYPE), engine.schedule_batch(ep3, 9, host[2]))
        # pick buffers that are only 8-byte aligned (the ABI's minimum): no TMA bulk store possible
        odd = [torch.zeros(R * 8 + 8, dtype=torch.uint8, device="cuda") for _ in range(3)]
        torch.cuda.synchronize()   # the fills ran on torch's default stream
        with torch.cuda.stream(stream):
            engine.schedule_batches_device(ep3, 21, [t.data_ptr() for t in d_reqs[:3]], R,
                                           [t.data_ptr() + 8 for t in odd], stream.cuda_stream)
        stream.synchronize()
        for b in range(3):
            assert np.array_equal(odd[b][8:].cpu().numpy().view(PICK_DTYPE), engine.schedule_batch(ep3, 21 + b, host[b])), (mode, b)
        # a queue longer than the kernel-parameter item table (97+ batches -> device item table)
        nq, Rq = 130, 3000
        big_in = torch.cat([d_reqs[b % nb][: Rq * 16] for b in range(nq)])
        big_out = torch.zeros(nq * Rq * 8, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            engine.schedule_batches_device(ep3, 1000, [big_in.data_ptr() + b * Rq * 16 for b in range(nq)], Rq,
                                           [big_out.data_ptr() + b * Rq * 8 for b in range(nq)], stream.cuda_stream)
        stream.synchronize()
        res = big_out.cpu().numpy().view(PICK_DTYPE).reshape(nq, Rq)
        for b in (0, 1, 95, 96, 97, 129):
            assert np.array_equal(res[b], engine.schedule_batch(ep3, 1000 + b, host[b % nb][:Rq])), (mode, b)
    finally:
        engine.close()


def test_doorbell_stream_matches_batch_path(oracle):
    """K3: micro-batches through the persistent doorbell kernel give exactly the results of
    lig_schedule_batch, across snapshot refreshes, for sizes 1..capacity; close/open cycles."""
    e = Engine(0, max_pods=512, max_adapters=64, max_batch=8192)
    try:
        snap = WL.make_snapshot(300, 24, seed=51)
        snap2 = WL.make_snapshot(300, 24, seed=52)
        reqs = WL.make_requests(4096, 24, seed=53)
        e.upload_snapshot(1, snap.packed)
        want_full = e.schedule_batch(1, 99, reqs)
        e.stream_open()
        e.stream_open()                                   # idempotent
        for n in (1, 2, 31, 256, 257, 1000, 4096):
            got = e.stream_submit(1, 99, np.ascontiguousarray(reqs[:n]))
            assert np.array_equal(got, want_full[:n]), n
        assert len(e.stream_submit(1, 99, reqs[:0])) == 0
        with pytest.raises(N.LigError):
            e.stream_submit(1, 99, WL.make_requests(4097, 24))
        # a snapshot refresh while the kernel is resident: the new tables must be seen (L2 reads)
        e.upload_snapshot(2, snap2.packed)
        got2 = e.stream_submit(2, 7, reqs)
        e.stream_close()
        assert np.array_equal(got2, e.schedule_batch(2, 7, reqs))
        want_o, _ = oracle.Pool(snap2.pod_records()).schedule_batch(snap2.adapter_names(), WL.UNKNOWN_MODEL, reqs, 7)
        assert np.array_equal(got2, want_o)
        with pytest.raises(N.LigError):                   # closed
            e.stream_submit(2, 7, reqs)
        e.stream_open()                                   # reopen: tickets restart
        assert np.array_equal(e.stream_submit(2, 7, reqs), got2)
        for i in range(300):                              # many tickets
            assert np.array_equal(e.stream_submit(2, i, np.ascontiguousarray(reqs[:8])), e.schedule_batch(2, i, np.ascontiguousarray(reqs[:8])))
        e.stream_close()
        e.stream_close()
    finally:
        e.close()


def test_async_submit_many_in_flight(oracle):
    """lig_schedule_batch_async/_wait: many batches in flight on one ctx from several threads while
    the snapshot is re-uploaded underneath them (the reader ring keeps a slot from being
    overwritten while a batch still reads it; an evicted epoch answers STALE_EPOCH, never garbage)."""
    import threading
    lib = N.load()
    P_, A_, R_ = 400, 32, 20_000
    snaps = [WL.make_snapshot(P_, A_, seed=70 + i) for i in range(4)]
    reqs = WL.make_requests(R_, A_, seed=80)
    wants = [oracle.Pool(s.pod_records()).schedule_batch(s.adapter_names(), WL.UNKNOWN_MODEL, reqs, 5, False,
                                                         oracle.hardware_threads())[0] for s in snaps]
    e = Engine(0, max_pods=512, max_adapters=32, max_batch=R_)
    nthreads, per_thread = 6, 40
    bufs = []
    try:
        for _ in range(nthreads):
            h_in, h_out = lib.lig_host_alloc(R_ * 16), lib.lig_host_alloc(R_ * 8)
            np.ctypeslib.as_array((C.c_uint8 * (R_ * 16)).from_address(h_in))[:] = reqs.view(np.uint8).reshape(-1)
            bufs.append((h_in, h_out))
        e.upload_snapshot(1, snaps[0].packed)
        stop = threading.Event()
        errors, done = [], [0]

        def refresher():
            ep = 1
            while not stop.is_set():
                ep += 1
                try:
                    e.upload_snapshot(ep, snaps[(ep - 1) % 4].packed)
                except Exception as ex:      # noqa: BLE001
                    errors.append(repr(ex))
                    return
                current[0] = ep

        current = [1]

        def caller(k):
            h_in, h_out = bufs[k]
            view = np.ctypeslib.as_array((C.c_uint8 * (R_ * 8)).from_address(h_out)).view(PICK_DTYPE)
            for _ in range(per_thread):
                ep = current[0]
                try:
                    t = e.schedule_batch_async(ep, 5, h_in, R_, h_out)
                    e.schedule_wait(t)
                except N.LigError as ex:
                    if ex.code == N.LIG_ERR_STALE_EPOCH:
                        continue             # two refreshes raced past this batch: the caller re-resolves
                    errors.append(repr(ex))
                    return
                if not np.array_equal(view, wants[(ep - 1) % 4]):
                    errors.append(f"thread {k}: wrong picks for epoch {ep}")
                    return
                done[0] += 1

        th = [threading.Thread(target=caller, args=(k,)) for k in range(nthreads)]
        rf = threading.Thread(target=refresher)
        rf.start()
        for t in th:
            t.start()
        for t in th:
            t.join()
        stop.set()
        rf.join()
        assert not errors, errors[:3]
        assert done[0] > nthreads * per_thread // 4
        # ticket exhaustion is an error code, not a hang
        tickets = []
        with pytest.raises(N.LigError) as ei:
            for _ in range(N.LIG_MAX_TICKETS + 1):
                tickets.append(e.schedule_batch_async(current[0], 5, bufs[0][0], 16, bufs[0][1]))
        assert ei.value.code == N.LIG_ERR_BUSY and len(tickets) == N.LIG_MAX_TICKETS
        for t in tickets:
            e.schedule_wait(t)
        with pytest.raises(N.LigError):           # pageable buffers are refused by the async form
            e.schedule_batch_async(current[0], 5, reqs.ctypes.data, 16, np.empty(16, dtype=PICK_DTYPE).ctypes.data)
    finally:
        e.close()
        for h_in, h_out in bufs:
            lib.lig_host_free(h_in)
            lib.lig_host_free(h_out)


def test_readers_on_many_streams_vs_device_uploads(oracle):
    """Queues on several caller streams read a slot while device-side uploads keep replacing the
    OTHER slot and then this one: every result must be that of the epoch it was enqueued against
    (a single idle event per slot lost all but the last reader)."""
    import torch
    P_, A_, R_ = 600, 40, 300_000
    snaps = [WL.make_snapshot(P_, A_, seed=90 + i) for i in range(3)]
    blobs = [torch.from_numpy(s.packed.blob()).cuda() for s in snaps]
    reqs = WL.make_requests(R_, A_, seed=99)
    d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).cuda()
    e = Engine(0, max_pods=1024, max_adapters=64, max_batch=R_)
    try:
        want = []
        for i, s in enumerate(snaps):
            e.upload_snapshot(100 + i, s.packed)
            want.append(e.schedule_batch(100 + i, 3, reqs))
        streams = [torch.cuda.Stream() for _ in range(5)]
        up = torch.cuda.Stream()
        outs = [torch.zeros(R_ * 8, dtype=torch.uint8, device="cuda") for _ in range(5 * 6)]
        torch.cuda.synchronize()   # the fills ran on torch's default stream
        expect = []
        ep = 200
        e.upload_snapshot_device(ep, P_, A_, blobs[0].data_ptr(), up.cuda_stream)
        cur = 0
        k = 0
        for it in range(6):
            for si, st in enumerate(streams):
                e.schedule_batches_device(ep, 3, [d_reqs.data_ptr()] * 3, R_, [outs[k].data_ptr()] * 3, st.cuda_stream)
                expect.append(cur)
                k += 1
            # replace the other slot, then the slot the streams above are still reading
            nxt = (cur + 1) % 3
            ep += 1
            e.upload_snapshot_device(ep, P_, A_, blobs[nxt].data_ptr(), up.cuda_stream)
            cur = nxt
        torch.cuda.synchronize()
        for i in range(k):
            # three batches wrote the same buffer with seeds 3,4,5: the last one wins
            got = outs[i].cpu().numpy().view(PICK_DTYPE)
            assert np.array_equal(got["n_survivors"], want[expect[i]]["n_survivors"]), i
            assert np.array_equal(got["status"], want[expect[i]]["status"]), i
            assert (got["pod_idx"] < P_).all()
    finally:
        e.close()


def test_fast_and_general_class_build_agree(monkeypatch):
    """Pools of up to 4096 pods are built by the register-resident fast kernel, larger ones (or
    LIG_FAST_BUILD=0) by the general one: every class table and every pick must be identical."""
    cases = [(WL.make_snapshot(P, A, seed=100 + P).packed, A) for P, A in
             [(1, 3), (31, 3), (33, 5), (100, 17), (512, 256), (1000, 40), (4095, 64), (4096, 1024)]]
    pools = [pack_pod_metrics(EDGE_POOLS[k]) for k in ("all_nan_kv", "q_all_high", "drop_all", "ties", "overfull", "int32_extremes")]
    cases += [(p, p.A) for p in pools]
    monkeypatch.setenv("LIG_FAST_BUILD", "0")
    general = Engine(0, max_pods=4096, max_adapters=1024, max_batch=1 << 16)
    monkeypatch.setenv("LIG_FAST_BUILD", "1")
    fast = Engine(0, max_pods=4096, max_adapters=1024, max_batch=1 << 16)
    try:
        for ep, (packed, A) in enumerate(cases, start=1):
            general.upload_snapshot(ep, packed)
            fast.upload_snapshot(ep, packed)
            reqs = np.concatenate([all_class_requests(A, extra_ids=(-1, A + 3)), WL.make_requests(20000, A, seed=ep)])
            reqs = np.ascontiguousarray(reqs)
            assert np.array_equal(general.schedule_batch(ep, 9, reqs), fast.schedule_batch(ep, 9, reqs)), (packed.P, A)
            step = max(1, (A + 1) // 64)
            for crit in (False, True):
                for a in list(range(0, A + 1, step)) + [A]:
                    g, f_ = general.read_class(ep, crit, a, packed.P), fast.read_class(ep, crit, a, packed.P)
                    assert g[0] == f_[0] and g[1] == f_[1] and g[2].tolist() == f_[2].tolist(), (packed.P, A, crit, a)
            # the resident queue kernel reads the COMPACT tables: same answers from both builds
            import torch
            d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).cuda()
            outs = []
            for e in (general, fast):
                d_out = [torch.zeros(len(reqs) * 8, dtype=torch.uint8, device="cuda") for _ in range(2)]
                e.schedule_batches_device(ep, 9, [d_reqs.data_ptr()] * 2, len(reqs), [t.data_ptr() for t in d_out], 0)
                torch.cuda.synchronize()
                outs.append(d_out[1].cpu().numpy().view(PICK_DTYPE))
                assert e.pick_kernel_info(ep)["tables_in_smem"]
            assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], general.schedule_batch(ep, 10, reqs))
    finally:
        general.close()
        fast.close()
