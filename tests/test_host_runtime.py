"""The native C++ host runtime (lig::scheduling::Scheduler, csrc/host/) above the C ABI."""
import numpy as np
import pytest

from helpers import golden_to_podmetrics
from llm_instance_gateway_b200 import host as H
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200.backend import Metrics, Pod, PodMetrics


def test_host_library_loads_and_exports():
    lib = H.load()
    for name in H.EXPORTED_SYMBOLS:
        assert hasattr(lib, name)


@pytest.mark.skipif(N.load().lig_device_count() > 0, reason="checks the no-GPU failure mode")
def test_host_scheduler_fails_loudly_without_gpu():
    prov = H.HostProvider([PodMetrics(Pod("p", "a"), Metrics())])
    with pytest.raises(H.HostSchedulerError) as ei:
        H.HostScheduler(prov, max_pods=8, max_adapters=8, max_batch=8)
    assert "no CPU path" in str(ei.value)
    prov.close()


def snapshot_to_podmetrics(snap):
    p = snap.packed
    return [PodMetrics(Pod(f"pod-{i}", f"address-{i}"),
                       Metrics(WaitingQueueSize=int(snap.q64[i]), KVCacheUsagePercent=float(p.kv[i]),
                               MaxActiveModels=int(snap.max_active64[i]),
                               ActiveModels={WL.adapter_name(a): 1 for a in snap.active[i]}))
            for i in range(p.P)]


@pytest.mark.gpu
def test_golden_vectors_through_cpp_scheduler(golden):
    for case in golden["TestFilter"]:
        if case["filter"]["name"] != "defaultFilter":
            continue
        prov = H.HostProvider([golden_to_podmetrics(p) for p in case["input"]])
        s = H.HostScheduler(prov, max_pods=64, max_adapters=64, max_batch=256)
        req = case["req"]
        code, pod, err = s.Schedule(req["model"], req["resolved_target_model"], req["critical"])
        if case["err"]:
            assert code == H.GRPC_RESOURCE_EXHAUSTED and pod is None
            assert err == ("failed to apply filter, resulted 0 pods, this should never happen: rpc error: "
                           "code = ResourceExhausted desc = dropping request due to limited backend resources")
        else:
            assert code == H.GRPC_OK and pod.Name == case["output"][0]["name"], case["name"]
        s.close()
        prov.close()
    case = golden["TestHandleRequestBody"][0]
    prov = H.HostProvider([golden_to_podmetrics(p) for p in case["pods"]])
    s = H.HostScheduler(prov, max_pods=64, max_adapters=64, max_batch=256)
    resolved = case["models"][case["request_model"]]["target_models"][0]["name"]
    code, pod, _ = s.Schedule(case["request_model"], resolved, False)
    assert code == 0 and pod.Address == "address-1" and pod.Name == "pod-1"
    s.close()
    prov.close()


@pytest.mark.gpu
def test_concurrent_callers_are_batched_and_correct(oracle):
    P, A = 300, 24
    snap = WL.make_snapshot(P, A, seed=31)
    prov = H.HostProvider(snapshot_to_podmetrics(snap))
    s = H.HostScheduler(prov, max_pods=512, max_adapters=64, max_batch=4096, flush_size=256, batch_window_us=200)
    models = [WL.adapter_name(a) for a in range(A)] + [WL.UNKNOWN_MODEL]
    models = models + models
    critical = [False] * (A + 1) + [True] * (A + 1)
    codes, pods = s.schedule_concurrent(64, 200, models, critical)
    pool = oracle.Pool(snap.pod_records())
    want = {}
    for m, c in set(zip(models, critical)):
        want[(m, c)] = pool.filter(m, c)
    n = len(codes)
    for i in range(n):
        m, c = models[i % len(models)], critical[i % len(models)]
        rc, survivors = want[(m, c)]
        if rc == oracle.LIGO_OK:
            assert codes[i] == H.GRPC_OK and pods[i] in survivors, (i, m, c)
        elif rc == oracle.LIGO_DROP:
            assert codes[i] == H.GRPC_RESOURCE_EXHAUSTED and pods[i] == -1
        else:
            assert codes[i] == H.GRPC_UNKNOWN
    st = s.stats()
    assert st["scheduled"] == n and st["batches"] < n / 4 and st["max_batch"] > 8, st
    # every survivor of a multi-survivor class gets picked eventually (uniform draw)
    big = max(want.items(), key=lambda kv: len(kv[1][1]))
    if len(big[1][1]) > 1:
        sel = [pods[i] for i in range(n) if (models[i % len(models)], critical[i % len(models)]) == big[0]]
        assert set(sel) == set(big[1][1])
    s.close()
    prov.close()


@pytest.mark.gpu
def test_refresh_follows_the_provider():
    mk = lambda q0, q1: [PodMetrics(Pod("pod-0", "address-0"), Metrics(WaitingQueueSize=q0, KVCacheUsagePercent=0.1)),
                         PodMetrics(Pod("pod-1", "address-1"), Metrics(WaitingQueueSize=q1, KVCacheUsagePercent=0.1))]
    prov = H.HostProvider(mk(0, 30))
    s = H.HostScheduler(prov, max_pods=8, max_adapters=8, max_batch=64)
    assert s.Schedule("m", "m", True)[1].Name == "pod-0"
    prov.set_pods(mk(30, 0))
    assert s.Schedule("m", "m", True)[1].Name == "pod-0"      # snapshot not refreshed yet
    s.Refresh()
    assert s.Schedule("m", "m", True)[1].Name == "pod-1"
    prov.set_pods(mk(30, 30))                                   # nobody has capacity for sheddable
    s.Refresh()
    assert s.Schedule("m", "m", False)[0] == H.GRPC_RESOURCE_EXHAUSTED
    assert s.Schedule("m", "m", True)[0] == H.GRPC_OK
    assert s.stats()["refreshes"] == 3
    s.close()
    # background refresher thread
    prov.set_pods(mk(0, 30))
    s = H.HostScheduler(prov, max_pods=8, max_adapters=8, max_batch=64, refresh_interval_ms=5)
    prov.set_pods(mk(30, 0))
    import time
    time.sleep(0.1)
    assert s.Schedule("m", "m", True)[1].Name == "pod-1"
    s.close()
    prov.close()


@pytest.mark.gpu
def test_reference_load_test_shape(oracle):
    """The reference's ghz load generator (pkg/ext-proc/test/benchmark/benchmark.go:20-29,45-110):
    200 fake pods, 5 adapters each ("adapter-<pod*5+j>"), all other metrics zero, requests cycling
    over the 1000 adapter names, no Criticality.  With MaxActiveModels == 0 nobody has room, so the
    only low-cost pod is the one holding the adapter: request n must land on pod (n % 1000) // 5."""
    n_pods, per_pod = 200, 5
    pods = [PodMetrics(Pod(f"pod-{i}", f"address-{i}"),
                       Metrics(ActiveModels={f"adapter-{i * per_pod + j}": 0 for j in range(per_pod)}))
            for i in range(n_pods)]
    prov = H.HostProvider(pods)
    s = H.HostScheduler(prov, max_pods=256, max_adapters=1024, max_batch=4096, flush_size=512,
                        batch_window_us=20, busy_poll=True, caller_spin_us=50)
    models = [f"adapter-{m}" for m in range(n_pods * per_pod)]
    codes, picked = s.schedule_concurrent(50, 400, models, [False] * len(models))
    assert (codes == H.GRPC_OK).all()
    want = np.array([(i % len(models)) // per_pod for i in range(len(codes))])
    assert np.array_equal(picked, want)
    # and the oracle agrees on a sample of the same requests
    pool = oracle.Pool([dict(name=p.Pod.Name, address=p.Pod.Address, waiting_queue_size=0,
                             kv_cache_usage_percent=0.0, max_active_models=0,
                             active_models=list(p.Metrics.ActiveModels)) for p in pods])
    for m in (0, 4, 5, 499, 999):
        assert pool.filter(models[m], False) == (oracle.LIGO_OK, [m // per_pod])
    st = s.stats()
    assert st["scheduled"] == len(codes) and st["batches"] < len(codes)
    s.close()
    prov.close()


@pytest.mark.gpu
def test_adapter_churn_across_refreshes(oracle):
    """The packer keeps adapter ids stable between refresh ticks and re-interns from scratch when
    churn outgrows max_adapters; picks must follow every snapshot exactly."""
    rng = np.random.default_rng(3)
    prov = H.HostProvider([])
    s = H.HostScheduler(prov, max_pods=64, max_adapters=12, max_batch=256)
    universe = [f"lora-{i}" for i in range(40)]
    for tick in range(12):
        live = list(rng.choice(universe, size=8, replace=False))          # 8 of 40 names alive per tick
        pods = [PodMetrics(Pod(f"pod-{i}", f"address-{i}"),
                           Metrics(WaitingQueueSize=int(rng.integers(0, 60)), KVCacheUsagePercent=float(np.round(rng.random(), 2)),
                                   MaxActiveModels=int(rng.integers(0, 4)),
                                   ActiveModels={a: 1 for a in rng.choice(live, size=int(rng.integers(0, 4)), replace=False)}))
                for i in range(int(rng.integers(1, 40)))]
        prov.set_pods(pods)
        s.Refresh()
        pool = oracle.Pool([dict(name=p.Pod.Name, address=p.Pod.Address, waiting_queue_size=p.Metrics.WaitingQueueSize,
                                 kv_cache_usage_percent=p.Metrics.KVCacheUsagePercent, max_active_models=p.Metrics.MaxActiveModels,
                                 active_models=list(p.Metrics.ActiveModels)) for p in pods])
        for model in live[:4] + ["never-loaded", universe[0]]:
            for crit in (False, True):
                rc, survivors = pool.filter(model, crit)
                code, pod, _ = s.Schedule(model, model, crit)
                if rc == oracle.LIGO_OK:
                    assert code == H.GRPC_OK and int(pod.Name.split("-")[1]) in survivors, (tick, model, crit)
                elif rc == oracle.LIGO_DROP:
                    assert code == H.GRPC_RESOURCE_EXHAUSTED
                else:
                    assert code == H.GRPC_UNKNOWN
    assert s.stats()["refreshes"] == 13
    s.close()
    prov.close()


@pytest.mark.gpu
def test_delta_refresh_small_churn_matches_oracle(oracle):
    """Tick after tick ~1 % of the pods change: the runtime uploads only the dirty pods
    (lig_update_snapshot) and every pick still follows the current snapshot exactly."""
    P, A = 1000, 40
    snap = WL.make_snapshot(P, A, seed=71)
    pods = snapshot_to_podmetrics(snap)
    prov = H.HostProvider(pods)
    s = H.HostScheduler(prov, max_pods=1024, max_adapters=64, max_batch=1024)
    rng = np.random.default_rng(72)
    models = [WL.adapter_name(a) for a in range(0, A, 5)] + [WL.UNKNOWN_MODEL]
    for tick in range(8):
        for i in rng.choice(P, size=10, replace=False):
            m = pods[i].Metrics
            pods[i] = PodMetrics(pods[i].Pod, Metrics(
                WaitingQueueSize=int(rng.integers(0, 70)), KVCacheUsagePercent=float(np.round(rng.random(), 3)),
                MaxActiveModels=m.MaxActiveModels,
                ActiveModels={WL.adapter_name(a): 1 for a in rng.choice(A, size=int(rng.integers(0, 4)), replace=False)}))
        prov.set_pods(pods)
        s.Refresh()
        st = s.stats()
        assert st["last_dirty_pods"] <= 10 and st["delta_refreshes"] == tick + 1, st
        pool = oracle.Pool([dict(name=p.Pod.Name, address=p.Pod.Address, waiting_queue_size=p.Metrics.WaitingQueueSize,
                                 kv_cache_usage_percent=p.Metrics.KVCacheUsagePercent, max_active_models=p.Metrics.MaxActiveModels,
                                 active_models=list(p.Metrics.ActiveModels)) for p in pods])
        for model in models:
            for crit in (False, True):
                rc, survivors = pool.filter(model, crit)
                code, pod, _ = s.Schedule(model, model, crit)
                if rc == oracle.LIGO_OK:
                    assert code == H.GRPC_OK and int(pod.Name.split("-")[1]) in survivors, (tick, model, crit)
                else:
                    assert code == (H.GRPC_RESOURCE_EXHAUSTED if rc == oracle.LIGO_DROP else H.GRPC_UNKNOWN)
    # a tick that changes most pods goes through the full upload again
    prov.set_pods(snapshot_to_podmetrics(WL.make_snapshot(P, A, seed=73)))
    s.Refresh()
    assert s.stats()["delta_refreshes"] == 8 and s.stats()["last_dirty_pods"] > P // 4
    s.close()
    prov.close()


@pytest.mark.gpu
def test_unrepresentable_pod_is_excluded_not_fatal():
    """One pod reporting a WaitingQueueSize beyond int32 must not freeze the pool on a stale
    snapshot: it is left out of the snapshot and counted."""
    mk = lambda q1: [PodMetrics(Pod("pod-0", "address-0"), Metrics(WaitingQueueSize=3, KVCacheUsagePercent=0.1)),
                     PodMetrics(Pod("pod-1", "address-1"), Metrics(WaitingQueueSize=q1, KVCacheUsagePercent=0.0)),
                     PodMetrics(Pod("pod-2", "address-2"), Metrics(WaitingQueueSize=40, KVCacheUsagePercent=0.1))]
    prov = H.HostProvider(mk(0))
    s = H.HostScheduler(prov, max_pods=8, max_adapters=8, max_batch=64)
    assert s.Schedule("m", "m", True)[1].Name == "pod-1"
    prov.set_pods(mk(2**40))
    s.Refresh()                                              # must not raise
    st = s.stats()
    assert st["excluded_pods"] == 1 and st["failed_refreshes"] == 0
    for _ in range(10):
        assert s.Schedule("m", "m", True)[1].Name == "pod-0"   # pod indices still map to the right pods
    prov.set_pods(mk(1))
    s.Refresh()
    assert s.stats()["excluded_pods"] == 0 and s.Schedule("m", "m", True)[1].Name == "pod-1"
    s.close()
    prov.close()


@pytest.mark.gpu
def test_schedule_model_is_the_resolve_step_plus_schedule(golden):
    """handlers/request.go:42-56 in the C++ runtime: the hermetic case (test/hermetic_test.go:37-104),
    a weighted split, an unknown model, a shed request."""
    from llm_instance_gateway_b200.backend import CRITICAL, InferenceModel, InferenceModelSpec, TargetModel
    case = golden["TestHandleRequestBody"][0]
    prov = H.HostProvider([golden_to_podmetrics(p) for p in case["pods"]])
    s = H.HostScheduler(prov, max_pods=64, max_adapters=64, max_batch=256, seed=5)
    ds = H.HostDataStore([
        InferenceModel("my-model", InferenceModelSpec(ModelName="my-model", TargetModels=[TargetModel("my-model-v1", 100)])),
        InferenceModel("split", InferenceModelSpec(ModelName="split", Criticality=CRITICAL,
                                                   TargetModels=[TargetModel("canary", 25), TargetModel("v1.1", 55), TargetModel("v1", 50)])),
        InferenceModel("plain", InferenceModelSpec(ModelName="plain")),
        InferenceModel("zero", InferenceModelSpec(ModelName="zero", TargetModels=[TargetModel("x", 0)]))])
    code, resolved, pod, err = s.ScheduleModel(ds, "my-model")
    assert (code, resolved, pod.Address) == (H.GRPC_OK, "my-model-v1", "address-1")
    seen = {}
    for _ in range(600):
        code, resolved, pod, _ = s.ScheduleModel(ds, "split")
        assert code == H.GRPC_OK
        seen[resolved] = seen.get(resolved, 0) + 1
    assert set(seen) == {"canary", "v1.1", "v1"} and seen["v1.1"] > seen["canary"]       # weights 25 / 55 / 50
    code, resolved, pod, err = s.ScheduleModel(ds, "plain")
    assert code == H.GRPC_OK and resolved == "plain"
    code, _, _, err = s.ScheduleModel(ds, "nope")
    assert code == H.GRPC_UNKNOWN and err == "error finding a model object in InferenceModel for input nope"
    code, _, _, err = s.ScheduleModel(ds, "zero")
    assert code == H.GRPC_UNKNOWN and err == "error getting target model name for model zero"
    # shed: ResourceExhausted survives the wrapping (-> 429, handlers/server.go:97-109)
    prov.set_pods([PodMetrics(Pod("pod-0", "address-0"), Metrics(WaitingQueueSize=10, KVCacheUsagePercent=0.9))])
    s.Refresh()
    code, _, _, err = s.ScheduleModel(ds, "plain")
    assert code == H.GRPC_RESOURCE_EXHAUSTED and err.startswith("failed to find target pod: failed to apply filter")
    ds.close()
    s.close()
    prov.close()


@pytest.mark.gpu
def test_one_scheduler_over_several_gpus(oracle):
    """Options::devices: the C++ runtime drives every GPU of the box through lig_group_* — what the
    Go adapter does with one process per ext-proc (main.go:137)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    P, A = 300, 24
    snap = WL.make_snapshot(P, A, seed=81)
    prov = H.HostProvider(snapshot_to_podmetrics(snap))
    s = H.HostScheduler(prov, max_pods=512, max_adapters=64, max_batch=4096, flush_size=512, batch_window_us=300,
                        devices=list(range(min(n, 4))))
    models = [WL.adapter_name(a) for a in range(A)] + [WL.UNKNOWN_MODEL]
    models = models + models
    critical = [False] * (A + 1) + [True] * (A + 1)
    codes, pods = s.schedule_concurrent(32, 200, models, critical)
    pool = oracle.Pool(snap.pod_records())
    for i in range(len(codes)):
        rc, survivors = pool.filter(models[i % len(models)], critical[i % len(models)])
        if rc == oracle.LIGO_OK:
            assert codes[i] == H.GRPC_OK and pods[i] in survivors, i
        else:
            assert codes[i] == (H.GRPC_RESOURCE_EXHAUSTED if rc == oracle.LIGO_DROP else H.GRPC_UNKNOWN)
    assert s.stats()["max_batch"] > 8
    s.close()
    prov.close()
