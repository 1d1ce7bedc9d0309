"""Delta snapshot upload (device half): lig_update_snapshot(new, base, dirty pods)
must leave exactly the tables a full lig_upload_snapshot of the modified pool would build."""
import numpy as np
import pytest

from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.engine import Engine
from llm_instance_gateway_b200.packer import pack_columns

pytestmark = pytest.mark.gpu


def mutate(snap, frac, seed):
    """A copy of the snapshot with ~frac of the pods changed (metrics and ActiveModels)."""
    rng = np.random.default_rng(seed)
    p = snap.packed
    P, A = p.P, p.A
    dirty = np.sort(rng.choice(P, size=max(1, int(P * frac)), replace=False))
    kv, q, ma = p.kv.copy(), snap.q64.copy(), snap.max_active64.copy()
    active = [list(a) for a in snap.active]
    for i in dirty:
        kv[i] = float(np.round(rng.random(), 3))
        q[i] = int(rng.integers(0, 80))
        ma[i] = int(rng.choice([0, 4, 8, 16]))
        active[i] = sorted(rng.choice(A, size=int(rng.integers(0, min(A, 6) + 1)), replace=False).tolist()) if A else []
    W = (P + 31) // 32
    bm = np.zeros((A, W), dtype=np.uint32)
    for pod, acts in enumerate(active):
        for a in acts:
            bm[a, pod >> 5] |= np.uint32(1 << (pod & 31))
    packed = pack_columns(kv, q, [len(a) for a in active], ma, bm)
    off = np.zeros(len(dirty) + 1, dtype=np.int32)
    ids = []
    for k, i in enumerate(dirty):
        ids.extend(active[i])
        off[k + 1] = len(ids)
    delta = dict(pod_idx=dirty.astype(np.int32), kv=packed.kv[dirty], q=packed.q[dirty], n_active=packed.n_active[dirty],
                 max_active=packed.max_active[dirty], adapter_offsets=off, adapter_ids=np.array(ids, dtype=np.int32))
    records = [dict(name=f"pod-{i}", address=f"address-{i}", waiting_queue_size=int(q[i]), kv_cache_usage_percent=float(kv[i]),
                    max_active_models=int(ma[i]), active_models=[WL.adapter_name(a) for a in active[i]]) for i in range(P)]
    return packed, delta, records


@pytest.mark.parametrize("cfg,frac", [("C2", 0.1), ("C3", 0.01), ("C4", 0.01), ("C5", 0.5)])
def test_delta_equals_full_upload(cfg, frac, oracle):
    c = WL.CONFIGS[cfg]
    snap = WL.make_snapshot(c["P"], c["A"], seed=31)
    packed2, delta, records2 = mutate(snap, frac, seed=32)
    reqs = WL.make_requests(min(c["R"], 50_000), c["A"], seed=33)
    with Engine(0, max_pods=c["P"], max_adapters=c["A"], max_batch=1 << 16) as e:
        e.upload_snapshot(1, snap.packed)
        before = e.schedule_batch(1, 5, reqs)
        e.update_snapshot(2, 1, **delta)
        got = e.schedule_batch(2, 5, reqs)
        assert np.array_equal(e.schedule_batch(1, 5, reqs), before)            # the base epoch is untouched
        e.upload_snapshot(3, packed2)                                          # evicts epoch 1
        want = e.schedule_batch(3, 5, reqs)
        assert np.array_equal(got, want) and not np.array_equal(got, before)
        for crit in (False, True):                                             # table by table, including never-requested classes
            for a in range(0, c["A"] + 1, max(1, c["A"] // 37)):
                assert [x.tolist() if hasattr(x, "tolist") else x for x in e.read_class(2, crit, a, c["P"])] == \
                       [x.tolist() if hasattr(x, "tolist") else x for x in e.read_class(3, crit, a, c["P"])], (crit, a)
        # and against the oracle on the modified pool
        ref, _ = oracle.Pool(records2).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL,
                                                      np.ascontiguousarray(reqs[:8192]), 5, False, oracle.hardware_threads())
        assert np.array_equal(got[:8192], ref)
        # chains of deltas: 3 -> 4 -> 5, each the base of the next
        packed3, delta3, _ = mutate(type(snap)(packed=packed2, active=[[a for a in range(c["A"]) if (int(packed2.bitmap[a, p >> 5]) >> (p & 31)) & 1]
                                                                          for p in range(c["P"])] if c["A"] * c["P"] < 200_000 else snap.active,
                                                 max_active64=packed2.max_active.astype(np.int64), q64=packed2.q.astype(np.int64)), 0.02, seed=34)
        if c["A"] * c["P"] < 200_000:
            e.update_snapshot(4, 3, **delta3)
            e.upload_snapshot(5, packed3)
            assert np.array_equal(e.schedule_batch(4, 6, reqs), e.schedule_batch(5, 6, reqs))
        # errors: unknown base, base that would be evicted, bad pod index
        with pytest.raises(N.LigError) as ei:
            e.update_snapshot(9, 77, **delta)
        assert ei.value.code == N.LIG_ERR_STALE_EPOCH
        bad = dict(delta)
        bad["pod_idx"] = delta["pod_idx"].copy()
        bad["pod_idx"][0] = c["P"]
        with pytest.raises(N.LigError):
            e.update_snapshot(9, 3 if c["A"] * c["P"] >= 200_000 else 5, **bad)
