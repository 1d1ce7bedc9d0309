"""GPU parity of the request pre-step fused into the pick: model id ->
InferenceModel lookup, weighted target draw, criticality, then Schedule — against the oracle's
restatement of handlers/request.go:42-56 + backend/datastore.go:70-105 (oracle/lig_oracle_models.c)
followed by the port of Scheduler.Schedule.  Bit-exact: (status, pod_idx, target_idx) per request."""
import numpy as np
import pytest

from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.backend import CRITICAL, InferenceModel, InferenceModelSpec, TargetModel
from llm_instance_gateway_b200.engine import Engine
from llm_instance_gateway_b200.packer import MPICK_DTYPE, REQ_DTYPE, pack_models

pytestmark = pytest.mark.gpu


def setup(oracle, cfg, snap_seed=WL.SNAPSHOT_SEED):
    c = WL.CONFIGS[cfg]
    snap = WL.make_snapshot(c["P"], c["A"], seed=snap_seed)
    models = WL.make_models(c["A"])
    pm = pack_models(models, snap.packed)
    return c, snap, models, pm, oracle.Pool(snap.pod_records()), oracle.Models(WL.model_records(models))


@pytest.mark.parametrize("cfg", ["C2", "C3", "C5"])
def test_models_batch_matches_oracle(cfg, oracle):
    c, snap, models, pm, pool, mo = setup(oracle, cfg)
    R = min(c["R"], 30000)
    ids = WL.make_model_requests(R, c["A"], seed=5)
    ids[:4] = [pm.n_models, pm.n_models + 5, 2**32 - 1, c["A"]]       # no such model x3, the pass-through model
    ids[4:12] = c["A"] + 1 + np.arange(8, dtype=np.uint32)          # weighted-split models
    with Engine(0, max_pods=c["P"], max_adapters=c["A"], max_batch=1 << 16) as e:
        e.upload_snapshot(1, snap.packed)
        with pytest.raises(N.LigError) as ei:                       # no model table yet
            e.schedule_models_batch(1, 9, ids)
        assert ei.value.code == N.LIG_ERR_NO_SNAPSHOT
        e.upload_models(1, pm)
        want = mo.schedule_batch(pool, ids, 9, first_index=77)
        got = e.schedule_models_batch(1, 9, ids, first_index=77)
        for f in ("status", "pod_idx", "target_idx"):
            bad = np.nonzero(got[f] != want[f])[0]
            assert bad.size == 0, (f, bad[:5], got[bad[:5]], want[bad[:5]], ids[bad[:5]])
        assert (got["status"] == N.LIG_NO_MODEL).any() and (got["target_idx"] == 255).any()
        if R >= 10000:
            assert ((got["target_idx"] > 0) & (got["target_idx"] < 255)).any()     # a non-first target was drawn
        # the descriptor the device built == the oracle's resolve; feeding it to the descriptor
        # path gives the same picks
        reqs, res = e.resolve_models(1, 9, ids, first_index=77)
        ok = res["status"] == N.LIG_OK
        assert np.array_equal(ok, want["status"] != N.LIG_NO_MODEL)
        assert np.array_equal(reqs["rand_key"], 77 + np.arange(R, dtype=np.uint64))
        for i in range(0, R, 97):
            if not ok[i]:
                continue
            rc, name, crit, k = mo.resolve(int(ids[i]), 9, 77 + i)
            assert rc == 0 and reqs[i]["flags"] == int(crit) and res[i]["target_idx"] == k
            assert int(reqs[i]["adapter_id"]) == snap.packed.adapter_id(name)
        picks = e.schedule_batch(1, 9, np.ascontiguousarray(reqs[ok]))
        assert np.array_equal(picks["pod_idx"], got["pod_idx"][ok].astype(np.int32))
        assert np.array_equal(picks["status"], got["status"][ok])
        # a shard with its offset gives the single-call result (counter-based keys)
        lo = R // 3
        assert np.array_equal(e.schedule_models_batch(1, 9, np.ascontiguousarray(ids[lo:]), first_index=77 + lo), got[lo:])
        # a new snapshot in the same slot invalidates the model table (it is interned against the old one)
        e.upload_snapshot(2, snap.packed)
        e.upload_snapshot(3, snap.packed)                            # evicts epoch 1
        with pytest.raises(N.LigError):
            e.schedule_models_batch(3, 9, ids)
        e.upload_models(3, pm)
        assert np.array_equal(e.schedule_models_batch(3, 9, ids, first_index=77), got)


@pytest.mark.parametrize("env", [{}, {"LIG_TAB_SMEM": "0"}, {"LIG_MODELS_GROUPS": "1", "LIG_MODELS_STAGES": "2"},
                                 {"LIG_MODELS_STAGES": "8"}, {"LIG_PICK_KERNEL": "merged"}],
                         ids=["default", "tables_global", "1group_2stage", "8stage", "plain_kernel"])
def test_models_device_queue_matches_host_call(env, oracle, monkeypatch):
    import torch
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c, snap, models, pm, pool, mo = setup(oracle, "C3", snap_seed=17)
    e = Engine(0, max_pods=c["P"], max_adapters=c["A"], max_batch=1 << 18)
    try:
        e.upload_snapshot(1, snap.packed)
        e.upload_models(1, pm)
        nb = 5
        stream = torch.cuda.Stream()
        for R in (100_000, 4096, 1024, 8, 100_003, 5):                # last two: not a multiple of 4 -> plain kernel
            host = [WL.make_model_requests(R, c["A"], seed=30 + b) for b in range(nb)]
            d_ids = [torch.from_numpy(h.view(np.uint8)).cuda() for h in host]
            d_out = [torch.zeros(R * 4, dtype=torch.uint8, device="cuda") for _ in range(nb)]
            torch.cuda.synchronize()   # the fills / copies ran on torch's default stream
            with torch.cuda.stream(stream):
                e.schedule_models_batches_device(1, 40, 1 << 33, [t.data_ptr() for t in d_ids], R,
                                                 [t.data_ptr() for t in d_out], stream.cuda_stream)
            stream.synchronize()
            for b in range(nb):
                got = d_out[b].cpu().numpy().view(MPICK_DTYPE)
                want = e.schedule_models_batch(1, 40 + b, host[b], first_index=1 << 33)
                assert np.array_equal(got, want), (env, R, b)
                if R == 100_000 and b == 0:
                    assert np.array_equal(want, mo.schedule_batch(pool, host[b], 40, first_index=1 << 33))
    finally:
        e.close()


def test_model_table_validation_and_edge_tables(oracle):
    snap = WL.make_snapshot(40, 4, seed=2)
    P = snap.packed
    mk = lambda name, crit, tms: InferenceModel(name, InferenceModelSpec(ModelName=name, Criticality=crit,
                                                                         TargetModels=[TargetModel(n, w) for n, w in tms]))
    models = [mk("plain", None, []),
              mk("one", CRITICAL, [(WL.adapter_name(1), 5)]),
              mk("zero-sum", None, [(WL.adapter_name(0), 0)]),          # packed as absent (Int31n(0) panics in Go)
              mk("pow2", CRITICAL, [(WL.adapter_name(0), 1), (WL.adapter_name(1), 1), (WL.adapter_name(2), 2)]),
              mk("big", None, [(WL.adapter_name(3), 2**30), (WL.adapter_name(2), 2**30 - 1)]),
              mk("zeros-inside", None, [(WL.adapter_name(0), 0), (WL.adapter_name(1), 3), ("nowhere", 0), (WL.adapter_name(2), 4)])]
    pm = pack_models(models, P)
    assert pm.present.tolist() == [1, 1, 0, 1, 1, 1]
    recs = WL.model_records(models)
    recs[2] = None
    mo, pool = oracle.Models(recs), oracle.Pool(snap.pod_records())
    ids = np.tile(np.arange(8, dtype=np.uint32), 4000)
    with Engine(0, max_pods=64, max_adapters=4, max_batch=1 << 16) as e:
        e.upload_snapshot(1, P)
        e.upload_models(1, pm)
        got = e.schedule_models_batch(1, 3, ids)
        want = mo.schedule_batch(pool, ids, 3)
        assert np.array_equal(got, want)
        assert (got["status"][ids == 2] == N.LIG_NO_MODEL).all() and (got["status"][ids >= 6] == N.LIG_NO_MODEL).all()
        t = got["target_idx"][ids == 5]
        assert set(t.tolist()) == {1, 3}                                  # zero-weight targets are never drawn
        t = got["target_idx"][ids == 3]
        frac = np.bincount(t, minlength=3) / len(t)
        assert np.abs(frac - np.array([0.25, 0.25, 0.5])).max() < 0.03
        # refused tables
        bad = pack_models(models, P)
        bad.target_weights = bad.target_weights.copy()
        bad.target_weights[0] = -1
        with pytest.raises(N.LigError) as ei:
            e.upload_models(1, bad)
        assert ei.value.code == N.LIG_ERR_RANGE
        # a refused upload changes nothing: the previous table keeps serving
        assert np.array_equal(e.schedule_models_batch(1, 3, ids), want)


def test_async_uploads_are_ordered_on_the_device(oracle):
    """lig_upload_snapshot_async / lig_upload_models_async return before the device has copied or
    built anything; a schedule call issued right behind them must still see exactly that snapshot
    and model table (device-side ordering), the previous epoch must stay servable, and the pinned
    staging copies must survive back-to-back async uploads of the same slot."""
    c = WL.CONFIGS["C3"]
    R = 20000
    ids = WL.make_model_requests(R, c["A"], seed=15)
    reqs = WL.make_requests(R, c["A"], seed=16)
    models = WL.make_models(c["A"])
    snaps = [WL.make_snapshot(c["P"], c["A"], seed=300 + i) for i in range(6)]
    pools = [oracle.Pool(s.pod_records()) for s in snaps]
    mo = oracle.Models(WL.model_records(models))
    with Engine(0, max_pods=c["P"], max_adapters=c["A"], max_batch=1 << 16) as e:
        for i, s in enumerate(snaps):
            ep = 50 + i
            e.upload_snapshot(ep, s.packed, block=False)
            e.upload_models(ep, pack_models(models, s.packed), block=False)
            got = e.schedule_models_batch(ep, 7 + i, ids, first_index=1000 * i)
            want = mo.schedule_batch(pools[i], ids, 7 + i, first_index=1000 * i)
            assert np.array_equal(got, want), i
            picks = e.schedule_batch(ep, 3, reqs)
            wantp, _ = pools[i].schedule_batch(s.adapter_names(), WL.UNKNOWN_MODEL, reqs, 3)
            assert np.array_equal(picks, wantp), i
            if i:       # the previous epoch is still resident and unchanged
                prev, _ = pools[i - 1].schedule_batch(snaps[i - 1].adapter_names(), WL.UNKNOWN_MODEL, reqs, 3)
                assert np.array_equal(e.schedule_batch(ep - 1, 3, reqs), prev), i
        # the same epoch re-uploaded many times in a row without any synchronising call in between
        # (each async upload re-stages the slot's pinned blob: the staging guard is what is tested)
        for k in range(12):
            s = snaps[k % 6]
            e.upload_snapshot(99, s.packed, block=False)
            e.upload_models(99, pack_models(models, s.packed), block=False)
        last = snaps[11 % 6]
        got = e.schedule_models_batch(99, 5, ids)
        assert np.array_equal(got, mo.schedule_batch(pools[11 % 6], ids, 5, first_index=0))
        # the device queue path right behind an async upload (header not back yet -> strided tables)
        import torch
        e.upload_snapshot(101, snaps[2].packed, block=False)
        d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).cuda()
        d_out = torch.zeros(R * 8, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        e.upload_snapshot(102, snaps[3].packed, block=False)
        e.schedule_batches_device(102, 3, [d_reqs.data_ptr()], R, [d_out.data_ptr()], st.cuda_stream)
        st.synchronize()
        wantp, _ = pools[3].schedule_batch(snaps[3].adapter_names(), WL.UNKNOWN_MODEL, reqs, 3)
        from llm_instance_gateway_b200.packer import PICK_DTYPE
        assert np.array_equal(d_out.cpu().numpy().view(PICK_DTYPE), wantp)
