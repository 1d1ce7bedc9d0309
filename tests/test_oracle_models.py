"""CPU tests of the oracle's restatement of the step before Schedule (handlers/request.go:42-56,
backend/datastore.go:70-105) and of the class-table CPU arm.

Draw parity with Go's seeded source is UNPINNED (oracle/lig_oracle_models.c header): the
reference's own known answers (datastore_test.go:78-89, seed 420) need Go's rngCooked table.  What
is pinned here: the selection loop for every possible randomVal on the reference's three weight
tables, the distribution of the defined draw, and the resolve semantics around it."""
import numpy as np
import pytest

from llm_instance_gateway_b200 import workload as WL
from oracle import lig_oracle_py as PY

# backend/datastore_test.go:9-76: the three tables, with the target the reference expects for its
# seed-420 draw (kept to show which branch of the loop each table exercises)
GO_TABLES = [([("canary", 50), ("v1", 50)], "canary"),
             ([("canary", 25), ("v1.1", 55), ("v1", 50)], "v1"),
             ([("canary", 20), ("v1.1", 20), ("v1", 10)], "v1.1")]


def go_loop(targets, random_val):
    """datastore.go:91-97, verbatim semantics."""
    for i, (_, w) in enumerate(targets):
        if random_val < w:
            return i
        random_val -= w
    return -1


def test_weighted_select_is_the_reference_loop_for_every_random_value(oracle):
    models = oracle.Models([dict(name=f"m{i}", critical=False, targets=t) for i, (t, _) in enumerate(GO_TABLES)])
    for i, (targets, want_name) in enumerate(GO_TABLES):
        total = sum(w for _, w in targets)
        hit = set()
        for v in range(total):
            k = models.weighted_select(i, v)
            assert k == go_loop(targets, v)
            hit.add(targets[k][0])
        assert models.weighted_select(i, total) == -1              # the `return ""` tail
        assert want_name in hit                                     # the Go test's answer is reachable
        # the seed-420 answers pin an interval of randomVal per table (consistent with one draw)
    assert go_loop(GO_TABLES[0][0], 49) == 0 and go_loop(GO_TABLES[1][0], 80) == 2 and go_loop(GO_TABLES[2][0], 20) == 1


def test_draw_distribution_matches_the_weights(oracle):
    models = oracle.Models([dict(name=f"m{i}", critical=False, targets=t) for i, (t, _) in enumerate(GO_TABLES)])
    n = 60_000
    for i, (targets, _) in enumerate(GO_TABLES):
        total = sum(w for _, w in targets)
        counts = np.zeros(len(targets))
        for key in range(n):
            counts[models.random_weighted_draw(i, (12345 ^ key) ^ oracle.LIGO_DRAW_DOMAIN)] += 1
        expect = np.array([w for _, w in targets]) / total
        assert np.abs(counts / n - expect).max() < 0.01, (i, counts / n, expect)
    # the draw is Int31n(sum) on the request's SplitMix stream, then the loop
    for key in (0, 1, 99, 2**40 + 7):
        st = (7 ^ key) ^ oracle.LIGO_DRAW_DOMAIN
        for i, (targets, _) in enumerate(GO_TABLES):
            r = oracle.int31n(st, sum(w for _, w in targets))
            assert models.random_weighted_draw(i, st) == go_loop(targets, r)


def test_resolve_semantics(oracle):
    models = oracle.Models([
        dict(name="plain", critical=False, targets=[]),                       # TargetModels empty: name passes through
        dict(name="crit-one", critical=True, targets=[("only", 7)]),
        None,                                                                 # no InferenceModel for this id
        dict(name="split", critical=True, targets=[("a", 1), ("b", 0), ("c", 3)]),
        dict(name="zero", critical=False, targets=[("a", 0)]),                # Go: Int31n(0) panics
    ])
    assert models.resolve(0, 1, 2) == (0, "plain", False, 255)
    assert models.resolve(1, 1, 2) == (0, "only", True, 0)
    assert models.resolve(2, 1, 2)[0] == oracle.LIGO_NO_MODEL                 # request.go:43-45
    assert models.resolve(9, 1, 2)[0] == oracle.LIGO_NO_MODEL
    seen = set()
    for key in range(400):
        rc, name, crit, k = models.resolve(3, 5, key)
        assert rc == 0 and crit and name == ["a", "b", "c"][k]
        seen.add(name)
    assert seen == {"a", "c"}                                                 # a zero-weight target is never drawn
    assert models.resolve(4, 1, 2)[0] == oracle.LIGO_NO_TARGET


@pytest.mark.parametrize("cfg", ["C1", "C2", "C5"])
def test_class_table_cpu_equals_the_port(cfg, oracle):
    c = WL.CONFIGS[cfg]
    snap = WL.make_snapshot(c["P"], c["A"])
    p = snap.packed
    reqs = WL.make_requests(min(c["R"], 20000), c["A"])
    reqs["adapter_id"][:50] = [-1, c["A"] + 3, 2**31 - 1, -(2**31), c["A"]] * 10      # out-of-range ids
    want, _ = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 77, False,
                                                             oracle.hardware_threads())
    tab = oracle.ClassTable(p.P, p.A, p.kv, p.q, p.n_active, p.max_active, p.bitmap, nthreads=4)
    got = tab.schedule_batch(reqs, 77, nthreads=3)
    assert np.array_equal(got, want)
    pool = oracle.Pool(snap.pod_records())
    names = snap.adapter_names() + [WL.UNKNOWN_MODEL]
    for crit in (False, True):
        for a in (0, c["A"] // 2, c["A"] - 1, c["A"]):
            rc, idx = pool.filter(names[a], crit)
            st, n, lst = tab.klass(crit, a)
            assert (st, n, lst) == (rc, len(idx) if rc == 0 else 0, idx if rc == 0 else []), (crit, a)


def test_class_table_cpu_adversarial_pools(oracle):
    import math
    rng = np.random.default_rng(5)
    kv_special = [0.0, -0.0, 0.8, 0.8000000000000002, 1.0, math.inf, -math.inf, math.nan, 5e-324]
    for it in range(60):
        P = int(rng.integers(0, 80))
        A = 4
        recs, act = [], []
        for i in range(P):
            a = sorted(rng.choice(A, size=int(rng.integers(0, 4)), replace=False).tolist())
            act.append(a)
            recs.append(dict(name=f"p{i}", address=f"a{i}",
                             waiting_queue_size=int(rng.choice([0, 5, 6, 49, 50, -1, 70])) if rng.random() < 0.3 else int(rng.integers(0, 60)),
                             kv_cache_usage_percent=float(rng.choice(kv_special)) if rng.random() < 0.25 else float(np.round(rng.random(), 2)),
                             max_active_models=int(rng.integers(0, 4)), active_models=[WL.adapter_name(x) for x in a]))
        from llm_instance_gateway_b200.packer import pack_columns
        W = (P + 31) // 32
        bm = np.zeros((A, W), dtype=np.uint32)
        for p_, a in enumerate(act):
            for x in a:
                bm[x, p_ >> 5] |= np.uint32(1 << (p_ & 31))
        pk = pack_columns([r["kv_cache_usage_percent"] for r in recs], [r["waiting_queue_size"] for r in recs],
                          [len(a) for a in act], [r["max_active_models"] for r in recs], bm)
        reqs = WL.make_requests(300, A, seed=it)
        want, _ = oracle.Pool(recs).schedule_batch([WL.adapter_name(a) for a in range(A)], WL.UNKNOWN_MODEL, reqs, it)
        got = oracle.ClassTable(pk.P, pk.A, pk.kv, pk.q, pk.n_active, pk.max_active, pk.bitmap).schedule_batch(reqs, it)
        assert np.array_equal(got, want), it


def test_models_batch_equals_resolve_then_schedule(oracle):
    A, P = 8, 64
    snap = WL.make_snapshot(P, A, seed=3)
    models = WL.make_models(A)
    mo = oracle.Models(WL.model_records(models))
    pool = oracle.Pool(snap.pod_records())
    ids = WL.make_model_requests(3000, A, seed=4)
    out = mo.schedule_batch(pool, ids, 11, first_index=1000)
    for i in range(0, 3000, 37):
        rc, name, crit, k = mo.resolve(int(ids[i]), 11, 1000 + i) if ids[i] < len(models) else (3, None, False, 255)
        if rc != 0:
            assert out[i]["status"] == rc and out[i]["pod_idx"] == -1
            continue
        st, pod, _ = pool.schedule(name, crit, 11, 1000 + i)
        assert (out[i]["status"], out[i]["pod_idx"], out[i]["target_idx"]) == (st, pod if st == 0 else -1, k)


def test_c_and_python_restatements_of_the_resolve_step_agree(oracle):
    """Two independently written restatements (C: lig_oracle_models.c, Python: lig_oracle_py.py) of
    request.go:42-56 + datastore.go:78-98 give the same draw for random model tables and keys."""
    from oracle import lig_oracle_py as PY
    rng = np.random.default_rng(17)
    for it in range(60):
        n_models = int(rng.integers(1, 12))
        recs, py_models = [], []
        for m in range(n_models):
            if rng.random() < 0.1:
                recs.append(None)
                py_models.append(None)
                continue
            nt = int(rng.integers(0, 6))
            tms = [(f"t{m}-{k}", int(rng.choice([0, 1, 2, 10, 50, 100, 2**20, 2**29]))) for k in range(nt)]
            recs.append(dict(name=f"m{m}", critical=bool(rng.integers(0, 2)), targets=tms))
            py_models.append(PY.InferenceModel(f"m{m}", bool(recs[-1]["critical"]), [PY.TargetModel(n, w) for n, w in tms]))
        mo = oracle.Models(recs)
        for _ in range(40):
            m = int(rng.integers(0, n_models + 2))
            seed, key = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**63))
            rc_c, name_c, crit_c, k_c = mo.resolve(m, seed, key) if m < n_models else (oracle.LIGO_NO_MODEL, None, False, 255)
            rc_p, name_p, crit_p, k_p = PY.resolve(py_models, m, seed, key)
            assert (rc_c, name_c) == (rc_p, name_p), (it, m, recs[m] if m < n_models else None)
            if rc_c == 0:
                assert (crit_c, k_c) == (crit_p, k_p)
