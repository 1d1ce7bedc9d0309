"""Seeded synthetic snapshots and request batches (the distributions of SURVEY.md section 8d).

Shared by bench.py and the parity tests so both see identical inputs.  Pure numpy, no GPU.
Adapter a is named ``adapter-<a>`` (the naming of the reference's load generator,
pkg/ext-proc/test/benchmark/benchmark.go:108-110); pods are ``pod-<i>`` / ``address-<i>``
(pkg/ext-proc/test/utils.go:73-80).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .packer import REQ_DTYPE, PackedModels, PackedSnapshot, pack_columns, pack_models
from .backend import CRITICAL, DEFAULT, SHEDDABLE, InferenceModel, InferenceModelSpec, Pod, TargetModel

SNAPSHOT_SEED = 0xC0FFEE
REQUEST_SEED = 0xBADC0DE
UNKNOWN_MODEL = "base-model"     # a ResolvedTargetModel that is in no pod's ActiveModels

# BASELINE.json configs (R, P, A)
CONFIGS = {
    "C1": dict(R=5000, P=8, A=4),
    "C2": dict(R=1024, P=64, A=32),
    "C3": dict(R=65536, P=512, A=256),
    "C4": dict(R=1 << 20, P=4096, A=1024),
    "C5": dict(R=100_000, P=256, A=64),
}


def adapter_name(a: int) -> str:
    return f"adapter-{a}"


def _zipf_weights(A: int, s: float = 1.1) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, A + 1, dtype=np.float64), s)
    return w / w.sum()


@dataclass
class SyntheticSnapshot:
    packed: PackedSnapshot
    active: List[List[int]]      # per pod: adapter ids in ActiveModels
    max_active64: np.ndarray     # the un-saturated Go-width value
    q64: np.ndarray

    def pod_records(self) -> List[dict]:
        p = self.packed
        return [dict(name=f"pod-{i}", address=f"address-{i}",
                     waiting_queue_size=int(self.q64[i]),
                     kv_cache_usage_percent=float(p.kv[i]),
                     max_active_models=int(self.max_active64[i]),
                     active_models=[adapter_name(a) for a in self.active[i]])
                for i in range(p.P)]

    def adapter_names(self) -> List[str]:
        return [adapter_name(a) for a in range(self.packed.A)]


def make_snapshot(P: int, A: int, seed: int = SNAPSHOT_SEED) -> SyntheticSnapshot:
    rng = np.random.default_rng(seed)
    # WaitingQueueSize: 70 % U{0..8}, 25 % U{9..60}, 5 % U{61..200}
    u = rng.random(P)
    q = np.where(u < 0.70, rng.integers(0, 9, P),
                 np.where(u < 0.95, rng.integers(9, 61, P), rng.integers(61, 201, P))).astype(np.int64)
    # KVCacheUsagePercent: U[0,1); half the pods rounded to a 1e-3 grid (exact ties)
    kv = rng.random(P)
    grid = rng.random(P) < 0.5
    kv = np.where(grid, np.round(kv * 1000.0) / 1000.0, kv)
    # MaxActiveModels in {0 (2 %), 4, 8, 16}
    ma = rng.choice(np.array([4, 8, 16]), size=P).astype(np.int64)
    ma[rng.random(P) < 0.02] = 0
    W = (P + 31) // 32
    bitmap = np.zeros((A, W), dtype=np.uint32)
    w = _zipf_weights(A) if A > 0 else None
    active: List[List[int]] = []
    na = np.zeros(P, dtype=np.int64)
    for p in range(P):
        size = int(rng.integers(0, ma[p] + 2))          # 0 .. MaxActive+1 (over-full allowed)
        size = min(size, A)
        ids = sorted(rng.choice(A, size=size, replace=False, p=w).tolist()) if size else []
        active.append(ids)
        na[p] = len(ids)
        for a in ids:
            bitmap[a, p >> 5] |= np.uint32(1 << (p & 31))
    ids_map = {adapter_name(a): a for a in range(A)}
    packed = pack_columns(kv, q, na, ma, bitmap, ids_map,
                          [Pod(f"pod-{i}", f"address-{i}") for i in range(P)])
    return SyntheticSnapshot(packed=packed, active=active, max_active64=ma, q64=q)


def make_requests(R: int, A: int, seed: int = REQUEST_SEED, out: np.ndarray = None) -> np.ndarray:
    """adapter ~ Zipf(1.1) over A with 3 % unknown (id = A); critical ~ Bernoulli(0.5)."""
    rng = np.random.default_rng(seed)
    reqs = out if out is not None else np.zeros(R, dtype=REQ_DTYPE)
    if A > 0:
        cdf = np.cumsum(_zipf_weights(A))
        ids = np.searchsorted(cdf, rng.random(R), side="right").astype(np.int32)
        ids = np.minimum(ids, A - 1)
    else:
        ids = np.zeros(R, dtype=np.int32)
    ids[rng.random(R) < 0.03] = A
    reqs["adapter_id"] = ids
    reqs["flags"] = (rng.random(R) < 0.5).astype(np.uint32)
    reqs["rand_key"] = rng.integers(0, 1 << 64, size=R, dtype=np.uint64)
    return reqs


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous request shard [lo, hi) of rank `rank` (what lig_group_schedule_batch and bench.py
    use): result order == request order, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    return total * rank // world, total * (rank + 1) // world


def algorithmic_bytes(R: int, P: int, A: int) -> int:
    """24 R + 16 P + 4 A ceil(P/32)  (SURVEY.md section 8d)."""
    return 24 * R + 16 * P + 4 * A * ((P + 31) // 32)


# ---- synthetic InferenceModels (the datastore the request pre-step reads) ------------------------
# The three weight tables of the reference's own test (backend/datastore_test.go:9-76).
WEIGHT_TABLES = ([("canary", 50), ("v1", 50)], [("canary", 25), ("v1.1", 55), ("v1", 50)],
                 [("canary", 20), ("v1.1", 20), ("v1", 10)])
N_SPLIT_MODELS = 64


def _mix(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
    return x ^ (x >> 31)


def make_models(A: int) -> List[InferenceModel]:
    """A + 1 + N_SPLIT_MODELS InferenceModels over the A adapters of a synthetic snapshot:
      model-<a>   one target model adapter-<a> (weight 100); Critical for ~half of them
      base-model  no TargetModels: the requested model name passes through (request.go:47)
      split-<j>   2-3 weighted target models with the weight tables of datastore_test.go:9-76,
                  the targets being adapters (or, for one in eight, a never-loaded adapter name)"""
    models = []
    for a in range(A):
        crit = CRITICAL if _mix(a) & 1 else (DEFAULT if _mix(a) & 2 else SHEDDABLE)
        models.append(InferenceModel(f"model-{a}", InferenceModelSpec(
            ModelName=f"model-{a}", Criticality=crit, TargetModels=[TargetModel(adapter_name(a), 100)])))
    models.append(InferenceModel(UNKNOWN_MODEL, InferenceModelSpec(ModelName=UNKNOWN_MODEL, Criticality=None)))
    for j in range(N_SPLIT_MODELS):
        table = WEIGHT_TABLES[j % 3]
        tms = []
        for k, (_, w) in enumerate(table):
            t = _mix(1000 * j + k) % max(A, 1)
            name = adapter_name(t) if (A and (_mix(7 * j + k) % 8)) else f"unloaded-{j}-{k}"
            tms.append(TargetModel(name, w))
        models.append(InferenceModel(f"split-{j}", InferenceModelSpec(
            ModelName=f"split-{j}", Criticality=CRITICAL if j & 1 else DEFAULT, TargetModels=tms)))
    return models


def make_model_requests(R: int, A: int, seed: int = REQUEST_SEED) -> np.ndarray:
    """uint32 model ids: 89 % model-<a> with a ~ Zipf(1.1), 3 % base-model, 8 % split-<j> (uniform),
    0.05 % ids no InferenceModel exists for (FetchModelData returns nil)."""
    rng = np.random.default_rng(seed)
    n_models = A + 1 + N_SPLIT_MODELS
    if A > 0:
        cdf = np.cumsum(_zipf_weights(A))
        ids = np.minimum(np.searchsorted(cdf, rng.random(R), side="right"), A - 1).astype(np.uint32)
    else:
        ids = np.full(R, A, dtype=np.uint32)
    u = rng.random(R)
    ids[u < 0.03] = A
    split = (u >= 0.03) & (u < 0.11)
    ids[split] = (A + 1 + rng.integers(0, N_SPLIT_MODELS, int(split.sum()))).astype(np.uint32)
    ids[u > 0.9995] = n_models + 5
    return np.ascontiguousarray(ids, dtype=np.uint32)


def model_records(models: List[InferenceModel]) -> List[dict]:
    return [dict(name=m.Spec.ModelName, critical=(m.Spec.Criticality == CRITICAL),
                 targets=[(t.Name, t.Weight) for t in m.Spec.TargetModels]) for m in models]
