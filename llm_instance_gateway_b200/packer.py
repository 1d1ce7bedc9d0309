"""Snapshot packer: ``[]*backend.PodMetrics`` -> the device layout of include/lig.h.

This is the per-refresh-tick step that replaces the reference's per-request
``AllPodMetrics()`` materialisation (pkg/ext-proc/scheduling/scheduler.go:114-115,
pkg/ext-proc/backend/provider.go:38-46): adapter names are interned to dense ids, pod metrics
become four columns, and ``ActiveModels`` membership becomes an adapter-major bitmap.  The same
for the datastore's InferenceModels (pkg/ext-proc/backend/datastore.go:70-98): model names become
dense model ids, ``Spec.TargetModels`` a CSR table of (adapter id, weight).

Pure numpy on purpose: nothing here loads the CUDA library, so tools that only need the synthetic
workload (the reference arm of bench.py) never touch liblig.so.  The C helpers lig_pack_pods /
lig_pack_snapshot do the same narrowing for non-Python hosts; tests/test_abi_cpu.py checks that
both produce the same bytes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from .backend import InferenceModel, IsCritical, Pod, PodMetrics

REQ_DTYPE = np.dtype([("adapter_id", "<i4"), ("flags", "<u4"), ("rand_key", "<u8")])
PICK_DTYPE = np.dtype([("pod_idx", "<i4"), ("status", "<u2"), ("n_survivors", "<u2")])
MPICK_DTYPE = np.dtype([("pod_idx", "<i2"), ("status", "u1"), ("target_idx", "u1")])
assert REQ_DTYPE.itemsize == 16 and PICK_DTYPE.itemsize == 8 and MPICK_DTYPE.itemsize == 4

LIG_MAX_ADAPTERS = 65534


class RangeError(ValueError):
    """A host value does not fit the device record (the C helpers return LIG_ERR_RANGE)."""


@dataclass
class PackedSnapshot:
    P: int
    A: int
    kv: np.ndarray          # float64[P]   Metrics.KVCacheUsagePercent
    q: np.ndarray           # int32[P]     Metrics.WaitingQueueSize
    n_active: np.ndarray    # uint16[P]    len(Metrics.ActiveModels)
    max_active: np.ndarray  # uint16[P]    Metrics.MaxActiveModels, saturated
    bitmap: np.ndarray      # uint32[A, ceil(P/32)]  adapter-major membership
    adapter_ids: Dict[str, int]
    pods: List[Pod]         # pod_idx -> backend.Pod (returned by value, backend/types.go:8-11)

    @property
    def W(self) -> int:
        return (self.P + 31) // 32

    def adapter_id(self, model_name: str) -> int:
        """Dense id of ResolvedTargetModel; A (= "in no pod's ActiveModels") when unknown."""
        return self.adapter_ids.get(model_name, self.A)

    def blob(self) -> np.ndarray:
        """The packed blob (uint8) exactly as lig_upload_snapshot_device expects it:
        kv[Ppad] | q[Ppad] | n_active[Ppad] | max_active[Ppad] | bitmap[A][W], 16-byte padded."""
        P, A, W = self.P, self.A, self.W
        Ppad = 32 * W
        total = max(16, (16 * Ppad + 4 * A * W + 15) // 16 * 16)
        out = np.zeros(total, dtype=np.uint8)
        out[0: 8 * P] = self.kv.view(np.uint8)
        o = 8 * Ppad
        out[o: o + 4 * P] = self.q.view(np.uint8)
        o += 4 * Ppad
        out[o: o + 2 * P] = self.n_active.view(np.uint8)
        o += 2 * Ppad
        out[o: o + 2 * P] = self.max_active.view(np.uint8)
        o += 2 * Ppad
        if A and W:
            bm = np.ascontiguousarray(self.bitmap, dtype=np.uint32).copy()
            if P & 31:
                bm[:, W - 1] &= np.uint32((1 << (P & 31)) - 1)      # bits of padding pods stay clear
            out[o: o + 4 * A * W] = bm.view(np.uint8).reshape(-1)
        return out

    def algorithmic_snapshot_bytes(self) -> int:
        """S(P, A) = 16 P + 4 A ceil(P/32)  (SURVEY.md section 8d)."""
        return 16 * self.P + 4 * self.A * self.W


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data if a is not None and a.size else None


def pack_columns(kv, q64, n_active64, max_active64, bitmap, adapter_ids=None, pods=None) -> PackedSnapshot:
    """Narrow Go-width columns to the device record (range-checked, never silently wrapped):
    q must fit int32, n_active in [0, 65534], max_active saturates to [0, 65535] — saturation cannot
    change `n_active < max_active` (canAcceptNewLoraPredicate, filter.go:175-177)."""
    P = int(len(kv))
    kv = np.ascontiguousarray(kv, dtype=np.float64)
    q64 = np.ascontiguousarray(q64, dtype=np.int64)
    na64 = np.ascontiguousarray(n_active64, dtype=np.int64)
    ma64 = np.ascontiguousarray(max_active64, dtype=np.int64)
    if P and (q64.min() < -(1 << 31) or q64.max() > (1 << 31) - 1):
        bad = int(np.nonzero((q64 < -(1 << 31)) | (q64 > (1 << 31) - 1))[0][0])
        raise RangeError(f"pod {bad}: WaitingQueueSize {int(q64[bad])} does not fit int32")
    if P and (na64.min() < 0 or na64.max() > LIG_MAX_ADAPTERS):
        bad = int(np.nonzero((na64 < 0) | (na64 > LIG_MAX_ADAPTERS))[0][0])
        raise RangeError(f"pod {bad}: len(ActiveModels) {int(na64[bad])} outside [0, {LIG_MAX_ADAPTERS}]")
    q = q64.astype(np.int32)
    na = na64.astype(np.uint16)
    ma = np.clip(ma64, 0, 65535).astype(np.uint16)
    W = (P + 31) // 32
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint32)
    A = int(bitmap.shape[0]) if bitmap.ndim == 2 else (bitmap.size // W if W else 0)
    bitmap = bitmap.reshape(A, W)
    return PackedSnapshot(P=P, A=A, kv=kv, q=q, n_active=na, max_active=ma, bitmap=bitmap,
                          adapter_ids=dict(adapter_ids or {}),
                          pods=list(pods) if pods is not None else [Pod(f"pod-{i}", f"address-{i}") for i in range(P)])


def pack_pod_metrics(pod_metrics: Sequence[PodMetrics]) -> PackedSnapshot:
    """Pack the slice a PodMetricsProvider returned, keeping its order as the pod index."""
    P = len(pod_metrics)
    adapter_ids: Dict[str, int] = {}
    for pm in pod_metrics:
        for name in pm.Metrics.ActiveModels:
            if name not in adapter_ids:
                adapter_ids[name] = len(adapter_ids)
    A = len(adapter_ids)
    W = (P + 31) // 32
    bitmap = np.zeros((A, W), dtype=np.uint32)
    for p, pm in enumerate(pod_metrics):
        for name in pm.Metrics.ActiveModels:
            bitmap[adapter_ids[name], p >> 5] |= np.uint32(1 << (p & 31))
    kv = np.array([pm.Metrics.KVCacheUsagePercent for pm in pod_metrics], dtype=np.float64)
    q = np.array([pm.Metrics.WaitingQueueSize for pm in pod_metrics], dtype=np.int64)
    na = np.array([len(pm.Metrics.ActiveModels) for pm in pod_metrics], dtype=np.int64)
    ma = np.array([pm.Metrics.MaxActiveModels for pm in pod_metrics], dtype=np.int64)
    return pack_columns(kv, q, na, ma, bitmap, adapter_ids, [pm.Pod for pm in pod_metrics])


# ---- the datastore's InferenceModels -> the model table of lig_upload_models -----------------------
@dataclass
class PackedModels:
    n_models: int
    target_offsets: np.ndarray      # int32[n_models + 1]
    target_adapter_ids: np.ndarray  # int32[n_targets]   Spec.TargetModels[k].Name interned against a snapshot
    target_weights: np.ndarray      # int32[n_targets]   Spec.TargetModels[k].Weight
    critical: np.ndarray            # uint8[n_models]    IsCritical(model)
    self_adapter_ids: np.ndarray    # int32[n_models]    the model's own name interned (TargetModels empty)
    present: np.ndarray             # uint8[n_models]    0 = FetchModelData returns nil
    model_ids: Dict[str, int]       # Spec.ModelName -> dense id
    target_names: List[List[str]]   # per model: TargetModels names (for the body's "model" rewrite)


def pack_models(models: Sequence[InferenceModel], snap: PackedSnapshot) -> PackedModels:
    """Intern the datastore's InferenceModels against `snap` (request.go:42-56, datastore.go:70-105).

    A model whose weights sum to zero cannot be drawn from (Go's Int31n panics on 0; the CRD comment
    promises "no valid target model"): it is packed as absent, so requests for it get LIG_NO_MODEL."""
    n = len(models)
    off = np.zeros(n + 1, dtype=np.int32)
    ids: List[int] = []
    wts: List[int] = []
    crit = np.zeros(n, dtype=np.uint8)
    self_ids = np.zeros(n, dtype=np.int32)
    present = np.ones(n, dtype=np.uint8)
    model_ids: Dict[str, int] = {}
    names: List[List[str]] = []
    for m, im in enumerate(models):
        model_ids[im.Spec.ModelName] = m
        crit[m] = 1 if IsCritical(im) else 0
        self_ids[m] = snap.adapter_id(im.Spec.ModelName)
        tms = im.Spec.TargetModels
        if tms and sum(t.Weight for t in tms) <= 0:
            present[m] = 0
            tms = []
        names.append([t.Name for t in tms])
        for t in tms:
            ids.append(snap.adapter_id(t.Name))
            wts.append(int(t.Weight))
        off[m + 1] = len(ids)
    return PackedModels(n_models=n, target_offsets=off,
                        target_adapter_ids=np.array(ids, dtype=np.int32), target_weights=np.array(wts, dtype=np.int32),
                        critical=crit, self_adapter_ids=self_ids, present=present, model_ids=model_ids,
                        target_names=names)
