"""Snapshot packer: ``[]*backend.PodMetrics`` -> the device layout of include/lig.h.

This is the per-refresh-tick step that replaces the reference's per-request
``AllPodMetrics()`` materialisation (pkg/ext-proc/scheduling/scheduler.go:114-115,
pkg/ext-proc/backend/provider.go:38-46): adapter names are interned to dense ids, pod metrics
become four columns, and ``ActiveModels`` membership becomes an adapter-major bitmap.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np

from . import _native as N
from .backend import Pod, PodMetrics

REQ_DTYPE = np.dtype([("adapter_id", "<i4"), ("flags", "<u4"), ("rand_key", "<u8")])
PICK_DTYPE = np.dtype([("pod_idx", "<i4"), ("status", "<u2"), ("n_survivors", "<u2")])
assert REQ_DTYPE.itemsize == 16 and PICK_DTYPE.itemsize == 8


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
        """The packed blob (uint8) exactly as lig_upload_snapshot_device expects it."""
        lib = N.load()
        out = np.zeros(lib.lig_snapshot_bytes(self.P, self.A), dtype=np.uint8)
        N.check(lib.lig_pack_snapshot(out.ctypes.data, self.P, self.A, _ptr(self.kv), _ptr(self.q),
                                      _ptr(self.n_active), _ptr(self.max_active),
                                      _ptr(self.bitmap)))
        return out

    def algorithmic_snapshot_bytes(self) -> int:
        """S(P, A) = 16 P + 4 A ceil(P/32)  (SURVEY.md section 8d)."""
        return 16 * self.P + 4 * self.A * self.W


def _ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


def pack_columns(kv, q64, n_active64, max_active64, bitmap, adapter_ids=None, pods=None) -> PackedSnapshot:
    """Narrow Go-width columns through lig_pack_pods (range-checked, never silently wrapped)."""
    lib = N.load()
    P = int(len(kv))
    kv = np.ascontiguousarray(kv, dtype=np.float64)
    q64 = np.ascontiguousarray(q64, dtype=np.int64)
    na64 = np.ascontiguousarray(n_active64, dtype=np.int64)
    ma64 = np.ascontiguousarray(max_active64, dtype=np.int64)
    q = np.zeros(P, dtype=np.int32)
    na = np.zeros(P, dtype=np.uint16)
    ma = np.zeros(P, dtype=np.uint16)
    N.check(lib.lig_pack_pods(P, _ptr(q64), _ptr(na64), _ptr(ma64), _ptr(q), _ptr(na), _ptr(ma)))
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
