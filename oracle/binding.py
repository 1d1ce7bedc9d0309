"""ctypes binding of oracle/liblig_oracle.so.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this; the product package (llm_instance_gateway_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblig_oracle.so")

LIGO_OK, LIGO_DROP, LIGO_EMPTY, LIGO_ERROR = 0, 1, 2, 3
REQ_DTYPE = np.dtype([("adapter_id", "<i4"), ("flags", "<u4"), ("rand_key", "<u8")])
PICK_DTYPE = np.dtype([("pod_idx", "<i4"), ("status", "<u2"), ("n_survivors", "<u2")])

FILTER_FUNCS = {
    "leastQueuingFilterFunc": 0, "leastKVCacheFilterFunc": 1, "lowLoRACostPredicate": 2,
    "loRAAffinityPredicate": 3, "canAcceptNewLoraPredicate": 4, "lowQueueingPodPredicate": 5,
    "criticalRequestPredicate": 6, "noQueueAndLessThanKVCacheThresholdPredicate": 7,
}

_lib = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    cpp = C.POINTER(C.c_char_p)
    lib.lig_oracle_pool_new.argtypes = [i32]
    lib.lig_oracle_pool_new.restype = vp
    lib.lig_oracle_pool_free.argtypes = [vp]
    lib.lig_oracle_pool_free.restype = None
    lib.lig_oracle_pool_set_pod.argtypes = [vp, i32, C.c_char_p, C.c_char_p, i64, dbl, i64, cpp, i32]
    lib.lig_oracle_pool_size.argtypes = [vp]
    lib.lig_oracle_set_thresholds.argtypes = [dbl, i64, i64]
    lib.lig_oracle_set_thresholds.restype = None
    lib.lig_oracle_filter.argtypes = [vp, C.c_char_p, i32, vp, C.POINTER(i32)]
    lib.lig_oracle_filter_func.argtypes = [vp, i32, C.c_char_p, i32, i64, dbl, vp, C.POINTER(i32)]
    lib.lig_oracle_filter_error_leaf.argtypes = [vp, vp, C.POINTER(i32)]
    lib.lig_oracle_schedule.argtypes = [vp, C.c_char_p, i32, u64, u64, C.POINTER(C.c_int32), C.POINTER(i32)]
    lib.lig_oracle_schedule_batch.argtypes = [vp, cpp, i32, C.c_char_p, vp, i32, u64, vp, vp, i32]
    lib.lig_oracle_soa_schedule_batch.argtypes = [i32, i32, vp, vp, vp, vp, vp, dbl, i64, i64, vp, i32, u64,
                                                  vp, vp, i32]
    lib.lig_oracle_splitmix64_next.argtypes = [C.POINTER(u64)]
    lib.lig_oracle_splitmix64_next.restype = u64
    lib.lig_oracle_int31n.argtypes = [C.POINTER(u64), C.c_int32]
    lib.lig_oracle_int31n.restype = C.c_int32
    _lib = lib
    return lib


def _names(names: Sequence[str]):
    arr = (C.c_char_p * max(len(names), 1))()
    for i, n in enumerate(names):
        arr[i] = n.encode()
    return arr


class Pool:
    """A fake PodMetricsProvider holding an ordered slice of pods (the oracle's input)."""

    def __init__(self, pods: Sequence[dict]):
        """pods: dicts with name, address, waiting_queue_size, kv_cache_usage_percent,
        max_active_models, active_models (list of names)."""
        self._lib = load()
        self.n = len(pods)
        self._p = self._lib.lig_oracle_pool_new(self.n)
        for i, p in enumerate(pods):
            act = list(p.get("active_models", []))
            rc = self._lib.lig_oracle_pool_set_pod(
                self._p, i, p.get("name", "").encode(), p.get("address", "").encode(),
                int(p.get("waiting_queue_size", 0)), float(p.get("kv_cache_usage_percent", 0.0)),
                int(p.get("max_active_models", 0)), _names(act), len(act))
            assert rc == 0

    def __del__(self):
        try:
            if self._p:
                self._lib.lig_oracle_pool_free(self._p)
                self._p = None
        except Exception:
            pass

    def filter(self, model: str, critical: bool) -> Tuple[int, List[int]]:
        out = np.zeros(max(self.n, 1), dtype=np.int32)
        n = C.c_int()
        rc = self._lib.lig_oracle_filter(self._p, model.encode(), int(critical), out.ctypes.data, C.byref(n))
        return rc, out[: n.value].tolist()

    def filter_func(self, which: str, model: Optional[str] = None, critical: bool = False,
                    q_thr: int = 0, kv_thr: float = 0.0) -> Tuple[int, List[int]]:
        out = np.zeros(max(self.n, 1), dtype=np.int32)
        n = C.c_int()
        rc = self._lib.lig_oracle_filter_func(self._p, FILTER_FUNCS[which], (model or "").encode(),
                                              int(critical), q_thr, kv_thr, out.ctypes.data, C.byref(n))
        return rc, out[: n.value].tolist()

    def filter_error_leaf(self) -> Tuple[int, List[int]]:
        out = np.zeros(max(self.n, 1), dtype=np.int32)
        n = C.c_int()
        rc = self._lib.lig_oracle_filter_error_leaf(self._p, out.ctypes.data, C.byref(n))
        return rc, out[: n.value].tolist()

    def schedule(self, model: str, critical: bool, seed: int, rand_key: int) -> Tuple[int, int, int]:
        pod, n = C.c_int32(), C.c_int()
        rc = self._lib.lig_oracle_schedule(self._p, model.encode(), int(critical), seed, rand_key,
                                           C.byref(pod), C.byref(n))
        return rc, pod.value, n.value

    def schedule_batch(self, adapter_names: Sequence[str], unknown_name: str, reqs: np.ndarray,
                       seed: int, want_masks: bool = False, nthreads: int = 1):
        assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
        R = int(reqs.shape[0])
        out = np.zeros(R, dtype=PICK_DTYPE)
        W = (self.n + 31) // 32
        masks = np.zeros((R, W), dtype=np.uint32) if want_masks else None
        rc = self._lib.lig_oracle_schedule_batch(
            self._p, _names(adapter_names), len(adapter_names), unknown_name.encode(),
            reqs.ctypes.data if R else None, R, seed, out.ctypes.data if R else None,
            masks.ctypes.data if (masks is not None and masks.size) else None, nthreads)
        assert rc == 0
        return out, masks


def soa_schedule_batch(P, A, kv, q, n_active, max_active, bitmap, reqs, seed, want_masks=False,
                       nthreads=1, thresholds=(0.8, 5, 50)):
    """The optimised-CPU variant on the packed columns (numpy arrays as in PackedSnapshot)."""
    lib = load()
    assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
    R = int(reqs.shape[0])
    out = np.zeros(R, dtype=PICK_DTYPE)
    W = (P + 31) // 32
    masks = np.zeros((R, W), dtype=np.uint32) if want_masks else None
    ptr = lambda a: a.ctypes.data if a is not None and a.size else None
    rc = lib.lig_oracle_soa_schedule_batch(
        P, A, ptr(np.ascontiguousarray(kv, dtype=np.float64)), ptr(np.ascontiguousarray(q, dtype=np.int32)),
        ptr(np.ascontiguousarray(n_active, dtype=np.uint16)), ptr(np.ascontiguousarray(max_active, dtype=np.uint16)),
        ptr(np.ascontiguousarray(bitmap, dtype=np.uint32)), thresholds[0], thresholds[1], thresholds[2],
        ptr(reqs), R, seed, ptr(out), ptr(masks), nthreads)
    assert rc == 0
    return out, masks


def set_thresholds(kv=0.8, qcrit=5, qlora=50) -> None:
    load().lig_oracle_set_thresholds(kv, qcrit, qlora)


def hardware_threads() -> int:
    return int(load().lig_oracle_hardware_threads())


def splitmix64_stream(state: int, n: int) -> List[int]:
    lib = load()
    st = C.c_uint64(state)
    return [int(lib.lig_oracle_splitmix64_next(C.byref(st))) for _ in range(n)]


def int31n(state: int, n: int) -> int:
    st = C.c_uint64(state)
    return int(load().lig_oracle_int31n(C.byref(st), n))
