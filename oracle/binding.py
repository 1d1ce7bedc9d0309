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
MPICK_DTYPE = np.dtype([("pod_idx", "<i2"), ("status", "u1"), ("target_idx", "u1")])
LIGO_NO_MODEL, LIGO_NO_TARGET = 3, 4
LIGO_DRAW_DOMAIN = 0xA0761D6478BD642F

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
    lib.lig_oracle_classtab_build.argtypes = [i32, i32, vp, vp, vp, vp, vp, dbl, i64, i64, i32]
    lib.lig_oracle_classtab_build.restype = vp
    lib.lig_oracle_classtab_free.argtypes = [vp]
    lib.lig_oracle_classtab_free.restype = None
    lib.lig_oracle_classtab_schedule_batch.argtypes = [vp, vp, i32, u64, vp, i32]
    lib.lig_oracle_classtab_class.argtypes = [vp, i32, i32, C.POINTER(i32), C.POINTER(i32), vp]
    lib.lig_oracle_models_new.argtypes = [i32]
    lib.lig_oracle_models_new.restype = vp
    lib.lig_oracle_models_free.argtypes = [vp]
    lib.lig_oracle_models_free.restype = None
    lib.lig_oracle_models_set.argtypes = [vp, i32, C.c_char_p, i32, cpp, vp, i32]
    lib.lig_oracle_weighted_select.argtypes = [vp, i32, C.c_int32]
    lib.lig_oracle_random_weighted_draw.argtypes = [vp, i32, u64]
    lib.lig_oracle_resolve.argtypes = [vp, i32, u64, u64, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i32)]
    lib.lig_oracle_schedule_models_batch.argtypes = [vp, vp, vp, i32, u64, u64, vp, i32]
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


def cpu_quota_cores():
    """cgroup v2 CPU quota of this container in cores (None = unlimited)."""
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        return None if quota == "max" else float(quota) / float(period)
    except Exception:
        return None


def usable_threads() -> int:
    """min(hardware threads, cgroup quota): more threads than the quota only adds contention."""
    n = hardware_threads()
    q = cpu_quota_cores()
    return max(1, min(n, int(q + 0.5))) if q else n


def splitmix64_stream(state: int, n: int) -> List[int]:
    lib = load()
    st = C.c_uint64(state)
    return [int(lib.lig_oracle_splitmix64_next(C.byref(st))) for _ in range(n)]


def int31n(state: int, n: int) -> int:
    st = C.c_uint64(state)
    return int(load().lig_oracle_int31n(C.byref(st), n))


class ClassTable:
    """The class-table CPU arm (lig_oracle_classtab.c): tree walked once per (critical, adapter)
    class per snapshot, then table lookup + Int31n per request."""

    def __init__(self, P, A, kv, q, n_active, max_active, bitmap, thresholds=(0.8, 5, 50), nthreads=1):
        self._lib = load()
        self._keep = [np.ascontiguousarray(kv, dtype=np.float64), np.ascontiguousarray(q, dtype=np.int32),
                      np.ascontiguousarray(n_active, dtype=np.uint16), np.ascontiguousarray(max_active, dtype=np.uint16),
                      np.ascontiguousarray(bitmap, dtype=np.uint32)]
        ptr = lambda a: a.ctypes.data if a.size else None
        self.P, self.A = P, A
        self._t = self._lib.lig_oracle_classtab_build(P, A, *[ptr(a) for a in self._keep], thresholds[0],
                                                      thresholds[1], thresholds[2], nthreads)
        assert self._t

    def __del__(self):
        try:
            if self._t:
                self._lib.lig_oracle_classtab_free(self._t)
                self._t = None
        except Exception:
            pass

    def schedule_batch(self, reqs: np.ndarray, seed: int, nthreads: int = 1, out: Optional[np.ndarray] = None):
        assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
        R = int(reqs.shape[0])
        if out is None:
            out = np.zeros(R, dtype=PICK_DTYPE)
        rc = self._lib.lig_oracle_classtab_schedule_batch(self._t, reqs.ctypes.data if R else None, R, seed,
                                                          out.ctypes.data if R else None, nthreads)
        assert rc == 0
        return out

    def klass(self, critical: bool, adapter_id: int):
        st, n = C.c_int(), C.c_int()
        lst = np.zeros(max(self.P, 1), dtype=np.uint16)
        assert self._lib.lig_oracle_classtab_class(self._t, int(critical), adapter_id, C.byref(st), C.byref(n),
                                                   lst.ctypes.data) == 0
        return st.value, n.value, lst[: n.value].tolist()


class Models:
    """The datastore's InferenceModels by dense id (lig_oracle_models.c)."""

    def __init__(self, models: Sequence[dict]):
        """models: dicts with name, critical, targets=[(name, weight), ...]; None = absent id."""
        self._lib = load()
        self.n = len(models)
        self._m = self._lib.lig_oracle_models_new(self.n)
        for i, m in enumerate(models):
            if m is None:
                continue
            tn = [t[0] for t in m.get("targets", [])]
            tw = np.array([t[1] for t in m.get("targets", [])], dtype=np.int32)
            rc = self._lib.lig_oracle_models_set(self._m, i, m["name"].encode(), int(bool(m.get("critical"))),
                                                 _names(tn), tw.ctypes.data if tw.size else None, len(tn))
            assert rc == 0

    def __del__(self):
        try:
            if self._m:
                self._lib.lig_oracle_models_free(self._m)
                self._m = None
        except Exception:
            pass

    def weighted_select(self, model: int, random_val: int) -> int:
        return int(self._lib.lig_oracle_weighted_select(self._m, model, random_val))

    def random_weighted_draw(self, model: int, state: int) -> int:
        return int(self._lib.lig_oracle_random_weighted_draw(self._m, model, state))

    def resolve(self, model: int, seed: int, rand_key: int):
        name, crit, tgt = C.c_char_p(), C.c_int(), C.c_int()
        rc = self._lib.lig_oracle_resolve(self._m, model, seed, rand_key, C.byref(name), C.byref(crit), C.byref(tgt))
        return rc, (name.value.decode() if rc == 0 else None), bool(crit.value), tgt.value

    def schedule_batch(self, pool: "Pool", model_ids: np.ndarray, seed: int, first_index: int = 0,
                       nthreads: int = 1) -> np.ndarray:
        assert model_ids.dtype == np.uint32 and model_ids.flags.c_contiguous
        R = int(model_ids.shape[0])
        out = np.zeros(R, dtype=MPICK_DTYPE)
        rc = self._lib.lig_oracle_schedule_models_batch(pool._p, self._m, model_ids.ctypes.data if R else None, R,
                                                        seed, first_index, out.ctypes.data if R else None, nthreads)
        assert rc == 0
        return out
