"""ctypes view of the native C++ host runtime (csrc/host/lig_host.hpp -> liblig_host.so).

The C++ ``lig::scheduling::Scheduler`` is the compiled-language counterpart of the Go adapter
(INTEGRATION.md): concurrent blocking ``Schedule`` calls folded into one C-ABI call per flush,
snapshot re-packed per refresh tick.  This module only drives it from Python for the tests and the
streaming benchmark.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .backend import Pod, PodMetrics

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblig_host.so")

GRPC_OK, GRPC_UNKNOWN, GRPC_RESOURCE_EXHAUSTED, GRPC_INTERNAL = 0, 2, 8, 13

EXPORTED_SYMBOLS = (
    "ligh_provider_new", "ligh_provider_free", "ligh_provider_set_pods", "ligh_scheduler_new",
    "ligh_scheduler_new2",
    "ligh_scheduler_free", "ligh_schedule", "ligh_refresh", "ligh_stats",
    "ligh_schedule_concurrent", "ligh_stream_bench", "ligh_refresh_timing", "ligh_flush_timing",
    "ligh_scheduler_new_devices", "ligh_datastore_new", "ligh_datastore_free", "ligh_datastore_set_model",
    "ligh_schedule_model",
)

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run `make -C llm_instance_gateway_b200/csrc`")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    cpp = C.POINTER(C.c_char_p)
    lib.ligh_provider_new.restype = vp
    lib.ligh_provider_free.argtypes = [vp]
    lib.ligh_provider_free.restype = None
    lib.ligh_provider_set_pods.argtypes = [vp, i32, cpp, cpp, vp, vp, vp, cpp, vp]
    lib.ligh_scheduler_new.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, u64, C.c_char_p, i32]
    lib.ligh_scheduler_new.restype = vp
    lib.ligh_scheduler_new2.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, u64, i32, i32, i32, C.c_char_p, i32]
    lib.ligh_scheduler_new2.restype = vp
    lib.ligh_scheduler_new_devices.argtypes = [vp, C.POINTER(i32), i32, i32, i32, i32, i32, i32, i32, u64, C.c_char_p, i32]
    lib.ligh_scheduler_new_devices.restype = vp
    lib.ligh_datastore_new.restype = vp
    lib.ligh_datastore_free.argtypes = [vp]
    lib.ligh_datastore_free.restype = None
    lib.ligh_datastore_set_model.argtypes = [vp, C.c_char_p, i32, i32, cpp, vp]
    lib.ligh_schedule_model.argtypes = [vp, vp, C.c_char_p, C.c_char_p, i32, C.c_char_p, i32, C.c_char_p, i32, C.c_char_p, i32]
    lib.ligh_scheduler_free.argtypes = [vp]
    lib.ligh_scheduler_free.restype = None
    lib.ligh_schedule.argtypes = [vp, C.c_char_p, C.c_char_p, i32, C.c_char_p, i32, C.c_char_p, i32,
                                  C.c_char_p, i32]
    lib.ligh_refresh.argtypes = [vp, C.c_char_p, i32]
    lib.ligh_stats.argtypes = [vp, vp]
    lib.ligh_stats.restype = None
    lib.ligh_refresh_timing.argtypes = [vp, vp]
    lib.ligh_refresh_timing.restype = None
    lib.ligh_flush_timing.argtypes = [vp, vp]
    lib.ligh_flush_timing.restype = None
    lib.ligh_schedule_concurrent.argtypes = [vp, i32, i32, cpp, vp, i32, vp, vp]
    lib.ligh_stream_bench.argtypes = [vp, C.c_double, C.c_double, i32, cpp, vp, i32, u64, vp, vp, i32,
                                      C.POINTER(i32), C.POINTER(i32)]
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)
    _lib = lib
    return lib


def _strs(items: Sequence[str]):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode()
    return arr


class HostSchedulerError(RuntimeError):
    pass


class HostProvider:
    """A mutable fake PodMetricsProvider living on the C++ side."""

    def __init__(self, pods: Sequence[PodMetrics] = ()):
        self._lib = load()
        self._p = self._lib.ligh_provider_new()
        self.set_pods(pods)

    def set_pods(self, pods: Sequence[PodMetrics]) -> None:
        self.pods = list(pods)
        n = len(self.pods)
        q = np.array([p.Metrics.WaitingQueueSize for p in self.pods], dtype=np.int64)
        kv = np.array([p.Metrics.KVCacheUsagePercent for p in self.pods], dtype=np.float64)
        ma = np.array([p.Metrics.MaxActiveModels for p in self.pods], dtype=np.int64)
        flat: List[str] = []
        off = [0]
        for p in self.pods:
            flat.extend(p.Metrics.ActiveModels.keys())
            off.append(len(flat))
        off = np.array(off, dtype=np.int32)
        rc = self._lib.ligh_provider_set_pods(
            self._p, n, _strs([p.Pod.Name for p in self.pods]), _strs([p.Pod.Address for p in self.pods]),
            q.ctypes.data, kv.ctypes.data, ma.ctypes.data, _strs(flat), off.ctypes.data)
        assert rc == 0

    def close(self):
        if self._p:
            self._lib.ligh_provider_free(self._p)
            self._p = None


class HostScheduler:
    def __init__(self, provider: HostProvider, device: int = 0, max_pods: int = 4096,
                 max_adapters: int = 1024, max_batch: int = 1 << 16, flush_size: int = 4096,
                 batch_window_us: int = 50, refresh_interval_ms: int = 0, seed: int = 1,
                 busy_poll: bool = False, caller_spin_us: int = 0, use_doorbell: bool = False,
                 devices: Optional[Sequence[int]] = None):
        self._lib = load()
        self.provider = provider
        err = C.create_string_buffer(512)
        if devices is not None and len(devices) > 1:
            arr = (C.c_int * len(devices))(*devices)
            self._s = self._lib.ligh_scheduler_new_devices(provider._p, arr, len(devices), max_pods, max_adapters,
                                                           max_batch, flush_size, batch_window_us, refresh_interval_ms,
                                                           seed, err, 512)
        else:
            self._s = self._lib.ligh_scheduler_new2(provider._p, device, max_pods, max_adapters, max_batch,
                                                    flush_size, batch_window_us, refresh_interval_ms, seed,
                                                    int(busy_poll), caller_spin_us, int(use_doorbell), err, 512)
        if not self._s:
            raise HostSchedulerError(err.value.decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_s", None):
            self._lib.ligh_scheduler_free(self._s)
            self._s = None

    def Schedule(self, model: str, resolved: str, critical: bool) -> Tuple[int, Optional[Pod], str]:  # noqa: N802
        name, addr = C.create_string_buffer(256), C.create_string_buffer(256)
        err = C.create_string_buffer(512)
        code = self._lib.ligh_schedule(self._s, model.encode(), resolved.encode(), int(critical),
                                       name, 256, addr, 256, err, 512)
        if code == GRPC_OK:
            return code, Pod(name.value.decode(), addr.value.decode()), ""
        return code, None, err.value.decode("utf-8", "replace")

    def Refresh(self) -> None:  # noqa: N802
        err = C.create_string_buffer(512)
        if self._lib.ligh_refresh(self._s, err, 512) != 0:
            raise HostSchedulerError(err.value.decode())

    def stats(self) -> dict:
        out = (C.c_uint64 * 9)()
        self._lib.ligh_stats(self._s, out)
        return dict(zip(("scheduled", "batches", "max_batch", "refreshes", "stale_retries", "failed_refreshes",
                         "excluded_pods", "delta_refreshes", "last_dirty_pods"), map(int, out)))

    def ScheduleModel(self, datastore: "HostDataStore", model: str):  # noqa: N802
        """resolve (FetchModelData + RandomWeightedDraw + IsCritical) + Schedule: (code, resolved, pod, err)."""
        resolved, name, addr = C.create_string_buffer(256), C.create_string_buffer(256), C.create_string_buffer(256)
        err = C.create_string_buffer(512)
        code = self._lib.ligh_schedule_model(self._s, datastore._d, model.encode(), resolved, 256, name, 256, addr, 256, err, 512)
        pod = Pod(name.value.decode(), addr.value.decode()) if code == GRPC_OK else None
        return code, resolved.value.decode(), pod, err.value.decode("utf-8", "replace")

    def refresh_timing(self) -> dict:
        out = (C.c_double * 2)()
        self._lib.ligh_refresh_timing(self._s, out)
        return {"pack_us": float(out[0]), "upload_us": float(out[1])}

    def flush_timing(self) -> dict:
        """Slowest single device call and slowest Flush (resolve + call + wake) since creation, us."""
        out = (C.c_double * 4)()
        self._lib.ligh_flush_timing(self._s, out)
        return {"max_device_call_us": float(out[0]), "max_flush_us": float(out[1]), "slowest_call_batch": int(out[2]),
                "slowest_call_cpu_us": float(out[3])}

    def schedule_concurrent(self, n_threads: int, per_thread: int, models: Sequence[str],
                            critical: Sequence[bool]):
        n = n_threads * per_thread
        codes = np.zeros(n, dtype=np.int32)
        pods = np.zeros(n, dtype=np.int32)
        crit = np.array([int(c) for c in critical], dtype=np.int32)
        rc = self._lib.ligh_schedule_concurrent(self._s, n_threads, per_thread, _strs(models),
                                                crit.ctypes.data, len(models), codes.ctypes.data,
                                                pods.ctypes.data)
        assert rc == 0
        return codes, pods

    def stream_bench(self, rate: float, seconds: float, n_threads: int, models: Sequence[str],
                     critical: Sequence[bool], seed: int = 1):
        cap = int(rate * seconds * 1.2) + 1024
        lat = np.zeros(cap, dtype=np.float32)
        svc = np.zeros(cap, dtype=np.float32)
        crit = np.array([int(c) for c in critical], dtype=np.int32)
        n_done, n_err = C.c_int(), C.c_int()
        rc = self._lib.ligh_stream_bench(self._s, rate, seconds, n_threads, _strs(models),
                                         crit.ctypes.data, len(models), seed, lat.ctypes.data, svc.ctypes.data, cap,
                                         C.byref(n_done), C.byref(n_err))
        assert rc == 0
        self.last_service_latency_us = svc[: n_done.value].copy()
        return lat[: n_done.value].copy(), n_err.value


class HostDataStore:
    """backend.FakeDataStore on the C++ side: model name -> InferenceModel."""

    def __init__(self, models=()):
        self._lib = load()
        self._d = self._lib.ligh_datastore_new()
        for m in models:
            self.set_model(m)

    def set_model(self, m) -> None:
        tms = m.Spec.TargetModels
        w = np.array([t.Weight for t in tms], dtype=np.int32)
        rc = self._lib.ligh_datastore_set_model(self._d, m.Spec.ModelName.encode(), int(m.Spec.Criticality == "Critical"),
                                                len(tms), _strs([t.Name for t in tms]), w.ctypes.data if len(tms) else None)
        assert rc == 0

    def close(self):
        if self._d:
            self._lib.ligh_datastore_free(self._d)
            self._d = None
