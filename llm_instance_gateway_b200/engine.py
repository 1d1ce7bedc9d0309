"""Thin object wrapper over the C ABI (one ``lig_ctx`` = one CUDA device)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from .packer import MPICK_DTYPE, PICK_DTYPE, REQ_DTYPE, PackedModels, PackedSnapshot, _ptr


class Engine:
    def __init__(self, device: int = 0, max_pods: int = 4096, max_adapters: int = 1024,
                 max_batch: int = 1 << 20, _borrowed_ctx=None):
        self._lib = N.load()
        self._owned = _borrowed_ctx is None
        if self._owned:
            self._ctx = C.c_void_p()
            N.check(self._lib.lig_create(C.byref(self._ctx), device, max_pods, max_adapters, max_batch))
        else:
            self._ctx = C.c_void_p(_borrowed_ctx)      # a member of an EngineGroup
        self.device = device
        self.max_batch = max_batch

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            if self._owned:
                self._lib.lig_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- thresholds (scheduler.go:15-24) ----
    def set_thresholds(self, kv_cache_threshold=0.8, queue_threshold_critical=5,
                       queueing_threshold_lora=50) -> None:
        t = N.LigThresholds(kv_cache_threshold, queue_threshold_critical, queueing_threshold_lora)
        N.check(self._lib.lig_set_thresholds(self._ctx, C.byref(t)))

    # ---- snapshots ----
    def upload_snapshot(self, epoch: int, snap: PackedSnapshot, block: bool = True) -> None:
        """block=False: lig_upload_snapshot_async (staged + enqueued, the device is not waited for)."""
        fn = self._lib.lig_upload_snapshot if block else self._lib.lig_upload_snapshot_async
        N.check(fn(self._ctx, epoch, snap.P, snap.A, _ptr(snap.kv), _ptr(snap.q), _ptr(snap.n_active),
                   _ptr(snap.max_active), _ptr(snap.bitmap)))

    def update_snapshot(self, new_epoch: int, base_epoch: int, pod_idx, kv, q, n_active, max_active,
                        adapter_offsets, adapter_ids) -> None:
        """Delta upload: base_epoch with the listed pods replaced becomes new_epoch."""
        a = [np.ascontiguousarray(pod_idx, dtype=np.int32), np.ascontiguousarray(kv, dtype=np.float64),
             np.ascontiguousarray(q, dtype=np.int32), np.ascontiguousarray(n_active, dtype=np.uint16),
             np.ascontiguousarray(max_active, dtype=np.uint16), np.ascontiguousarray(adapter_offsets, dtype=np.int32),
             np.ascontiguousarray(adapter_ids, dtype=np.int32)]
        N.check(self._lib.lig_update_snapshot(self._ctx, new_epoch, base_epoch, len(a[0]), *[_ptr(x) for x in a]))

    def upload_snapshot_device(self, epoch: int, P: int, A: int, d_blob: int, stream: int = 0) -> None:
        N.check(self._lib.lig_upload_snapshot_device(self._ctx, epoch, P, A, d_blob, stream or None))

    # ---- host-buffer hot path ----
    def schedule_batch(self, epoch: int, seed: int, reqs: np.ndarray,
                       out: Optional[np.ndarray] = None) -> np.ndarray:
        assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
        R = int(reqs.shape[0])
        if out is None:
            out = np.empty(R, dtype=PICK_DTYPE)
        N.check(self._lib.lig_schedule_batch(self._ctx, epoch, seed, _ptr(reqs), R, _ptr(out)))
        return out

    def schedule_batch_ptr(self, epoch: int, seed: int, h_reqs: int, R: int, h_out: int) -> None:
        N.check(self._lib.lig_schedule_batch(self._ctx, epoch, seed, h_reqs, R, h_out))

    def schedule_batch_async(self, epoch: int, seed: int, h_reqs: int, R: int, h_out: int) -> int:
        """Page-locked buffers only; returns a ticket for schedule_wait."""
        t = C.c_int(-1)
        N.check(self._lib.lig_schedule_batch_async(self._ctx, epoch, seed, h_reqs, R, h_out, C.byref(t)))
        return t.value

    def schedule_wait(self, ticket: int) -> None:
        N.check(self._lib.lig_schedule_wait(self._ctx, ticket))

    # ---- model requests: the resolve step of HandleRequestBody on the device ----
    def upload_models(self, epoch: int, models: PackedModels, block: bool = True) -> None:
        fn = self._lib.lig_upload_models if block else self._lib.lig_upload_models_async
        N.check(fn(self._ctx, epoch, models.n_models, _ptr(models.target_offsets),
                   _ptr(models.target_adapter_ids), _ptr(models.target_weights),
                   _ptr(models.critical), _ptr(models.self_adapter_ids), _ptr(models.present)))

    def schedule_models_batch(self, epoch: int, seed: int, model_ids: np.ndarray, first_index: int = 0,
                              out: Optional[np.ndarray] = None) -> np.ndarray:
        assert model_ids.dtype == np.uint32 and model_ids.flags.c_contiguous
        R = int(model_ids.shape[0])
        if out is None:
            out = np.empty(R, dtype=MPICK_DTYPE)
        N.check(self._lib.lig_schedule_models_batch(self._ctx, epoch, seed, first_index, _ptr(model_ids), R, _ptr(out)))
        return out

    def schedule_models_batch_ptr(self, epoch: int, seed: int, first_index: int, h_ids: int, R: int, h_out: int) -> None:
        N.check(self._lib.lig_schedule_models_batch(self._ctx, epoch, seed, first_index, h_ids, R, h_out))

    def schedule_models_batches_device(self, epoch: int, seed: int, first_index: int, d_ids_ptrs, R: int,
                                       d_out_ptrs, stream: int = 0) -> None:
        n = len(d_ids_ptrs)
        a = (C.c_void_p * n)(*d_ids_ptrs)
        b = (C.c_void_p * n)(*d_out_ptrs)
        N.check(self._lib.lig_schedule_models_batches_device(self._ctx, epoch, seed, first_index, a, R, b, n,
                                                             stream or None))

    def resolve_models(self, epoch: int, seed: int, model_ids: np.ndarray, first_index: int = 0):
        assert model_ids.dtype == np.uint32 and model_ids.flags.c_contiguous
        R = int(model_ids.shape[0])
        reqs = np.empty(R, dtype=REQ_DTYPE)
        out = np.empty(R, dtype=MPICK_DTYPE)
        N.check(self._lib.lig_resolve_models(self._ctx, epoch, seed, first_index, _ptr(model_ids), R,
                                             _ptr(reqs), _ptr(out)))
        return reqs, out

    def schedule_batch_feedback_device(self, epoch: int, seed: int, d_reqs: int, R: int, d_out: int, sub_batch: int,
                                       n_windows: int = 0, d_hist: int = 0, stream: int = 0) -> None:
        """Opt-in in-batch load feedback: windows of `sub_batch` requests, the picks of each window
        (all ranks) are added to the pods' queue sizes before the next window is scheduled."""
        N.check(self._lib.lig_schedule_batch_feedback_device(self._ctx, epoch, seed, d_reqs, R, d_out, sub_batch,
                                                             n_windows, d_hist or None, stream or None))

    def pick_kernel_info(self, epoch: int) -> dict:
        name = C.create_string_buffer(64)
        grid, threads, tb, in_smem = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        N.check(self._lib.lig_pick_kernel_info(self._ctx, epoch, name, 64, C.byref(grid), C.byref(threads),
                                               C.byref(tb), C.byref(in_smem)))
        return {"kernel": name.value.decode(), "grid": grid.value, "threads": threads.value,
                "table_bytes": tb.value, "tables_in_smem": bool(in_smem.value)}

    # ---- one process per GPU: NCCL inside the library ----
    def comm_unique_id(self) -> bytes:
        buf = (C.c_char * N.LIG_COMM_ID_BYTES)()
        N.check(self._lib.lig_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, n_ranks: int, rank: int, unique_id: bytes) -> None:
        buf = (C.c_char * N.LIG_COMM_ID_BYTES).from_buffer_copy(unique_id)
        N.check(self._lib.lig_comm_init_rank(self._ctx, n_ranks, rank, buf))

    def comm_upload_snapshot_device(self, epoch: int, P: int, A: int, d_blob: int, root: int = 0,
                                    stream: int = 0) -> None:
        N.check(self._lib.lig_comm_upload_snapshot_device(self._ctx, epoch, P, A, d_blob or None, root,
                                                          stream or None))

    def comm_upload_snapshot(self, epoch: int, P: int, A: int, snap: Optional[PackedSnapshot], root: int = 0) -> None:
        if snap is None:
            N.check(self._lib.lig_comm_upload_snapshot(self._ctx, epoch, P, A, None, None, None, None, None, root))
        else:
            N.check(self._lib.lig_comm_upload_snapshot(self._ctx, epoch, P, A, _ptr(snap.kv), _ptr(snap.q),
                                                       _ptr(snap.n_active), _ptr(snap.max_active),
                                                       _ptr(snap.bitmap), root))

    def comm_allreduce_i32(self, d_values: int, n: int, stream: int = 0) -> None:
        N.check(self._lib.lig_comm_allreduce_i32(self._ctx, d_values, n, stream or None))

    def schedule_scan(self, epoch: int, seed: int, reqs: np.ndarray, want_masks: bool = True,
                      W: int = 0) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
        R = int(reqs.shape[0])
        out = np.empty(R, dtype=PICK_DTYPE)
        masks = np.zeros((R, W), dtype=np.uint32) if want_masks else None
        N.check(self._lib.lig_schedule_scan(self._ctx, epoch, seed, _ptr(reqs), R, _ptr(out),
                                            _ptr(masks) if masks is not None else None))
        return out, masks

    # ---- HBM-resident hot path (device pointers, async on `stream`) ----
    def schedule_batch_device(self, epoch: int, seed: int, d_reqs: int, R: int, d_out: int,
                              stream: int = 0) -> None:
        N.check(self._lib.lig_schedule_batch_device(self._ctx, epoch, seed, d_reqs, R, d_out,
                                                    stream or None))

    def schedule_batches_device(self, epoch: int, seed: int, d_reqs_ptrs, R: int, d_out_ptrs,
                                stream: int = 0) -> None:
        """Queue len(d_reqs_ptrs) resident batches back to back (one pass through the ABI)."""
        n = len(d_reqs_ptrs)
        assert n == len(d_out_ptrs)
        a = (C.c_void_p * n)(*d_reqs_ptrs)
        b = (C.c_void_p * n)(*d_out_ptrs)
        N.check(self._lib.lig_schedule_batches_device(self._ctx, epoch, seed, a, R, b, n,
                                                      stream or None))

    def schedule_scan_device(self, epoch: int, seed: int, d_reqs: int, R: int, d_out: int,
                             d_masks: int = 0, stream: int = 0) -> None:
        N.check(self._lib.lig_schedule_scan_device(self._ctx, epoch, seed, d_reqs, R, d_out,
                                                   d_masks or None, stream or None))

    # ---- streaming doorbell (persistent kernel) ----
    def stream_open(self) -> None:
        N.check(self._lib.lig_stream_open(self._ctx))

    def stream_close(self) -> None:
        N.check(self._lib.lig_stream_close(self._ctx))

    def stream_submit(self, epoch: int, seed: int, reqs: np.ndarray) -> np.ndarray:
        assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
        out = np.empty(len(reqs), dtype=PICK_DTYPE)
        N.check(self._lib.lig_stream_submit(self._ctx, epoch, seed, _ptr(reqs), len(reqs), _ptr(out)))
        return out

    def read_class(self, epoch: int, critical: bool, adapter_id: int, P: int):
        status, n = C.c_int(), C.c_int()
        lst = np.zeros(max(P, 1), dtype=np.uint16)
        N.check(self._lib.lig_read_class(self._ctx, epoch, int(bool(critical)), adapter_id,
                                         C.byref(status), C.byref(n), _ptr(lst)))
        return status.value, n.value, lst[: n.value].copy()

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.lig_kernel_launches(self._ctx))

    @property
    def sm_count(self) -> int:
        return int(self._lib.lig_sm_count(self._ctx))


class EngineGroup:
    """Several GPUs owned by one process (lig_group_*): the snapshot is replicated with one in-library
    ncclBroadcast per upload, the request batch shards contiguously, results stay in request order."""

    def __init__(self, devices, max_pods: int = 4096, max_adapters: int = 1024, max_batch: int = 1 << 20):
        self._lib = N.load()
        self._g = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        N.check(self._lib.lig_group_create(C.byref(self._g), arr, len(devices), max_pods, max_adapters, max_batch))
        self.devices = list(devices)
        self.max_batch = max_batch

    def close(self) -> None:
        if getattr(self, "_g", None):
            self._lib.lig_group_destroy(self._g)
            self._g = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self) -> int:
        return int(self._lib.lig_group_size(self._g))

    def member(self, i: int) -> Engine:
        ctx = self._lib.lig_group_ctx(self._g, i)
        if not ctx:
            raise IndexError(i)
        return Engine(self.devices[i], max_batch=self.max_batch, _borrowed_ctx=ctx)

    def set_thresholds(self, kv_cache_threshold=0.8, queue_threshold_critical=5, queueing_threshold_lora=50) -> None:
        t = N.LigThresholds(kv_cache_threshold, queue_threshold_critical, queueing_threshold_lora)
        N.check(self._lib.lig_group_set_thresholds(self._g, C.byref(t)))

    def upload_snapshot(self, epoch: int, snap: PackedSnapshot) -> None:
        N.check(self._lib.lig_group_upload_snapshot(self._g, epoch, snap.P, snap.A, _ptr(snap.kv), _ptr(snap.q),
                                                    _ptr(snap.n_active), _ptr(snap.max_active), _ptr(snap.bitmap)))

    def schedule_batch(self, epoch: int, seed: int, reqs: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        assert reqs.dtype == REQ_DTYPE and reqs.flags.c_contiguous
        R = int(reqs.shape[0])
        if out is None:
            out = np.empty(R, dtype=PICK_DTYPE)
        N.check(self._lib.lig_group_schedule_batch(self._g, epoch, seed, _ptr(reqs), R, _ptr(out)))
        return out

    def schedule_batch_ptr(self, epoch: int, seed: int, h_reqs: int, R: int, h_out: int) -> None:
        N.check(self._lib.lig_group_schedule_batch(self._g, epoch, seed, h_reqs, R, h_out))
