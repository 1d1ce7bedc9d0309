"""ctypes binding of the C ABI in include/lig.h (liblig.so, built in-tree for sm_100a).

There is no fallback: if the library is missing, import fails loudly with the build command; if
there is no CUDA device, ``lig_create`` returns LIG_ERR_CUDA and ``LigError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblig.so")

LIG_OK, LIG_DROP, LIG_EMPTY, LIG_NO_MODEL = 0, 1, 2, 3
LIG_DRAW_DOMAIN = 0xA0761D6478BD642F
LIG_ERR_INVALID, LIG_ERR_CUDA, LIG_ERR_STALE_EPOCH, LIG_ERR_NO_SNAPSHOT, LIG_ERR_RANGE = -1, -2, -3, -4, -5
LIG_ERR_BUSY, LIG_ERR_NCCL = -6, -7
LIG_MAX_TICKETS, LIG_COMM_ID_BYTES = 256, 128
LIG_ABI_VERSION = 2
LIG_REQ_CRITICAL = 1
LIG_MAX_PODS, LIG_MAX_ADAPTERS = 32768, 65534

# Every symbol include/lig.h declares (tests check the library exports all of them).
EXPORTED_SYMBOLS = (
    "lig_create", "lig_destroy", "lig_set_thresholds", "lig_get_thresholds", "lig_snapshot_bytes",
    "lig_pack_pods", "lig_pack_snapshot", "lig_upload_snapshot", "lig_upload_snapshot_device",
    "lig_schedule_batch", "lig_schedule_batch_device", "lig_schedule_batches_device", "lig_schedule_scan_device",
    "lig_schedule_scan", "lig_read_class", "lig_last_error", "lig_version", "lig_abi_version",
    "lig_device_count", "lig_kernel_launches", "lig_sm_count", "lig_host_alloc", "lig_host_free",
    "lig_stream_capacity", "lig_stream_open", "lig_stream_submit", "lig_stream_close",
    "lig_schedule_batch_async", "lig_schedule_wait",
    "lig_group_create", "lig_group_destroy", "lig_group_size", "lig_group_ctx", "lig_group_set_thresholds",
    "lig_group_upload_snapshot", "lig_group_schedule_batch",
    "lig_comm_unique_id", "lig_comm_init_rank", "lig_comm_upload_snapshot_device", "lig_comm_upload_snapshot",
    "lig_comm_allreduce_i32",
    "lig_upload_models", "lig_schedule_models_batch", "lig_schedule_models_batches_device", "lig_resolve_models",
    "lig_pick_kernel_info", "lig_schedule_batch_feedback_device", "lig_update_snapshot",
    "lig_upload_snapshot_async", "lig_upload_models_async",
)


class LigReq(C.Structure):
    _fields_ = [("adapter_id", C.c_int32), ("flags", C.c_uint32), ("rand_key", C.c_uint64)]


class LigPick(C.Structure):
    _fields_ = [("pod_idx", C.c_int32), ("status", C.c_uint16), ("n_survivors", C.c_uint16)]


class LigThresholds(C.Structure):
    _fields_ = [("kv_cache_threshold", C.c_double), ("queue_threshold_critical", C.c_int64),
                ("queueing_threshold_lora", C.c_int64)]


class LigError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lig error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load liblig.so (once).  Raises if it has not been built — no silent fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C llm_instance_gateway_b200/csrc` (needs nvcc, sm_100a). "
            "There is no CPU fallback for the scheduling path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    lib.lig_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    lib.lig_destroy.argtypes = [vp]
    lib.lig_destroy.restype = None
    lib.lig_set_thresholds.argtypes = [vp, C.POINTER(LigThresholds)]
    lib.lig_get_thresholds.argtypes = [vp, C.POINTER(LigThresholds)]
    lib.lig_snapshot_bytes.argtypes = [i32, i32]
    lib.lig_snapshot_bytes.restype = C.c_size_t
    lib.lig_pack_pods.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    lib.lig_pack_snapshot.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.lig_upload_snapshot.argtypes = [vp, u64, i32, i32, vp, vp, vp, vp, vp]
    lib.lig_upload_snapshot_async.argtypes = [vp, u64, i32, i32, vp, vp, vp, vp, vp]
    lib.lig_upload_snapshot_device.argtypes = [vp, u64, i32, i32, vp, vp]
    lib.lig_schedule_batch.argtypes = [vp, u64, u64, vp, i32, vp]
    lib.lig_schedule_batch_device.argtypes = [vp, u64, u64, vp, i32, vp, vp]
    lib.lig_schedule_batches_device.argtypes = [vp, u64, u64, vp, i32, vp, i32, vp]
    lib.lig_schedule_scan_device.argtypes = [vp, u64, u64, vp, i32, vp, vp, vp]
    lib.lig_schedule_scan.argtypes = [vp, u64, u64, vp, i32, vp, vp]
    lib.lig_read_class.argtypes = [vp, u64, i32, i32, C.POINTER(i32), C.POINTER(i32), vp]
    lib.lig_last_error.restype = C.c_char_p
    lib.lig_version.restype = C.c_char_p
    lib.lig_kernel_launches.argtypes = [vp]
    lib.lig_kernel_launches.restype = u64
    lib.lig_sm_count.argtypes = [vp]
    lib.lig_host_alloc.argtypes = [C.c_size_t]
    lib.lig_host_alloc.restype = vp
    lib.lig_host_free.argtypes = [vp]
    lib.lig_host_free.restype = None
    lib.lig_stream_open.argtypes = [vp]
    lib.lig_stream_submit.argtypes = [vp, u64, u64, vp, i32, vp]
    lib.lig_stream_close.argtypes = [vp]
    lib.lig_schedule_batch_async.argtypes = [vp, u64, u64, vp, i32, vp, C.POINTER(i32)]
    lib.lig_schedule_wait.argtypes = [vp, i32]
    lib.lig_group_create.argtypes = [C.POINTER(vp), C.POINTER(i32), i32, i32, i32, i32]
    lib.lig_group_destroy.argtypes = [vp]
    lib.lig_group_destroy.restype = None
    lib.lig_group_size.argtypes = [vp]
    lib.lig_group_ctx.argtypes = [vp, i32]
    lib.lig_group_ctx.restype = vp
    lib.lig_group_set_thresholds.argtypes = [vp, C.POINTER(LigThresholds)]
    lib.lig_group_upload_snapshot.argtypes = [vp, u64, i32, i32, vp, vp, vp, vp, vp]
    lib.lig_group_schedule_batch.argtypes = [vp, u64, u64, vp, i32, vp]
    lib.lig_comm_unique_id.argtypes = [vp]
    lib.lig_comm_init_rank.argtypes = [vp, i32, i32, vp]
    lib.lig_comm_upload_snapshot_device.argtypes = [vp, u64, i32, i32, vp, i32, vp]
    lib.lig_comm_upload_snapshot.argtypes = [vp, u64, i32, i32, vp, vp, vp, vp, vp, i32]
    lib.lig_comm_allreduce_i32.argtypes = [vp, vp, i32, vp]
    lib.lig_upload_models.argtypes = [vp, u64, i32, vp, vp, vp, vp, vp, vp]
    lib.lig_upload_models_async.argtypes = [vp, u64, i32, vp, vp, vp, vp, vp, vp]
    lib.lig_schedule_models_batch.argtypes = [vp, u64, u64, u64, vp, i32, vp]
    lib.lig_schedule_models_batches_device.argtypes = [vp, u64, u64, u64, vp, i32, vp, i32, vp]
    lib.lig_resolve_models.argtypes = [vp, u64, u64, u64, vp, i32, vp, vp]
    lib.lig_schedule_batch_feedback_device.argtypes = [vp, u64, u64, vp, i32, vp, i32, i32, vp, vp]
    lib.lig_update_snapshot.argtypes = [vp, u64, u64, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.lig_pick_kernel_info.argtypes = [vp, u64, C.c_char_p, i32, C.POINTER(i32), C.POINTER(i32),
                                         C.POINTER(i32), C.POINTER(i32)]
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)  # AttributeError if the library does not export it
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise LigError(rc, load().lig_last_error().decode("utf-8", "replace"))
