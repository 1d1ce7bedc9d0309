"""Host-side mirror of the reference's ``scheduling`` package interface, backed by the GPU.

Same names, argument meaning and error behaviour as the reference so that it drops in behind
``handlers.Scheduler`` (pkg/ext-proc/handlers/server.go:37-39):

    scheduler = NewScheduler(pod_metrics_provider)       # scheduling/scheduler.go:93-99
    pod = scheduler.Schedule(LLMRequest(...))            # scheduling/scheduler.go:113-122

``Schedule`` returns a ``backend.Pod`` or raises ``StatusError``; its ``.code`` is
``"ResourceExhausted"`` when the tree shed the request (scheduler.go:83-89), which the ext-proc
server turns into HTTP 429 (handlers/server.go:97-109).  ``ScheduleBatch`` is the micro-batched
form the Go adapter uses under load (one C-ABI call per batch).  There is no CPU path here:
without liblig.so or without a CUDA device construction fails.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Protocol, Sequence, Tuple

import numpy as np

from . import _native as N
from .backend import Pod, PodMetrics
from .engine import Engine
from .packer import REQ_DTYPE, PackedSnapshot, pack_pod_metrics


@dataclass
class LLMRequest:                            # scheduling/types.go:4-11
    Model: str = ""
    TargetModels: Dict[str, int] = field(default_factory=dict)
    ResolvedTargetModel: str = ""
    Critical: bool = False


class PodMetricsProvider(Protocol):          # scheduling/scheduler.go:108-110
    def AllPodMetrics(self) -> Sequence[PodMetrics]: ...


class StatusError(Exception):
    """A Go error carrying a gRPC status code (google.golang.org/grpc/status)."""

    def __init__(self, code: str, message: str):
        super().__init__(message)
        self.code = code


_DROP_DESC = "dropping request due to limited backend resources"          # scheduler.go:87


def _schedule_error(status: int) -> StatusError:
    # fmt.Errorf("failed to apply filter, resulted %v pods, this should never happen: %w", len(pods), err)
    #                                                                       scheduler.go:117
    if status == N.LIG_DROP:
        inner = f"rpc error: code = ResourceExhausted desc = {_DROP_DESC}"
        return StatusError("ResourceExhausted",
                           f"failed to apply filter, resulted 0 pods, this should never happen: {inner}")
    # tree returned ([], nil): Go prints a nil error under %w as %!w(<nil>); status.Code(err) of a
    # plain fmt error is Unknown.
    return StatusError("Unknown",
                       "failed to apply filter, resulted 0 pods, this should never happen: %!w(<nil>)")


class Scheduler:
    def __init__(self, podMetricsProvider: PodMetricsProvider, device: int = 0,
                 max_pods: int = 4096, max_adapters: int = 1024, max_batch: int = 1 << 16,
                 seed: Optional[int] = None):
        self.podMetricsProvider = podMetricsProvider
        self._engine = Engine(device, max_pods, max_adapters, max_batch)
        self._epoch = 0
        self._snap: Optional[PackedSnapshot] = None
        self._rng = random.Random(seed)      # the Go global source is auto-seeded (scheduler.go:120)
        self._seed = self._rng.getrandbits(64)

    def close(self) -> None:
        self._engine.close()

    # The reference re-reads the provider on every Schedule call; here the snapshot is re-packed
    # once per refresh tick (refreshMetricsInterval, main.go:39).  A provider that exposes
    # ``Version()`` is re-packed only when the version changes; otherwise every call re-packs,
    # which is exactly the reference's per-call semantics.
    def Refresh(self) -> PackedSnapshot:
        snap = pack_pod_metrics(list(self.podMetricsProvider.AllPodMetrics()))
        self._epoch += 1
        self._engine.upload_snapshot(self._epoch, snap)
        self._snap = snap
        self._version = self._provider_version()
        return snap

    def _provider_version(self):
        v = getattr(self.podMetricsProvider, "Version", None)
        return v() if callable(v) else None

    def _current(self) -> PackedSnapshot:
        ver = self._provider_version()
        if self._snap is None or ver is None or ver != self._version:
            return self.Refresh()
        return self._snap

    def ScheduleBatch(self, reqs: Sequence[LLMRequest]) -> List[Tuple[Optional[Pod], Optional[StatusError]]]:
        snap = self._current()
        arr = np.zeros(len(reqs), dtype=REQ_DTYPE)
        for i, r in enumerate(reqs):
            arr[i] = (snap.adapter_id(r.ResolvedTargetModel), N.LIG_REQ_CRITICAL if r.Critical else 0,
                      self._rng.getrandbits(64))
        picks = self._engine.schedule_batch(self._epoch, self._seed, arr)
        out = []
        for pk in picks:
            if pk["status"] == N.LIG_OK:
                out.append((snap.pods[int(pk["pod_idx"])], None))
            else:
                out.append((None, _schedule_error(int(pk["status"]))))
        return out

    def Schedule(self, req: LLMRequest) -> Pod:                     # scheduler.go:113-122
        pod, err = self.ScheduleBatch([req])[0]
        if err is not None:
            raise err
        return pod

    # Filter-level view (the GPU analogue of Filter.Filter, filter.go:12-15): survivor slice.
    def Filter(self, req: LLMRequest) -> Tuple[List[Pod], Optional[StatusError]]:
        snap = self._current()
        arr = np.zeros(1, dtype=REQ_DTYPE)
        arr[0] = (snap.adapter_id(req.ResolvedTargetModel), N.LIG_REQ_CRITICAL if req.Critical else 0, 0)
        picks, masks = self._engine.schedule_scan(self._epoch, self._seed, arr, True, snap.W)
        st = int(picks[0]["status"])
        survivors = [snap.pods[p] for p in range(snap.P) if (int(masks[0, p >> 5]) >> (p & 31)) & 1]
        if st == N.LIG_DROP:
            return [], StatusError("ResourceExhausted", _DROP_DESC)
        return survivors, None


def NewScheduler(pmp: PodMetricsProvider, **kw) -> Scheduler:       # scheduler.go:93-99
    return Scheduler(pmp, **kw)
