"""Pure-Python restatement of the reference scheduler hot path.  TEST INFRASTRUCTURE ONLY.

The checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg may import this module.  It is the small, obviously-correct twin of
``oracle/lig_oracle.c`` (which is the fast one); the two are cross-checked against each other by
``tests/test_oracle_cross.py`` and both are pinned against the reference's own golden vectors
(``tests/golden/go_filter_test_vectors.json``).

Restates (paths relative to the reference repo root, ``pkg/ext-proc/``):
  scheduling/scheduler.go:15-122   constants, ``defaultFilter`` tree, ``Scheduler.Schedule``
  scheduling/filter.go:44-187      node semantics, range filters, predicates
  scheduling/types.go:4-11         ``LLMRequest``
  backend/types.go:8-31            ``Pod``, ``Metrics``, ``PodMetrics``

Parity pinning: survivor sets pinned by the reference's tests; the final ``rand.Intn`` pick is
unpinned in the reference (auto-seeded global source, scheduler.go:120) and is defined by
``include/lig.h`` (SplitMix64 source + Go's ``Int31n``).
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

MAX_INT = (1 << 63) - 1            # math.MaxInt on 64-bit Go
MAX_FLOAT64 = sys.float_info.max   # math.MaxFloat64
_U64 = (1 << 64) - 1

LIGO_OK, LIGO_DROP, LIGO_EMPTY, LIGO_ERROR = 0, 1, 2, 3


# ---- backend/types.go:8-31 -------------------------------------------------------------------
@dataclass(frozen=True)
class Pod:
    name: str = ""
    address: str = ""


@dataclass
class Metrics:
    active_models: Dict[str, int] = field(default_factory=dict)
    max_active_models: int = 0
    running_queue_size: int = 0
    waiting_queue_size: int = 0
    kv_cache_usage_percent: float = 0.0
    kv_cache_max_token_capacity: int = 0


@dataclass
class PodMetrics:
    pod: Pod = field(default_factory=Pod)
    metrics: Metrics = field(default_factory=Metrics)


# ---- scheduling/types.go:4-11 ----------------------------------------------------------------
@dataclass
class LLMRequest:
    model: str = ""
    target_models: Dict[str, int] = field(default_factory=dict)
    resolved_target_model: str = ""
    critical: bool = False


class FilterError(Exception):
    """A Go ``error`` value returned by a filter func."""


class ResourceExhausted(FilterError):
    """status.Errorf(codes.ResourceExhausted, ...) — scheduler.go:87."""


# scheduler.go:15-24
@dataclass
class Thresholds:
    kv_cache_threshold: float = 0.8
    queue_threshold_critical: int = 5
    queueing_threshold_lora: int = 50


FilterFunc = Callable[[LLMRequest, List[PodMetrics]], Tuple[Optional[List[PodMetrics]], Optional[FilterError]]]


def _go_int64(x: int) -> int:
    """Wrap to a Go int (two's complement, 64 bit)."""
    x &= _U64
    return x - (1 << 64) if x >> 63 else x


def _go_div(a: int, b: int) -> int:
    """Go integer division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


# ---- filter.go:12-73 -------------------------------------------------------------------------
@dataclass
class Filter:
    name: str
    filter: FilterFunc
    next_on_success: Optional["Filter"] = None
    next_on_failure: Optional["Filter"] = None
    next_on_success_or_failure: Optional["Filter"] = None

    def Filter(self, req, pods):  # noqa: N802  (keeps the reference's method name)
        filtered, err = self.filter(req, pods)                                 # filter.go:47
        nxt = self.next_on_success_or_failure                                  # filter.go:49
        if err is None and filtered is not None and len(filtered) > 0:         # filter.go:50
            if self.next_on_success is None and self.next_on_success_or_failure is None:
                return filtered, err                                           # filter.go:51-54
            if self.next_on_success is not None:
                nxt = self.next_on_success                                     # filter.go:55-57
            return nxt.Filter(req, filtered)                                   # filter.go:60
        if self.next_on_failure is None and self.next_on_success_or_failure is None:
            return filtered, err                                               # filter.go:62-65
        if self.next_on_failure is not None:
            nxt = self.next_on_failure                                         # filter.go:66-68
        return nxt.Filter(req, pods)                                           # filter.go:71


def to_filter_func(pp) -> FilterFunc:                                          # filter.go:79-93
    def fn(req, pods):
        filtered = [pod for pod in pods if pp(req, pod)]
        if len(filtered) == 0:
            return None, FilterError("no pods left")
        return filtered, None
    return fn


def least_queuing_filter_func(req, pods):                                      # filter.go:102-122
    mn, mx = MAX_INT, 0
    for pod in pods:
        q = pod.metrics.waiting_queue_size
        if q <= mn:
            mn = q
        if q >= mx:
            mx = q
    filtered = []
    for pod in pods:
        q = pod.metrics.waiting_queue_size
        thr = _go_int64(mn + _go_div(_go_int64(mx - mn), len(pods)))
        if q >= mn and q <= thr:
            filtered.append(pod)
    return filtered, None


def least_kv_cache_filter_func(req, pods):                                     # filter.go:134-154
    mn, mx = MAX_FLOAT64, 0.0
    for pod in pods:
        kv = pod.metrics.kv_cache_usage_percent
        if kv <= mn:
            mn = kv
        if kv >= mx:
            mx = kv
    filtered = []
    for pod in pods:
        kv = pod.metrics.kv_cache_usage_percent
        if kv >= mn and kv <= mn + (mx - mn) / float(len(pods)):
            filtered.append(pod)
    return filtered, None


class Tree:
    """The filter tree of scheduler.go:26-91 instantiated for one set of thresholds."""

    def __init__(self, thr: Optional[Thresholds] = None):
        t = thr or Thresholds()
        self.thresholds = t

        def low_queueing_pod_predicate(req, pod):                              # filter.go:124-126
            return pod.metrics.waiting_queue_size < t.queueing_threshold_lora

        def lora_affinity_predicate(req, pod):                                 # filter.go:169-172
            return req.resolved_target_model in pod.metrics.active_models

        def can_accept_new_lora_predicate(req, pod):                           # filter.go:175-177
            return len(pod.metrics.active_models) < pod.metrics.max_active_models

        def low_lora_cost_predicate(req, pod):                                 # filter.go:163-166
            return (req.resolved_target_model in pod.metrics.active_models
                    or len(pod.metrics.active_models) < pod.metrics.max_active_models)

        def critical_request_predicate(req, pod):                              # filter.go:179-181
            return req.critical

        self.predicates = {
            "lowQueueingPodPredicate": low_queueing_pod_predicate,
            "loRAAffinityPredicate": lora_affinity_predicate,
            "canAcceptNewLoraPredicate": can_accept_new_lora_predicate,
            "lowLoRACostPredicate": low_lora_cost_predicate,
            "criticalRequestPredicate": critical_request_predicate,
        }

        def drop(req, pods):                                                   # scheduler.go:83-89
            return [], ResourceExhausted("dropping request due to limited backend resources")

        queue_lora_and_kv = Filter(                                            # scheduler.go:35-46
            "least queuing", least_queuing_filter_func,
            next_on_success_or_failure=Filter(
                "low cost LoRA", to_filter_func(low_lora_cost_predicate),
                next_on_success_or_failure=Filter("least KV cache percent",
                                                  least_kv_cache_filter_func)))
        queue_and_kv = Filter(                                                 # scheduler.go:49-56
            "least queuing", least_queuing_filter_func,
            next_on_success_or_failure=Filter("least KV cache percent",
                                              least_kv_cache_filter_func))
        low_latency = Filter(                                                  # scheduler.go:58-72
            "low queueing filter", to_filter_func(low_queueing_pod_predicate),
            next_on_success=Filter(
                "affinity LoRA", to_filter_func(lora_affinity_predicate),
                next_on_success=queue_and_kv,
                next_on_failure=Filter("can accept LoRA Adapter",
                                       to_filter_func(can_accept_new_lora_predicate),
                                       next_on_success_or_failure=queue_and_kv)),
            next_on_failure=queue_lora_and_kv)
        sheddable = Filter(                                                    # scheduler.go:74-90
            "has capacity for sheddable requests",
            to_filter_func(no_queue_and_less_than_kv_cache_threshold_predicate(
                t.queue_threshold_critical, t.kv_cache_threshold)),
            next_on_success=queue_lora_and_kv,
            next_on_failure=Filter("drop request", drop))
        self.default_filter = Filter(                                          # scheduler.go:26-31
            "critical request", to_filter_func(critical_request_predicate),
            next_on_success=low_latency, next_on_failure=sheddable)


def no_queue_and_less_than_kv_cache_threshold_predicate(queue_threshold, kv_cache_threshold):
    def pred(req, pod):                                                        # filter.go:183-187
        return (pod.metrics.waiting_queue_size <= queue_threshold
                and pod.metrics.kv_cache_usage_percent <= kv_cache_threshold)
    return pred


# ---- the pick (include/lig.h) ----------------------------------------------------------------
class SplitMix64Source:
    """A math/rand.Source: Int63() = next() >> 1."""

    def __init__(self, state: int):
        self.state = state & _U64

    def next(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & _U64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _U64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _U64
        return z ^ (z >> 31)

    def int63(self) -> int:
        return self.next() >> 1

    def int31(self) -> int:                       # Rand.Int31: int32(r.Int63() >> 32)
        return self.int63() >> 32

    def int31n(self, n: int) -> int:              # Rand.Int31n (Go 1.22 math/rand)
        if n & (n - 1) == 0:
            return self.int31() & (n - 1)
        mx = (1 << 31) - 1 - (1 << 31) % n
        v = self.int31()
        while v > mx:
            v = self.int31()
        return v % n


class Scheduler:
    """scheduler.go:93-122 with an injected provider (any object with ``AllPodMetrics()``)."""

    def __init__(self, pod_metrics_provider, thresholds: Optional[Thresholds] = None):
        self.pod_metrics_provider = pod_metrics_provider
        self.tree = Tree(thresholds)
        self.filter = self.tree.default_filter

    def filter_only(self, req: LLMRequest):
        """(status, survivors) of defaultFilter.Filter over the provider's slice."""
        pods, err = self.filter.Filter(req, self.pod_metrics_provider.AllPodMetrics())
        pods = pods or []
        if isinstance(err, ResourceExhausted):
            return LIGO_DROP, pods
        if err is not None:
            return LIGO_ERROR, pods
        return (LIGO_OK if pods else LIGO_EMPTY), pods

    def Schedule(self, req: LLMRequest, seed: int = 0, rand_key: int = 0):  # noqa: N802
        """Returns (status, pod_index_or_-1, n_survivors, survivors)."""
        all_pods = self.pod_metrics_provider.AllPodMetrics()
        pods, err = self.filter.Filter(req, all_pods)                          # scheduler.go:115
        pods = pods or []
        if err is not None or len(pods) == 0:                                  # scheduler.go:116
            if isinstance(err, ResourceExhausted):
                return LIGO_DROP, -1, 0, pods
            if err is not None:
                return LIGO_ERROR, -1, 0, pods
            return LIGO_EMPTY, -1, 0, pods
        i = SplitMix64Source(seed ^ rand_key).int31n(len(pods))                # scheduler.go:120
        chosen = pods[i]                                                       # scheduler.go:121
        idx = next(k for k, p in enumerate(all_pods) if p is chosen)
        return LIGO_OK, idx, len(pods), pods


class StaticProvider:
    """A fake PodMetricsProvider returning a fixed ordered slice (scheduler.go:108-110)."""

    def __init__(self, pods: Sequence[PodMetrics]):
        self._pods = list(pods)

    def AllPodMetrics(self):  # noqa: N802
        return list(self._pods)


# ---- the step before Schedule: handlers/request.go:42-56, backend/datastore.go:70-105 ----------------
# Readable twin of oracle/lig_oracle_models.c (draw parity against Go's seeded source is unpinned,
# see that file's header; the draw is defined on SplitMix64Source(seed ^ rand_key ^ DRAW_DOMAIN)).
DRAW_DOMAIN = 0xA0761D6478BD642F
LIGO_NO_MODEL, LIGO_NO_TARGET = 3, 4


@dataclass
class TargetModel:                      # api/v1alpha1 TargetModel
    name: str
    weight: int


@dataclass
class InferenceModel:                   # api/v1alpha1 InferenceModel, the fields the path reads
    name: str
    critical: bool = False
    target_models: List[TargetModel] = field(default_factory=list)


def random_weighted_draw(model: InferenceModel, source: SplitMix64Source) -> str:   # datastore.go:78-98
    weights = 0
    for tm in model.target_models:
        weights = _go_int32(weights + tm.weight)
    if weights <= 0:
        raise ValueError("invalid argument to Int31n")        # Go panics
    random_val = source.int31n(weights)                                              # datastore.go:90
    for tm in model.target_models:                                                   # datastore.go:91-97
        if random_val < tm.weight:
            return tm.name
        random_val -= tm.weight
    return ""


def _go_int32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x >= (1 << 31) else x


def resolve(models: Sequence[Optional[InferenceModel]], model_id: int, seed: int, rand_key: int):
    """request.go:42-56 -> (status, resolved target model name, critical, target index)."""
    if not (0 <= model_id < len(models)) or models[model_id] is None:                # request.go:42-45
        return LIGO_NO_MODEL, None, False, 255
    m = models[model_id]
    name, idx = m.name, 255
    if m.target_models:                                                              # request.go:46-51
        try:
            name = random_weighted_draw(m, SplitMix64Source(seed ^ rand_key ^ DRAW_DOMAIN))
        except ValueError:
            return LIGO_NO_TARGET, None, False, 255
        if name == "":
            return LIGO_NO_TARGET, None, False, 255
        # the index of the FIRST target the loop would stop at (names may repeat)
        src = SplitMix64Source(seed ^ rand_key ^ DRAW_DOMAIN)
        rv = src.int31n(sum(t.weight for t in m.target_models))
        idx = 0
        for k, tm in enumerate(m.target_models):
            if rv < tm.weight:
                idx = k
                break
            rv -= tm.weight
    return LIGO_OK, name, m.critical, idx
