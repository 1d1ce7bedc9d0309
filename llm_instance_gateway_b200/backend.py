"""Host-side mirror of the reference's ``backend`` record types (the hot path's input).

Field names are the reference's Go field names on purpose, so that tests written against this
package read like ``pkg/ext-proc/scheduling/filter_test.go``:

    PodMetrics(Pod=Pod(Name="pod1"),
               Metrics=Metrics(WaitingQueueSize=0, KVCacheUsagePercent=0.2, MaxActiveModels=2,
                               ActiveModels={"foo": 1, "bar": 1}))

Reference: pkg/ext-proc/backend/types.go:8-31 (Pod, Metrics, PodMetrics), :37-53 (Clone).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass(frozen=True)
class Pod:                                   # backend/types.go:8-11
    Name: str = ""
    Address: str = ""

    def __str__(self) -> str:                # backend/types.go:13-15
        return self.Name + ":" + self.Address


@dataclass
class Metrics:                               # backend/types.go:17-26
    ActiveModels: Dict[str, int] = field(default_factory=dict)
    MaxActiveModels: int = 0
    RunningQueueSize: int = 0
    WaitingQueueSize: int = 0
    KVCacheUsagePercent: float = 0.0
    KvCacheMaxTokenCapacity: int = 0


@dataclass
class PodMetrics:                            # backend/types.go:28-31
    Pod: Pod = field(default_factory=Pod)
    Metrics: Metrics = field(default_factory=Metrics)

    def Clone(self) -> "PodMetrics":         # backend/types.go:37-53
        # The reference's Clone does not copy MaxActiveModels; kept as is (it is not on the
        # scheduling path, only the scraper calls it).
        m = self.Metrics
        return PodMetrics(Pod=self.Pod, Metrics=Metrics(
            ActiveModels=dict(m.ActiveModels), RunningQueueSize=m.RunningQueueSize,
            WaitingQueueSize=m.WaitingQueueSize, KVCacheUsagePercent=m.KVCacheUsagePercent,
            KvCacheMaxTokenCapacity=m.KvCacheMaxTokenCapacity))


# ---- api/v1alpha1/inferencemodel_types.go: the fields the request path reads --------------------
CRITICAL, DEFAULT, SHEDDABLE = "Critical", "Default", "Sheddable"     # inferencemodel_types.go:105-111


@dataclass
class TargetModel:                           # inferencemodel_types.go: TargetModel{Name, Weight}
    Name: str = ""
    Weight: int = 0


@dataclass
class InferenceModelSpec:                    # inferencemodel_types.go:40-69
    ModelName: str = ""
    Criticality: Optional[str] = None
    TargetModels: List[TargetModel] = field(default_factory=list)


@dataclass
class InferenceModel:
    Name: str = ""
    Spec: InferenceModelSpec = field(default_factory=InferenceModelSpec)


def IsCritical(model: InferenceModel) -> bool:          # backend/datastore.go:100-105
    return model.Spec.Criticality is not None and model.Spec.Criticality == CRITICAL


class FakeDataStore:                         # backend/fake.go: FakeDataStore{Res map[string]*InferenceModel}
    def __init__(self, Res: Optional[Dict[str, InferenceModel]] = None):
        self.Res = dict(Res or {})

    def FetchModelData(self, modelName: str) -> Optional[InferenceModel]:   # backend/datastore.go:70-76
        return self.Res.get(modelName)


_M64 = (1 << 64) - 1
DRAW_DOMAIN = 0xA0761D6478BD642F            # LIG_DRAW_DOMAIN, include/lig.h


class SplitMixSource:
    """rand.Source64 over SplitMix64: the injected source the ABI defines the draw and the pick on."""

    def __init__(self, state: int):
        self.state = state & _M64

    def Uint64(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & _M64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def Int63(self) -> int:
        return self.Uint64() >> 1

    def Int31(self) -> int:                  # math/rand: Int31() = Int63() >> 32
        return self.Int63() >> 32

    def Int31n(self, n: int) -> int:         # math/rand (Go 1.22) Rand.Int31n
        if n <= 0:
            raise ValueError("invalid argument to Int31n")      # Go panics
        if n & (n - 1) == 0:
            return self.Int31() & (n - 1)
        mx = (1 << 31) - 1 - ((1 << 31) % n)
        v = self.Int31()
        while v > mx:
            v = self.Int31()
        return v % n


def RandomWeightedDraw(model: InferenceModel, source: SplitMixSource) -> str:      # backend/datastore.go:78-98
    """The reference seeds a fresh source per call (rand.NewSource(rand.Int63()), unseeded); here the
    source is injected (include/lig.h: the request's private stream)."""
    weights = 0
    for tm in model.Spec.TargetModels:
        weights += tm.Weight
    randomVal = source.Int31n(weights)
    for tm in model.Spec.TargetModels:
        if randomVal < tm.Weight:
            return tm.Name
        randomVal -= tm.Weight
    return ""


# ---- backend/vllm/metrics.go:73-166: scraped metric families -> PodMetrics -------------------------
LoraRequestInfoMetricName = "vllm:lora_requests_info"
LoraRequestInfoRunningAdaptersMetricName = "running_lora_adapters"
LoraRequestInfoMaxAdaptersMetricName = "max_lora"
RunningQueueSizeMetricName = "vllm:num_requests_running"
WaitingQueueSizeMetricName = "vllm:num_requests_waiting"
KVCacheUsagePercentMetricName = "vllm:gpu_cache_usage_perc"


def _go_atoi(s: str) -> int:
    """strconv.Atoi: optional sign, decimal digits only (no spaces, no underscores)."""
    body = s[1:] if s[:1] in "+-" else s
    if not body or not all("0" <= ch <= "9" for ch in body):
        raise ValueError(f'strconv.Atoi: parsing "{s}": invalid syntax')
    return int(s)


def getLatestMetric(metricFamilies: Dict[str, list], metricName: str):        # metrics.go:151-171
    """A metric family is a list of dicts {"value": float, "timestamp_ms": int, "labels": {...}}.
    vLLM sets no timestamps, so ">=" makes this the LAST series of the family."""
    mf = metricFamilies.get(metricName)
    if mf is None:
        raise KeyError(f'metric family "{metricName}" not found')
    if len(mf) == 0:
        raise KeyError(f'no metrics available for "{metricName}"')
    latestTs, latest = 0, None
    for m in mf:
        if m.get("timestamp_ms", 0) >= latestTs:
            latestTs, latest = m.get("timestamp_ms", 0), m
    return latest


def getLatestLoraMetric(metricFamilies: Dict[str, list]):                     # metrics.go:131-148
    """The series with the largest VALUE (the value is the series' creation timestamp)."""
    mf = metricFamilies.get(LoraRequestInfoMetricName)
    if mf is None:
        raise KeyError(f'metric family "{LoraRequestInfoMetricName}" not found')
    latestTs, latest = 0.0, None
    for m in mf:
        if m.get("value", 0.0) > latestTs:
            latestTs, latest = m.get("value", 0.0), m
    return latest


def promToPodMetrics(metricFamilies: Dict[str, list], existing: PodMetrics):   # metrics.go:73-129
    """Returns (updated PodMetrics, list of error strings).  Starts from existing.Clone(), which
    does NOT carry MaxActiveModels over (backend/types.go:37-53): a scrape without a parsable
    max_lora label leaves MaxActiveModels at 0 — the quirk the scheduler then sees."""
    errs: List[str] = []
    updated = existing.Clone()
    for name, field_, conv in ((RunningQueueSizeMetricName, "RunningQueueSize", int),
                               (WaitingQueueSizeMetricName, "WaitingQueueSize", int),
                               (KVCacheUsagePercentMetricName, "KVCacheUsagePercent", float)):
        try:
            m = getLatestMetric(metricFamilies, name)
            setattr(updated.Metrics, field_, conv(m["value"]))       # int(float64): truncation toward zero
        except KeyError as ex:
            errs.append(str(ex.args[0]))
    lora = None
    try:
        lora = getLatestLoraMetric(metricFamilies)
    except KeyError as ex:
        errs.append(str(ex.args[0]))
    if lora is not None:
        updated.Metrics.ActiveModels = {}
        for lname, lvalue in lora.get("labels", {}).items():
            if lname == LoraRequestInfoRunningAdaptersMetricName and lvalue != "":
                for adapter in lvalue.split(","):
                    updated.Metrics.ActiveModels[adapter] = 0
            if lname == LoraRequestInfoMaxAdaptersMetricName and lvalue != "":
                try:
                    updated.Metrics.MaxActiveModels = _go_atoi(lvalue)
                except ValueError as ex:
                    updated.Metrics.MaxActiveModels = 0              # strconv.Atoi returns 0 with the error
                    errs.append(str(ex))
    return updated, errs
