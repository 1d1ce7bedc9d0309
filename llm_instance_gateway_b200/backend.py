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
