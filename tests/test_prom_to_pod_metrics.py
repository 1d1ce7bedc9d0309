"""The field path in front of the snapshot packer: scraped vLLM metric families
-> backend.PodMetrics, pinned to the reference's own cases (backend/vllm/metrics_test.go:14-203),
plus the Clone quirk of backend/types.go:37-53 that decides what the scheduler sees."""
from llm_instance_gateway_b200 import backend as B
from llm_instance_gateway_b200.backend import Metrics, Pod, PodMetrics, promToPodMetrics


def families(max_lora):
    g = lambda pairs: [{"value": v, "timestamp_ms": ts} for v, ts in pairs]
    return {
        B.RunningQueueSizeMetricName: g([(10, 100), (15, 200)]),          # the second is the latest
        B.WaitingQueueSizeMetricName: g([(20, 100), (25, 200)]),
        B.KVCacheUsagePercentMetricName: g([(0.8, 100), (0.9, 200)]),
        B.LoraRequestInfoMetricName: [
            {"value": 100, "labels": {B.LoraRequestInfoRunningAdaptersMetricName: "lora3,lora4",
                                      B.LoraRequestInfoMaxAdaptersMetricName: max_lora}},
            {"value": 90, "labels": {B.LoraRequestInfoRunningAdaptersMetricName: "lora2",
                                     B.LoraRequestInfoMaxAdaptersMetricName: "2"}},
        ],
    }


def test_all_metrics_available():                                          # metrics_test.go:22-120
    updated, errs = promToPodMetrics(families("2"), PodMetrics())
    assert errs == []
    assert updated.Metrics == Metrics(RunningQueueSize=15, WaitingQueueSize=25, KVCacheUsagePercent=0.9,
                                      ActiveModels={"lora3": 0, "lora4": 0}, MaxActiveModels=2)


def test_invalid_max_lora():                                               # metrics_test.go:121-219
    updated, errs = promToPodMetrics(families("2a"), PodMetrics())
    assert len(errs) == 1 and 'strconv.Atoi: parsing "2a": invalid syntax' in errs[0]
    assert updated.Metrics == Metrics(RunningQueueSize=15, WaitingQueueSize=25, KVCacheUsagePercent=0.9,
                                      ActiveModels={"lora3": 0, "lora4": 0}, MaxActiveModels=0)


def test_clone_quirk_and_missing_families():
    """existing.Clone() drops MaxActiveModels (types.go:37-53): without a lora series the updated
    metrics report 0 adapters' room even if the previous scrape said 4; other fields persist."""
    existing = PodMetrics(Pod("p", "a"), Metrics(ActiveModels={"x": 1}, MaxActiveModels=4, WaitingQueueSize=7,
                                                 KVCacheUsagePercent=0.5, RunningQueueSize=3))
    updated, errs = promToPodMetrics({}, existing)
    assert len(errs) == 4 and all("not found" in e for e in errs)
    assert updated.Pod == existing.Pod
    assert updated.Metrics.MaxActiveModels == 0                           # the quirk
    assert (updated.Metrics.WaitingQueueSize, updated.Metrics.KVCacheUsagePercent, updated.Metrics.ActiveModels) == (7, 0.5, {"x": 1})
    # an empty running_lora_adapters label clears ActiveModels; a gauge is truncated like int(float64)
    fam = {B.WaitingQueueSizeMetricName: [{"value": 3.9}],
           B.LoraRequestInfoMetricName: [{"value": 5, "labels": {B.LoraRequestInfoRunningAdaptersMetricName: "",
                                                                   B.LoraRequestInfoMaxAdaptersMetricName: "8"}}]}
    updated, errs = promToPodMetrics(fam, existing)
    assert updated.Metrics.WaitingQueueSize == 3 and updated.Metrics.ActiveModels == {} and updated.Metrics.MaxActiveModels == 8
    assert len(errs) == 2
