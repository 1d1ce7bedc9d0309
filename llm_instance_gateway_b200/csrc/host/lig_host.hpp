// lig_host.hpp — native host side above the C ABI (include/lig.h): the C++ counterpart of the Go
// adapter a maintainer adds to the reference.  It mirrors the reference's
// scheduling-package interface for the hot path — same names, argument meaning, error behaviour:
//
//   backend::Pod / Metrics / PodMetrics     pkg/ext-proc/backend/types.go:8-31
//   scheduling::LLMRequest                   pkg/ext-proc/scheduling/types.go:4-11
//   scheduling::PodMetricsProvider           pkg/ext-proc/scheduling/scheduler.go:108-110
//   scheduling::NewScheduler / Scheduler     pkg/ext-proc/scheduling/scheduler.go:93-122
//
// Scheduler::Schedule is goroutine-style safe: many threads call it concurrently and block for
// their pick (one caller per Envoy stream, pkg/ext-proc/handlers/request.go:72); a batcher thread
// folds the concurrent calls into one lig_schedule_batch call per flush.  The snapshot is
// re-packed once per refresh tick (refreshMetricsInterval, pkg/ext-proc/main.go:39) instead of
// once per request.  No CPU scheduling path exists here: without liblig.so + a CUDA device,
// NewScheduler fails.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

struct lig_ctx;
struct lig_group;
struct lig_req;
struct lig_pick;

namespace lig {

namespace backend {

struct Pod {                         // backend/types.go:8-11
  std::string Name;
  std::string Address;
  std::string String() const { return Name + ":" + Address; }   // types.go:13-15
};

struct Metrics {                     // backend/types.go:17-26
  std::map<std::string, int> ActiveModels;
  int64_t MaxActiveModels = 0;
  int64_t RunningQueueSize = 0;
  int64_t WaitingQueueSize = 0;
  double KVCacheUsagePercent = 0.0;
  int64_t KvCacheMaxTokenCapacity = 0;
};

struct PodMetrics {                  // backend/types.go:28-31
  Pod pod;
  Metrics metrics;
};

// api/v1alpha1/inferencemodel_types.go: the fields the request path reads
struct TargetModel {
  std::string Name;
  int32_t Weight = 0;
};
struct InferenceModel {
  std::string ModelName;             // Spec.ModelName: the key requests carry in "model"
  bool Critical = false;             // Spec.Criticality != nil && *Spec.Criticality == Critical
  std::vector<TargetModel> TargetModels;
};

class ModelDataStore {               // handlers/server.go:47-49
 public:
  virtual ~ModelDataStore() = default;
  virtual std::shared_ptr<const InferenceModel> FetchModelData(const std::string& modelName) = 0;
};

// A rand.Source over SplitMix64: the injected source include/lig.h defines the draw and the pick on.
struct SplitMixSource {
  uint64_t state;
  uint64_t Uint64();
  int32_t Int31() { return (int32_t)(Uint64() >> 33); }          // Int63() >> 32
  int32_t Int31n(int32_t n);                                      // math/rand (Go 1.22)
};

// RandomWeightedDraw(model, seed)                                 backend/datastore.go:78-98
// Returns "" like the reference when no target is hit; weights summing to <= 0 (where Go's
// Int31n panics) also return "".
std::string RandomWeightedDraw(const InferenceModel& model, SplitMixSource source);

}  // namespace backend

namespace scheduling {

struct LLMRequest {                  // scheduling/types.go:4-11
  std::string Model;
  std::map<std::string, int> TargetModels;
  std::string ResolvedTargetModel;
  bool Critical = false;
};

// google.golang.org/grpc/codes values the path can produce.
enum Code : int { OK = 0, Unknown = 2, ResourceExhausted = 8, Internal = 13 };

struct Status {                      // a Go `error` carrying a gRPC status code
  int code = OK;
  std::string message;
  bool ok() const { return code == OK; }
};

class PodMetricsProvider {           // scheduling/scheduler.go:108-110
 public:
  virtual ~PodMetricsProvider() = default;
  virtual std::vector<std::shared_ptr<const backend::PodMetrics>> AllPodMetrics() = 0;
};

struct Options {
  int device = 0;
  // More than one entry: ONE scheduler owns all listed GPUs (lig_group_*): the snapshot is
  // replicated by an in-library ncclBroadcast per refresh, every flushed batch is sharded by request.
  std::vector<int> devices;
  int max_pods = 4096;
  int max_adapters = 1024;
  int max_batch = 1 << 16;
  // A batch is flushed when it has flush_size requests or its oldest request is batch_window old.
  int flush_size = 4096;
  std::chrono::microseconds batch_window{50};
  uint64_t seed = 0;                 // 0 = seed from std::random_device (Go's auto-seeded source)
  // >0: a refresher thread re-packs the provider's snapshot at this period.
  std::chrono::milliseconds refresh_interval{0};
  // Latency knobs: the batcher thread polls its queue instead of sleeping on a condition
  // variable, and callers spin this long on their result before blocking in the kernel.
  // Flush micro-batches through the persistent doorbell kernel (lig_stream_submit) instead of a
  // kernel launch per flush.  Nothing else in the process may then synchronise the whole device.
  bool use_doorbell = false;
  bool busy_poll = false;
  std::chrono::microseconds caller_spin{0};
  double kv_cache_threshold = 0.8;   // scheduler.go:15-24
  int64_t queue_threshold_critical = 5;
  int64_t queueing_threshold_lora = 50;
};

struct Stats {
  uint64_t scheduled = 0;      // Schedule calls completed
  uint64_t batches = 0;        // lig_schedule_batch calls issued
  uint64_t max_batch = 0;      // largest batch flushed
  uint64_t refreshes = 0;      // snapshots uploaded
  uint64_t stale_retries = 0;  // batches re-resolved after LIG_ERR_STALE_EPOCH
  uint64_t delta_refreshes = 0;    // refreshes that went through lig_update_snapshot (dirty pods only)
  uint64_t last_dirty_pods = 0;    // pods whose metrics changed at the last refresh
  uint64_t failed_refreshes = 0;   // Refresh() calls that kept the previous snapshot (see last_refresh_error)
  uint64_t excluded_pods = 0;      // pods left out of the last snapshot: a metric did not fit the device record
  double last_pack_us = 0;     // last Refresh: provider slice -> columns + bitmap (host)
  double last_upload_us = 0;   // last Refresh: lig_upload_snapshot (H2D + class tables)
  double max_device_call_us = 0;   // slowest single lig_schedule_batch call (launch + kernel + synchronise)
  double max_flush_us = 0;         // slowest Flush: resolve + device call + waking the callers
  uint64_t slowest_call_batch = 0; // which batch (0-based count of device calls) that slowest call was
  double slowest_call_cpu_us = 0;  // CPU time the batcher thread burnt inside that call: ~ its wall time when it
                                   // was spinning on a busy device, far less when the thread itself was descheduled
};

class Scheduler {
 public:
  ~Scheduler();
  Scheduler(const Scheduler&) = delete;
  Scheduler& operator=(const Scheduler&) = delete;

  // Schedule finds the target pod based on metrics and the requested lora adapter.
  // Returns OK and fills *targetPod, or: ResourceExhausted "dropping request due to limited
  // backend resources" wrapped exactly like scheduler.go:117 (-> HTTP 429 at
  // handlers/server.go:97-109); Unknown for the "resulted 0 pods" case; Internal for a
  // batch-level (CUDA) failure.                                       scheduler.go:113-122
  Status Schedule(const LLMRequest& req, backend::Pod* targetPod);

  // The resolve step of HandleRequestBody (handlers/request.go:42-56) followed by Schedule:
  // FetchModelData, RandomWeightedDraw when TargetModels is non-empty, IsCritical.  Fills
  // *resolvedTargetModel (the body's "model" is rewritten with it, request.go:60-69).  Errors as the
  // reference: Unknown "error finding a model object in InferenceModel for input <model>" /
  // "error getting target model name for model <name>".
  Status ScheduleModel(backend::ModelDataStore& datastore, const std::string& model, std::string* resolvedTargetModel,
                       backend::Pod* targetPod);

  // Re-read the provider and upload a new snapshot epoch (call on every metrics refresh).
  Status Refresh();
 private:
  Status RefreshImpl();
 public:

  Stats stats() const;
  std::string last_refresh_error() const;

 private:
  friend Status NewScheduler(std::shared_ptr<PodMetricsProvider>, const Options&,
                             std::unique_ptr<Scheduler>*);
  Scheduler() = default;

  using InternTable = std::unordered_map<std::string, int>;
  struct Snapshot {
    uint64_t epoch = 0;
    int A = 0;
    std::shared_ptr<const InternTable> adapter_ids;   // shared between snapshots until a name is added
    std::vector<backend::Pod> pods;
  };
  // Packer state kept between refreshes (guarded by refresh_mu_): adapter ids are stable while
  // the table fits max_adapters, and a pod whose ActiveModels keys did not change since the last
  // tick reuses its encoded ids (sequential string compares instead of hash lookups).
  struct PodMemo {
    std::vector<std::string> names;
    std::vector<int> ids;
  };
  // last uploaded columns, to find the dirty pods of a tick (delta upload)
  std::vector<double> prev_kv_;
  std::vector<int32_t> prev_q_;
  std::vector<uint16_t> prev_na_, prev_ma_;
  int prev_A_ = -1;
  uint64_t prev_epoch_ = 0;
  std::atomic<uint64_t> model_draws_{0};
  std::shared_ptr<InternTable> intern_;
  std::vector<PodMemo> memo_;
  // Lives on the caller's stack.  `done` goes 0 -> 1 (result published, the caller may read it)
  // -> 2 (the notifier will not touch the Waiter again, the caller may return and destroy it).
  struct Waiter {
    const LLMRequest* req = nullptr;
    Status status;
    backend::Pod pod;
    std::atomic<uint32_t> done{0};
  };

  void BatcherLoop();
  void RefresherLoop();
  void Flush(std::vector<Waiter*>& batch);

  std::shared_ptr<PodMetricsProvider> pmp_;
  Options opt_;
  lig_ctx* ctx_ = nullptr;
  lig_group* group_ = nullptr;     // set instead of ctx_ when Options::devices lists several GPUs
  lig_req* h_reqs_ = nullptr;      // pinned, device-mapped (lig_host_alloc)
  lig_pick* h_picks_ = nullptr;
  uint64_t seed_ = 0;
  uint64_t rng_state_ = 0;

  mutable std::mutex snap_mu_;
  std::shared_ptr<const Snapshot> snap_;
  uint64_t next_epoch_ = 1;
  std::mutex refresh_mu_;

  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<Waiter*> pending_;
  std::atomic<int> pending_count_{0};
  std::chrono::steady_clock::time_point oldest_;
  bool stop_ = false;
  std::atomic<bool> stop_flag_{false};
  std::thread batcher_, refresher_;

  mutable std::mutex stats_mu_;
  Stats stats_;
  std::string last_refresh_error_;
};

// NewScheduler                                                        scheduler.go:93-99
Status NewScheduler(std::shared_ptr<PodMetricsProvider> pmp, const Options& opt,
                    std::unique_ptr<Scheduler>* out);

}  // namespace scheduling
}  // namespace lig
