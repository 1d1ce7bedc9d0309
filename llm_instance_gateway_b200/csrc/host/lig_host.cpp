// lig_host.cpp — see lig_host.hpp.  Talks to the device only through the C ABI of include/lig.h.
#include "lig_host.hpp"

#include <time.h>
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../../include/lig.h"

namespace lig {
namespace scheduling {

namespace {

Status Errorf(int code, const std::string& msg) { return Status{code, msg}; }

Status LigFailure(const char* what) {
  return Errorf(Internal, std::string(what) + ": " + lig_last_error());
}

// status.Errorf(codes.ResourceExhausted, ...) printed through %w            scheduler.go:87,117
const char kDropInner[] =
    "rpc error: code = ResourceExhausted desc = dropping request due to limited backend resources";

inline uint64_t splitmix_next(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace

Status NewScheduler(std::shared_ptr<PodMetricsProvider> pmp, const Options& opt,
                    std::unique_ptr<Scheduler>* out) {
  if (!pmp || !out) return Errorf(Internal, "NewScheduler: nil provider");
  std::unique_ptr<Scheduler> s(new Scheduler());
  s->pmp_ = std::move(pmp);
  s->opt_ = opt;
  lig_thresholds t{opt.kv_cache_threshold, opt.queue_threshold_critical, opt.queueing_threshold_lora};
  if (opt.devices.size() > 1) {
    if (opt.use_doorbell) return Errorf(Internal, "the doorbell stream is a single-device mode");
    if (lig_group_create(&s->group_, opt.devices.data(), (int)opt.devices.size(), opt.max_pods, opt.max_adapters,
                         opt.max_batch) != 0)
      return LigFailure("lig_group_create");
    if (lig_group_set_thresholds(s->group_, &t) != 0) return LigFailure("lig_group_set_thresholds");
  } else {
    const int dev = opt.devices.size() == 1 ? opt.devices[0] : opt.device;
    if (lig_create(&s->ctx_, dev, opt.max_pods, opt.max_adapters, opt.max_batch) != 0)
      return LigFailure("lig_create");
    if (lig_set_thresholds(s->ctx_, &t) != 0) return LigFailure("lig_set_thresholds");
  }
  s->h_reqs_ = static_cast<lig_req*>(lig_host_alloc((size_t)opt.max_batch * sizeof(lig_req)));
  s->h_picks_ = static_cast<lig_pick*>(lig_host_alloc((size_t)opt.max_batch * sizeof(lig_pick)));
  if (!s->h_reqs_ || !s->h_picks_) return LigFailure("lig_host_alloc");
  uint64_t seed = opt.seed;
  if (seed == 0) {
    std::random_device rd;
    seed = ((uint64_t)rd() << 32) ^ rd();
  }
  s->seed_ = seed;
  s->rng_state_ = seed ^ 0xD1B54A32D192ED03ull;
  Status st = s->Refresh();
  if (!st.ok()) return st;
  if (opt.use_doorbell && lig_stream_open(s->ctx_) != 0) return LigFailure("lig_stream_open");
  s->batcher_ = std::thread(&Scheduler::BatcherLoop, s.get());
  if (opt.refresh_interval.count() > 0) s->refresher_ = std::thread(&Scheduler::RefresherLoop, s.get());
  *out = std::move(s);
  return Status{};
}

Scheduler::~Scheduler() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
    stop_flag_.store(true);
  }
  cv_.notify_all();
  if (batcher_.joinable()) batcher_.join();
  if (refresher_.joinable()) refresher_.join();
  if (ctx_ && opt_.use_doorbell) lig_stream_close(ctx_);
  if (h_reqs_) lig_host_free(h_reqs_);
  if (h_picks_) lig_host_free(h_picks_);
  if (ctx_) lig_destroy(ctx_);
  if (group_) lig_group_destroy(group_);
}

// One pack per refresh tick replaces the per-request AllPodMetrics() of scheduler.go:114-115.
// A failed refresh keeps the previous snapshot — like a failed scrape keeps stale metrics
// (backend/provider.go:151-156) — but never silently: it is counted and its reason kept.
Status Scheduler::Refresh() {
  Status st = RefreshImpl();
  if (!st.ok()) {
    std::lock_guard<std::mutex> sk(stats_mu_);
    stats_.failed_refreshes++;
    last_refresh_error_ = st.message;
    fprintf(stderr, "lig: snapshot refresh failed, keeping the previous one: %s\n", st.message.c_str());
  }
  return st;
}

std::string Scheduler::last_refresh_error() const {
  std::lock_guard<std::mutex> sk(stats_mu_);
  return last_refresh_error_;
}

Status Scheduler::RefreshImpl() {
  std::lock_guard<std::mutex> rk(refresh_mu_);
  const auto t_begin = std::chrono::steady_clock::now();
  auto pods = pmp_->AllPodMetrics();
  // A pod whose metrics do not fit the device record (WaitingQueueSize outside int32,
  // len(ActiveModels) above the record's limit) is left out of this snapshot and counted; one bad
  // scrape must not freeze the whole pool on a stale snapshot.
  uint64_t excluded = 0;
  {
    size_t keep = 0;
    for (size_t i = 0; i < pods.size(); ++i) {
      const auto& m = pods[i]->metrics;
      const bool fits = m.WaitingQueueSize >= INT32_MIN && m.WaitingQueueSize <= INT32_MAX &&
                        m.ActiveModels.size() <= (size_t)LIG_MAX_ADAPTERS;
      if (fits) pods[keep++] = pods[i];
      else ++excluded;
    }
    pods.resize(keep);
  }
  const int P = (int)pods.size();
  const int W = (P + 31) / 32;
  auto snap = std::make_shared<Snapshot>();
  snap->pods.reserve(P);
  std::vector<double> kv(P);
  std::vector<int64_t> q(P), na(P), ma(P);
  if (!intern_) intern_ = std::make_shared<InternTable>();
  bool table_changed = false;
  std::vector<char> memo_dirty;
  for (int attempt = 0; attempt < 2; ++attempt) {
    memo_.resize((size_t)P);
    memo_dirty.assign((size_t)P, 0);
    snap->pods.clear();
    for (int p = 0; p < P; ++p) {
      const backend::PodMetrics& pm = *pods[p];
      snap->pods.push_back(pm.pod);
      kv[p] = pm.metrics.KVCacheUsagePercent;
      q[p] = pm.metrics.WaitingQueueSize;
      na[p] = (int64_t)pm.metrics.ActiveModels.size();
      ma[p] = pm.metrics.MaxActiveModels;
      PodMemo& m = memo_[(size_t)p];
      bool same = m.names.size() == pm.metrics.ActiveModels.size();
      if (same) {
        size_t k = 0;
        for (const auto& kvp : pm.metrics.ActiveModels)     // std::map: ordered keys
          if (m.names[k++] != kvp.first) { same = false; break; }
      }
      if (same) continue;
      memo_dirty[(size_t)p] = 1;
      m.names.clear();
      m.ids.clear();
      for (const auto& kvp : pm.metrics.ActiveModels) {
        auto it = intern_->find(kvp.first);
        if (it == intern_->end()) {
          if (!table_changed) {                              // copy on write: older snapshots keep theirs
            intern_ = std::make_shared<InternTable>(*intern_);
            table_changed = true;
          }
          it = intern_->emplace(kvp.first, (int)intern_->size()).first;
        }
        m.names.push_back(kvp.first);
        m.ids.push_back(it->second);
      }
    }
    if ((int)intern_->size() <= opt_.max_adapters || attempt == 1) break;
    // the grow-only table outgrew the device capacity (adapters churned): re-intern from scratch
    intern_ = std::make_shared<InternTable>();
    table_changed = true;
    memo_.clear();
  }
  const int A = (int)intern_->size();
  snap->A = A;
  snap->adapter_ids = intern_;
  std::vector<uint32_t> bitmap((size_t)A * W, 0u);
  for (int p = 0; p < P; ++p)
    for (int id : memo_[(size_t)p].ids) bitmap[(size_t)id * W + (p >> 5)] |= 1u << (p & 31);
  std::vector<int32_t> q32(P);
  std::vector<uint16_t> na16(P), ma16(P);
  if (lig_pack_pods(P, q.data(), na.data(), ma.data(), q32.data(), na16.data(), ma16.data()) != 0)
    return LigFailure("lig_pack_pods");
  snap->epoch = next_epoch_++;
  const auto t_packed = std::chrono::steady_clock::now();
  // Which pods changed since the last uploaded tick?  With the adapter table and the pool size
  // unchanged and few dirty pods, only the delta crosses PCIe (lig_update_snapshot); the base is
  // the previous epoch, still resident in the other slot.
  std::vector<int32_t> dirty;
  const bool comparable = !group_ && prev_epoch_ != 0 && !table_changed && prev_A_ == A && (int)prev_q_.size() == P;
  if (comparable) {
    for (int p = 0; p < P; ++p)
      if (memo_dirty[(size_t)p] || prev_q_[(size_t)p] != q32[(size_t)p] || prev_na_[(size_t)p] != na16[(size_t)p] ||
          prev_ma_[(size_t)p] != ma16[(size_t)p] || memcmp(&prev_kv_[(size_t)p], &kv[(size_t)p], sizeof(double)) != 0)
        dirty.push_back(p);
  }
  bool used_delta = false;
  if (comparable && (int)dirty.size() * 4 <= P) {
    std::vector<double> dkv(dirty.size());
    std::vector<int32_t> dq(dirty.size()), doff(dirty.size() + 1, 0), dids;
    std::vector<uint16_t> dna(dirty.size()), dma(dirty.size());
    for (size_t i = 0; i < dirty.size(); ++i) {
      const size_t p = (size_t)dirty[i];
      dkv[i] = kv[p]; dq[i] = q32[p]; dna[i] = na16[p]; dma[i] = ma16[p];
      dids.insert(dids.end(), memo_[p].ids.begin(), memo_[p].ids.end());
      doff[i + 1] = (int32_t)dids.size();
    }
    const int rc = lig_update_snapshot(ctx_, snap->epoch, prev_epoch_, (int)dirty.size(), dirty.data(), dkv.data(),
                                       dq.data(), dna.data(), dma.data(), doff.data(), dids.data());
    if (rc == 0) used_delta = true;
    else if (rc != LIG_ERR_INVALID && rc != LIG_ERR_STALE_EPOCH) return LigFailure("lig_update_snapshot");
  }
  if (!used_delta) {
    const int rc = group_ ? lig_group_upload_snapshot(group_, snap->epoch, P, A, kv.data(), q32.data(), na16.data(),
                                                      ma16.data(), bitmap.data())
                          : lig_upload_snapshot(ctx_, snap->epoch, P, A, kv.data(), q32.data(), na16.data(),
                                                ma16.data(), bitmap.data());
    if (rc != 0) return LigFailure(group_ ? "lig_group_upload_snapshot" : "lig_upload_snapshot");
  }
  prev_kv_ = kv;
  prev_q_ = q32;
  prev_na_ = na16;
  prev_ma_ = ma16;
  prev_A_ = A;
  prev_epoch_ = snap->epoch;
  {
    std::lock_guard<std::mutex> lk(snap_mu_);
    snap_ = std::move(snap);
  }
  const auto t_done = std::chrono::steady_clock::now();
  std::lock_guard<std::mutex> sk(stats_mu_);
  stats_.refreshes++;
  stats_.delta_refreshes += used_delta ? 1 : 0;
  stats_.last_dirty_pods = comparable ? dirty.size() : (uint64_t)P;
  stats_.excluded_pods = excluded;
  stats_.last_pack_us = std::chrono::duration<double, std::micro>(t_packed - t_begin).count();
  stats_.last_upload_us = std::chrono::duration<double, std::micro>(t_done - t_packed).count();
  return Status{};
}

void Scheduler::RefresherLoop() {
  std::unique_lock<std::mutex> lk(mu_);
  while (!stop_) {
    cv_.wait_for(lk, opt_.refresh_interval, [&] { return stop_; });
    if (stop_) break;
    lk.unlock();
    Refresh();   // failures are counted and logged there
    lk.lock();
  }
}

Status Scheduler::Schedule(const LLMRequest& req, backend::Pod* targetPod) {
  Waiter w;
  w.req = &req;
  {
    std::lock_guard<std::mutex> lk(mu_);
    if (stop_) return Errorf(Internal, "scheduler is shut down");
    if (pending_.empty()) oldest_ = std::chrono::steady_clock::now();
    pending_.push_back(&w);
    pending_count_.store((int)pending_.size(), std::memory_order_release);
  }
  if (!opt_.busy_poll) cv_.notify_all();
  if (opt_.caller_spin.count() > 0) {
    const auto until = std::chrono::steady_clock::now() + opt_.caller_spin;
    while (w.done.load(std::memory_order_acquire) == 0 && std::chrono::steady_clock::now() < until) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  w.done.wait(0, std::memory_order_acquire);
  // the notifier may still be inside notify_one() on this stack object: leave only once it says
  // it is finished (done == 2); a handful of spins at most
  while (w.done.load(std::memory_order_acquire) != 2) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if (w.status.ok() && targetPod) *targetPod = std::move(w.pod);
  return w.status;
}

void Scheduler::BatcherLoop() {
  std::vector<Waiter*> batch;
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    if (opt_.busy_poll) {
      lk.unlock();
      while (pending_count_.load(std::memory_order_acquire) == 0) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if (stop_flag_.load(std::memory_order_relaxed)) break;
      }
      lk.lock();
      if (pending_.empty()) {
        if (stop_) break;
        continue;
      }
      const auto deadline = oldest_ + opt_.batch_window;
      while ((int)pending_.size() < opt_.flush_size && std::chrono::steady_clock::now() < deadline && !stop_) {
        lk.unlock();
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        lk.lock();
      }
    } else {
      cv_.wait(lk, [&] { return stop_ || !pending_.empty(); });
      if (pending_.empty() && stop_) break;
      const auto deadline = oldest_ + opt_.batch_window;
      cv_.wait_until(lk, deadline, [&] { return stop_ || (int)pending_.size() >= opt_.flush_size; });
    }
    batch.clear();
    batch.swap(pending_);
    pending_count_.store(0, std::memory_order_release);
    lk.unlock();
    Flush(batch);
    lk.lock();
  }
}

void Scheduler::Flush(std::vector<Waiter*>& batch) {
  size_t done = 0;
  const auto flush_t0 = std::chrono::steady_clock::now();
  double device_us = 0, device_cpu_us = 0;
  auto thread_cpu_us = [] {
    timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return (double)ts.tv_sec * 1e6 + (double)ts.tv_nsec * 1e-3;
  };
  while (done < batch.size()) {
    const int n = (int)std::min(batch.size() - done, (size_t)opt_.max_batch);
    std::shared_ptr<const Snapshot> snap;
    Status failure;
    int rc = 0;
    for (int attempt = 0; attempt < 4; ++attempt) {
      {
        std::lock_guard<std::mutex> lk(snap_mu_);
        snap = snap_;
      }
      for (int i = 0; i < n; ++i) {
        const LLMRequest& r = *batch[done + i]->req;
        auto it = snap->adapter_ids->find(r.ResolvedTargetModel);
        // a model in no pod's ActiveModels: id A matches no pod, like a Go map miss (filter.go:170).
        // An id interned after this snapshot was packed (>= A) is equally "in no pod" for it.
        h_reqs_[i].adapter_id = (it == snap->adapter_ids->end() || it->second >= snap->A) ? snap->A : it->second;
        h_reqs_[i].flags = r.Critical ? LIG_REQ_CRITICAL : 0u;
        h_reqs_[i].rand_key = splitmix_next(rng_state_);
      }
      const auto call_t0 = std::chrono::steady_clock::now();
      const double cpu_t0 = thread_cpu_us();
      rc = group_ ? lig_group_schedule_batch(group_, snap->epoch, seed_, h_reqs_, n, h_picks_)
           : (opt_.use_doorbell && n <= lig_stream_capacity())
               ? lig_stream_submit(ctx_, snap->epoch, seed_, h_reqs_, n, h_picks_)
               : lig_schedule_batch(ctx_, snap->epoch, seed_, h_reqs_, n, h_picks_);
      const double call_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - call_t0).count();
      if (call_us > device_us) {
        device_us = call_us;
        device_cpu_us = thread_cpu_us() - cpu_t0;
      }
      if (rc != LIG_ERR_STALE_EPOCH) break;   // two refreshes raced past this batch: re-resolve
      std::lock_guard<std::mutex> sk(stats_mu_);
      stats_.stale_retries++;
    }
    if (rc != 0) failure = LigFailure("lig_schedule_batch");
    for (int i = 0; i < n; ++i) {
      Waiter* w = batch[done + i];
      if (rc != 0) {
        w->status = failure;
      } else if (h_picks_[i].status == LIG_OK) {
        w->pod = snap->pods[(size_t)h_picks_[i].pod_idx];          // pods[i].Pod  scheduler.go:121
      } else if (h_picks_[i].status == LIG_DROP) {
        w->status = Errorf(ResourceExhausted,                         // scheduler.go:117
                           std::string("failed to apply filter, resulted 0 pods, this should never "
                                       "happen: ") + kDropInner);
      } else {
        w->status = Errorf(Unknown, "failed to apply filter, resulted 0 pods, this should never "
                                    "happen: %!w(<nil>)");
      }
    }
    {
      std::lock_guard<std::mutex> sk(stats_mu_);
      stats_.scheduled += (uint64_t)n;
      stats_.batches++;
      stats_.max_batch = std::max<uint64_t>(stats_.max_batch, (uint64_t)n);
    }
    for (int i = 0; i < n; ++i) {
      Waiter* w = batch[done + i];
      w->done.store(1, std::memory_order_release);
      w->done.notify_one();
      w->done.store(2, std::memory_order_release);   // last access: the caller may now destroy *w
    }
    done += (size_t)n;
  }
  const double flush_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - flush_t0).count();
  std::lock_guard<std::mutex> sk(stats_mu_);
  if (device_us > stats_.max_device_call_us) {
    stats_.max_device_call_us = device_us;
    stats_.slowest_call_cpu_us = device_cpu_us;
    stats_.slowest_call_batch = stats_.batches ? stats_.batches - 1 : 0;
  }
  stats_.max_flush_us = std::max(stats_.max_flush_us, flush_us);
}

Status Scheduler::ScheduleModel(backend::ModelDataStore& datastore, const std::string& model,
                                std::string* resolvedTargetModel, backend::Pod* targetPod) {
  auto modelObj = datastore.FetchModelData(model);                              // request.go:42-45
  if (!modelObj) return Errorf(Unknown, "error finding a model object in InferenceModel for input " + model);
  std::string modelName = model;
  if (!modelObj->TargetModels.empty()) {                                        // request.go:46-51
    const uint64_t key = model_draws_.fetch_add(1, std::memory_order_relaxed) + 1;
    modelName = backend::RandomWeightedDraw(*modelObj, backend::SplitMixSource{seed_ ^ key ^ LIG_DRAW_DOMAIN});
    if (modelName.empty()) return Errorf(Unknown, "error getting target model name for model " + modelObj->ModelName);
  }
  LLMRequest req;                                                               // request.go:52-56
  req.Model = model;
  req.ResolvedTargetModel = modelName;
  req.Critical = modelObj->Critical;
  Status st = Schedule(req, targetPod);
  if (!st.ok()) st.message = "failed to find target pod: " + st.message;        // request.go:73-75 (%w keeps the code)
  if (resolvedTargetModel) *resolvedTargetModel = modelName;
  return st;
}

Stats Scheduler::stats() const {
  std::lock_guard<std::mutex> sk(stats_mu_);
  return stats_;
}

}  // namespace scheduling

namespace backend {

uint64_t SplitMixSource::Uint64() {
  uint64_t z = (state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int32_t SplitMixSource::Int31n(int32_t n) {            // math/rand (Go 1.22) Rand.Int31n; n > 0
  if ((n & (n - 1)) == 0) return Int31() & (n - 1);
  const int32_t mx = (int32_t)((1u << 31) - 1 - (uint32_t)((1ull << 31) % (uint64_t)n));
  int32_t v = Int31();
  while (v > mx) v = Int31();
  return v % n;
}

std::string RandomWeightedDraw(const InferenceModel& model, SplitMixSource source) {   // datastore.go:78-98
  int32_t weights = 0;
  for (const TargetModel& tm : model.TargetModels) weights = (int32_t)((uint32_t)weights + (uint32_t)tm.Weight);
  if (weights <= 0) return "";                          // Go's Int31n panics here
  int32_t randomVal = source.Int31n(weights);
  for (const TargetModel& tm : model.TargetModels) {
    if (randomVal < tm.Weight) return tm.Name;
    randomVal -= tm.Weight;
  }
  return "";
}

}  // namespace backend
}  // namespace lig
