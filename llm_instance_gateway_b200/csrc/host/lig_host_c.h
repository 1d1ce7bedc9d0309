/* lig_host_c.h — C entry points over the C++ host runtime (lig_host.hpp), for ctypes-driven tests
 * and the streaming benchmark.  Not part of the drop-in boundary (that is include/lig.h). */
#ifndef LIG_HOST_C_H_
#define LIG_HOST_C_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ligh_provider ligh_provider;   /* a mutable fake PodMetricsProvider (backend/fake.go:10-21 analogue) */
typedef struct ligh_scheduler ligh_scheduler;

ligh_provider* ligh_provider_new(void);
void ligh_provider_free(ligh_provider*);
/* Replace the provider's slice.  active_flat holds all pods' ActiveModels keys back to back;
 * pod i owns active_flat[active_offsets[i] .. active_offsets[i+1]). */
int ligh_provider_set_pods(ligh_provider*, int n, const char* const* names, const char* const* addrs,
                           const int64_t* waiting_queue_size, const double* kv_cache_usage_percent,
                           const int64_t* max_active_models, const char* const* active_flat,
                           const int* active_offsets);

ligh_scheduler* ligh_scheduler_new(ligh_provider*, int device, int max_pods, int max_adapters,
                                   int max_batch, int flush_size, int batch_window_us,
                                   int refresh_interval_ms, uint64_t seed, char* err, int err_cap);
/* Same, plus the latency knobs: busy_poll (batcher polls instead of sleeping), caller_spin_us,
 * use_doorbell (flush through the persistent doorbell kernel instead of a launch per flush). */
ligh_scheduler* ligh_scheduler_new2(ligh_provider*, int device, int max_pods, int max_adapters,
                                    int max_batch, int flush_size, int batch_window_us,
                                    int refresh_interval_ms, uint64_t seed, int busy_poll,
                                    int caller_spin_us, int use_doorbell, char* err, int err_cap);
void ligh_scheduler_free(ligh_scheduler*);
/* Scheduler.Schedule: returns the gRPC code (0 OK, 8 ResourceExhausted, 2 Unknown, 13 Internal). */
int ligh_schedule(ligh_scheduler*, const char* model, const char* resolved_target_model, int critical,
                  char* name, int name_cap, char* addr, int addr_cap, char* err, int err_cap);
int ligh_refresh(ligh_scheduler*, char* err, int err_cap);
/* several GPUs behind one scheduler (lig_group_*) */
ligh_scheduler* ligh_scheduler_new_devices(ligh_provider*, const int* devices, int n_devices, int max_pods,
                                           int max_adapters, int max_batch, int flush_size, int window_us,
                                           int refresh_ms, uint64_t seed, char* err, int err_cap);
/* a fake ModelDataStore (backend/fake.go) and the resolve + Schedule call of HandleRequestBody */
typedef struct ligh_datastore ligh_datastore;
ligh_datastore* ligh_datastore_new(void);
void ligh_datastore_free(ligh_datastore*);
int ligh_datastore_set_model(ligh_datastore*, const char* model_name, int critical, int n_targets,
                             const char* const* target_names, const int32_t* weights);
int ligh_schedule_model(ligh_scheduler*, ligh_datastore*, const char* model, char* resolved, int resolved_cap,
                        char* name, int name_cap, char* addr, int addr_cap, char* err, int err_cap);
void ligh_stats(ligh_scheduler*, uint64_t out[9]); /* scheduled, batches, max_batch, refreshes, stale_retries */
void ligh_refresh_timing(ligh_scheduler*, double out[2]); /* last Refresh: host pack us, lig_upload_snapshot us */
void ligh_flush_timing(ligh_scheduler*, double out[4]);   /* slowest device call us, slowest Flush us, index of that call, thread CPU us inside it */

/* n_threads caller threads each issue `per_thread` blocking Schedule calls (model i of the
 * request table, round-robin); out_codes/out_pod (n_threads*per_thread) receive the code and the
 * index of the returned pod in the provider's slice (-1 on error). */
int ligh_schedule_concurrent(ligh_scheduler*, int n_threads, int per_thread,
                             const char* const* resolved_models, const int* critical, int n_models,
                             int* out_codes, int* out_pod);

/* Streaming load (BASELINE.json configs[4]): Poisson arrivals at `rate` req/s for `seconds`,
 * spread over n_threads caller threads; latency = completion - scheduled arrival.  Fills
 * lat_us[0..*n_done) (capacity cap).  Returns 0. */
int ligh_stream_bench(ligh_scheduler*, double rate, double seconds, int n_threads,
                      const char* const* resolved_models, const int* critical, int n_models,
                      uint64_t seed, float* lat_us, float* svc_us, int cap, int* n_done, int* n_errors);

#ifdef __cplusplus
}
#endif
#endif
