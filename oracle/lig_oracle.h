/*
 * lig_oracle.h — CPU restatement of the reference scheduler hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library (liblig.so) does not
 * link, include or call anything in oracle/.
 *
 * What it restates (file:line relative to the reference repo, kubernetes-sigs/llm-instance-gateway
 * @ 8e96339; the Go toolchain is absent in this image, so the reference itself cannot run here):
 *   pkg/ext-proc/scheduling/scheduler.go:15-122   constants, defaultFilter tree, Scheduler.Schedule
 *   pkg/ext-proc/scheduling/filter.go:44-187      node semantics, range filters, predicates
 *   pkg/ext-proc/scheduling/types.go:4-11         LLMRequest
 *   pkg/ext-proc/backend/types.go:8-31            Pod, Metrics, PodMetrics
 * It keeps the reference's data structures and pass structure on purpose (slice of pointers to
 * pod structs, a string-keyed ActiveModels map per pod, a fresh slice per filter stage, the
 * snapshot slice materialised twice per Schedule call) so that it can double as the timed
 * "reference CPU path" (cpu_baseline.kind = "port").
 *
 * Parity pinning: the survivor sets are pinned against every golden vector of the reference's own
 * tests (tests/golden/go_filter_test_vectors.json, extracted from scheduling/filter_test.go and
 * test/hermetic_test.go by tests/golden/extract_go_vectors.py).  The final random pick is NOT
 * pinned by any reference test (rand.Intn on the auto-seeded global source, scheduler.go:120):
 * pick parity is "unpinned" and defined by include/lig.h instead (SplitMix64 source + Go's
 * published Int31n algorithm).
 */
#ifndef LIG_ORACLE_H_
#define LIG_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 16-byte / 8-byte records as include/lig.h (restated here so oracle/ has no dependency on
 * the product tree). */
typedef struct { int32_t adapter_id; uint32_t flags; uint64_t rand_key; } lig_oracle_req;
typedef struct { int32_t pod_idx; uint16_t status; uint16_t n_survivors; } lig_oracle_pick;

enum { LIGO_OK = 0, LIGO_DROP = 1, LIGO_EMPTY = 2, LIGO_ERROR = 3 };

typedef struct lig_oracle_pool lig_oracle_pool; /* a fake PodMetricsProvider: an ordered slice */

lig_oracle_pool* lig_oracle_pool_new(int n_pods);
void lig_oracle_pool_free(lig_oracle_pool*);
/* Define pod i.  active_models: n_active NUL-terminated names (keys of Metrics.ActiveModels). */
int lig_oracle_pool_set_pod(lig_oracle_pool*, int i, const char* name, const char* address,
                            int64_t waiting_queue_size, double kv_cache_usage_percent,
                            int64_t max_active_models, const char* const* active_models,
                            int n_active);
int lig_oracle_pool_size(const lig_oracle_pool*);

void lig_oracle_set_thresholds(double kv_cache_threshold, int64_t queue_threshold_critical,
                               int64_t queueing_threshold_lora);

/* defaultFilter.Filter(req, pool) — scheduler.go:26-31 / filter.go:44-73.
 * Writes the survivor pod indices (in slice order) to out_idx[0..*n_out); returns LIGO_OK,
 * LIGO_DROP (ResourceExhausted error), LIGO_EMPTY ([] with nil error) or LIGO_ERROR. */
int lig_oracle_filter(const lig_oracle_pool*, const char* resolved_target_model, int critical,
                      int32_t* out_idx, int* n_out);

/* Individual filter funcs, for the reference's TestFilterFunc cases (filter_test.go:217-409).
 * `which`: 0 leastQueuingFilterFunc, 1 leastKVCacheFilterFunc, 2 lowLoRACostPredicate,
 * 3 loRAAffinityPredicate, 4 canAcceptNewLoraPredicate, 5 lowQueueingPodPredicate,
 * 6 criticalRequestPredicate, 7 noQueueAndLessThanKVCacheThresholdPredicate(q_thr, kv_thr).
 * Returns 0 (nil error) or LIGO_ERROR ("no pods left"). */
int lig_oracle_filter_func(const lig_oracle_pool*, int which, const char* resolved_target_model,
                           int critical, int64_t q_thr, double kv_thr, int32_t* out_idx,
                           int* n_out);

/* A single-node tree whose filterFunc returns (nil, error) and has no successors
 * (filter_test.go:21-27): returns LIGO_ERROR with *n_out = 0. */
int lig_oracle_filter_error_leaf(const lig_oracle_pool*, int32_t* out_idx, int* n_out);

/* Scheduler.Schedule(req) with the pick defined by include/lig.h.  scheduler.go:113-122. */
int lig_oracle_schedule(const lig_oracle_pool*, const char* resolved_target_model, int critical,
                        uint64_t seed, uint64_t rand_key, int32_t* pod_idx, int* n_survivors);

/* Batch form over interned requests: adapter_names[a] is the model name of adapter id a
 * (0 <= a < n_adapters); ids outside that range schedule `unknown_model_name`.
 * masks (nullable): R x ceil(P/32) survivor words.  nthreads > 1 partitions the requests
 * statically over that many pthreads (the goroutine-per-request analogue). */
int lig_oracle_schedule_batch(const lig_oracle_pool*, const char* const* adapter_names,
                              int n_adapters, const char* unknown_model_name,
                              const lig_oracle_req* reqs, int R, uint64_t seed,
                              lig_oracle_pick* out, uint32_t* masks, int nthreads);

/* "Optimised CPU" fairness datapoint (lig_oracle_soa.c): the same per-request tree walk on the
 * GPU's data layout (columns + adapter-major bitmap + mask words), no allocation per request.
 * Column meanings as in include/lig.h; thresholds passed explicitly. */
int lig_oracle_soa_schedule_batch(int P, int A, const double* kv, const int32_t* q,
                                  const uint16_t* n_active, const uint16_t* max_active,
                                  const uint32_t* bitmap_adapter_major, double kv_cache_threshold,
                                  int64_t queue_threshold_critical, int64_t queueing_threshold_lora,
                                  const lig_oracle_req* reqs, int R, uint64_t seed,
                                  lig_oracle_pick* out, uint32_t* masks, int nthreads);

/* "Class-table CPU" fairness datapoint and fast whole-shard checker (lig_oracle_classtab.c): the
 * same algorithmic restructuring the GPU path uses, on the host — the tree is walked once per
 * request class (critical, adapter) per snapshot (2(A+1) walks, `nthreads` threads), a request is
 * then a table lookup + Int31n + one list read. */
typedef struct lig_oracle_classtab lig_oracle_classtab;
lig_oracle_classtab* lig_oracle_classtab_build(int P, int A, const double* kv, const int32_t* q,
                                               const uint16_t* n_active, const uint16_t* max_active,
                                               const uint32_t* bitmap_adapter_major,
                                               double kv_cache_threshold, int64_t queue_threshold_critical,
                                               int64_t queueing_threshold_lora, int nthreads);
void lig_oracle_classtab_free(lig_oracle_classtab*);
int lig_oracle_classtab_schedule_batch(const lig_oracle_classtab*, const lig_oracle_req* reqs, int R,
                                       uint64_t seed, lig_oracle_pick* out, int nthreads);
int lig_oracle_classtab_class(const lig_oracle_classtab*, int critical, int adapter_id, int* status,
                              int* n, uint16_t* list);

/* ---- the step before Schedule (lig_oracle_models.c): handlers/request.go:42-56,
 * backend/datastore.go:70-105.  A lig_oracle_models is the datastore's InferenceModels map keyed by
 * a dense model id.  Draw parity against Go's seeded source is UNPINNED (see the file header); the
 * draw is defined on the request's private SplitMix64 stream seed ^ rand_key ^ LIGO_DRAW_DOMAIN. */
#define LIGO_DRAW_DOMAIN 0xA0761D6478BD642Full
enum { LIGO_NO_MODEL = 3, LIGO_NO_TARGET = 4 };
typedef struct { int16_t pod_idx; uint8_t status; uint8_t target_idx; } lig_oracle_mpick;   /* 4 B */
typedef struct lig_oracle_models lig_oracle_models;
lig_oracle_models* lig_oracle_models_new(int n_models);
void lig_oracle_models_free(lig_oracle_models*);
int lig_oracle_models_set(lig_oracle_models*, int i, const char* name, int critical,
                          const char* const* target_names, const int32_t* weights, int n_targets);
int lig_oracle_weighted_select(const lig_oracle_models*, int model, int32_t randomVal);
int lig_oracle_random_weighted_draw(const lig_oracle_models*, int model, uint64_t state);
int lig_oracle_resolve(const lig_oracle_models*, int model, uint64_t seed, uint64_t rand_key,
                       const char** resolved, int* critical, int* target_idx);
int lig_oracle_schedule_models_batch(const lig_oracle_pool*, const lig_oracle_models*,
                                     const uint32_t* model_ids, int R, uint64_t seed,
                                     uint64_t first_index, lig_oracle_mpick* out, int nthreads);

/* The pick primitives, exposed for known-answer tests. */
uint64_t lig_oracle_splitmix64_next(uint64_t* state);
int32_t  lig_oracle_int31n(uint64_t* state, int32_t n);

int lig_oracle_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
