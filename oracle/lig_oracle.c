/*
 * lig_oracle.c — CPU restatement of the reference scheduler hot path.  TEST INFRASTRUCTURE ONLY
 * (see lig_oracle.h for the rules on who may load it and for the parity-pinning statement).
 *
 * Every function cites the reference lines it follows (paths relative to the reference repo
 * root, pkg/ext-proc/...).  Written from the behaviour of that code, in C, keeping its shape:
 * pods are a slice of pointers, ActiveModels is a string-keyed hash map, every stage appends
 * into a fresh slice, and the filter tree is a graph of nodes with the three successor fields.
 */
#define _GNU_SOURCE
#include "lig_oracle.h"

#include <float.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* A Go-like map[string]int (only key presence and len() are ever read on this path).          */

typedef struct {
  char** keys;    /* owned copies; NULL = empty slot */
  uint32_t cap;   /* power of two, >= 2 * len */
  uint32_t len;
} strmap;

static uint64_t str_hash(const char* s) { /* FNV-1a; Go uses a different (AES/wyhash) string
                                             hash, but any hash gives the same lookups */
  uint64_t h = 1469598103934665603ull;
  for (; *s; ++s) { h ^= (unsigned char)*s; h *= 1099511628211ull; }
  return h;
}

static void strmap_init(strmap* m, uint32_t n) {
  uint32_t cap = 4;
  while (cap < 2 * n + 2) cap <<= 1;
  m->keys = (char**)calloc(cap, sizeof(char*));
  m->cap = cap;
  m->len = 0;
}

static void strmap_put(strmap* m, const char* k) {
  uint32_t i = (uint32_t)str_hash(k) & (m->cap - 1);
  while (m->keys[i]) {
    if (strcmp(m->keys[i], k) == 0) return; /* Go map: duplicate key overwrites, len unchanged */
    i = (i + 1) & (m->cap - 1);
  }
  m->keys[i] = strdup(k);
  m->len++;
}

static int strmap_has(const strmap* m, const char* k) {
  if (!m->keys) return 0;
  uint32_t i = (uint32_t)str_hash(k) & (m->cap - 1);
  while (m->keys[i]) {
    if (strcmp(m->keys[i], k) == 0) return 1;
    i = (i + 1) & (m->cap - 1);
  }
  return 0;
}

static void strmap_free(strmap* m) {
  if (!m->keys) return;
  for (uint32_t i = 0; i < m->cap; ++i) free(m->keys[i]);
  free(m->keys);
  m->keys = NULL;
  m->cap = m->len = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* backend/types.go:8-31                                                                        */

typedef struct { char* Name; char* Address; } Pod;                         /* types.go:8-11  */

typedef struct {                                                           /* types.go:17-26 */
  strmap  ActiveModels;
  int64_t MaxActiveModels;
  int64_t RunningQueueSize;
  int64_t WaitingQueueSize;
  double  KVCacheUsagePercent;
  int64_t KvCacheMaxTokenCapacity;
} Metrics;

typedef struct { Pod pod; Metrics m; } PodMetrics;                         /* types.go:28-31 */

typedef struct {                                                           /* scheduling/types.go:4-11 */
  const char* Model;
  const char* ResolvedTargetModel;
  int Critical;
} LLMRequest;

struct lig_oracle_pool { PodMetrics* pods; int n; };

/* a Go slice of *PodMetrics */
typedef struct { PodMetrics** v; int len; int cap; } podslice;

static void slice_append(podslice* s, PodMetrics* p) { /* append(): amortised doubling */
  if (s->len == s->cap) {
    int ncap = s->cap ? 2 * s->cap : 1;
    s->v = (PodMetrics**)realloc(s->v, (size_t)ncap * sizeof(PodMetrics*));
    s->cap = ncap;
  }
  s->v[s->len++] = p;
}

/* Everything one Schedule call allocates, released together at the end (the Go GC's job). */
typedef struct { void* ptr[24]; int n; } garbage;
static void gc_track(garbage* g, void* p) { if (p && g->n < 24) g->ptr[g->n++] = p; }
static void gc_release(garbage* g) { for (int i = 0; i < g->n; ++i) free(g->ptr[i]); g->n = 0; }

/* error values */
typedef struct { int code; const char* msg; } lig_err;
static const lig_err ERR_NO_PODS_LEFT = {LIGO_ERROR, "no pods left"};                /* filter.go:89 */
static const lig_err ERR_RESOURCE_EXHAUSTED = {LIGO_DROP,                            /* scheduler.go:87 */
    "dropping request due to limited backend resources"};
static const lig_err ERR_FILTER_ERROR = {LIGO_ERROR, "filter error"};                /* filter_test.go:24 */

/* scheduler.go:15-24 — compile-time constants in the reference; variables here so the tests can
 * reproduce filter_test.go:306, which instantiates the predicate with (0, 0.8). */
static double  kvCacheThreshold       = 0.8;
static int64_t queueThresholdCritical = 5;
static int64_t queueingThresholdLoRA  = 50;

void lig_oracle_set_thresholds(double kv, int64_t qcrit, int64_t qlora) {
  kvCacheThreshold = kv;
  queueThresholdCritical = qcrit;
  queueingThresholdLoRA = qlora;
}

/* ------------------------------------------------------------------------------------------ */
/* filter.go:12-42 — the node type                                                              */

struct filter;
typedef podslice (*filterFunc)(const struct filter* self, const LLMRequest* req, podslice pods,
                               const lig_err** err, garbage* g);
typedef int (*podPredicate)(const struct filter* self, const LLMRequest* req,
                            const PodMetrics* pod);

typedef struct filter {
  const char* name;
  filterFunc filter;
  const struct filter* nextOnSuccess;
  const struct filter* nextOnFailure;
  const struct filter* nextOnSuccessOrFailure;
  /* closure state of toFilterFunc(pp) and of noQueueAndLessThanKVCacheThresholdPredicate(q, kv) */
  podPredicate pp;
  int use_closure_thresholds;
  int64_t queueThreshold;
  double kvThreshold;
} filter;

/* filter.go:44-73 — (f *filter) Filter */
static podslice filter_Filter(const filter* f, const LLMRequest* req, podslice pods,
                              const lig_err** err, garbage* g) {
  const lig_err* e = NULL;
  podslice filtered = f->filter(f, req, pods, &e, g);               /* filter.go:47 */

  const filter* next = f->nextOnSuccessOrFailure;                   /* filter.go:49 */
  if (e == NULL && filtered.len > 0) {                              /* filter.go:50 */
    if (f->nextOnSuccess == NULL && f->nextOnSuccessOrFailure == NULL) {
      *err = e;                                                     /* filter.go:51-54 */
      return filtered;
    }
    if (f->nextOnSuccess != NULL) next = f->nextOnSuccess;          /* filter.go:55-57 */
    return filter_Filter(next, req, filtered, err, g);              /* filter.go:60: filtered set */
  } else {
    if (f->nextOnFailure == NULL && f->nextOnSuccessOrFailure == NULL) {
      *err = e;                                                     /* filter.go:62-65 */
      return filtered;
    }
    if (f->nextOnFailure != NULL) next = f->nextOnFailure;          /* filter.go:66-68 */
    return filter_Filter(next, req, pods, err, g);                  /* filter.go:71: this node's input */
  }
}

/* filter.go:79-93 — toFilterFunc */
static podslice predicate_filter(const filter* self, const LLMRequest* req, podslice pods,
                                 const lig_err** err, garbage* g) {
  podslice filtered = {NULL, 0, 0};                                 /* []*PodMetrics{} */
  for (int i = 0; i < pods.len; ++i) {
    if (self->pp(self, req, pods.v[i])) slice_append(&filtered, pods.v[i]);
  }
  gc_track(g, filtered.v);
  if (filtered.len == 0) {                                          /* filter.go:88-90 */
    *err = &ERR_NO_PODS_LEFT;
    podslice nil = {NULL, 0, 0};
    return nil;
  }
  *err = NULL;
  return filtered;
}

/* filter.go:102-122 — leastQueuingFilterFunc.  Go `int` is 64-bit on every supported server
 * platform; `/` is truncated division. */
static podslice leastQueuingFilterFunc(const filter* self, const LLMRequest* req, podslice pods,
                                       const lig_err** err, garbage* g) {
  (void)self; (void)req;
  int64_t min = INT64_MAX;                                          /* math.MaxInt */
  int64_t max = 0;
  podslice filtered = {NULL, 0, 0};
  for (int i = 0; i < pods.len; ++i) {                              /* filter.go:107-114 */
    int64_t q = pods.v[i]->m.WaitingQueueSize;
    if (q <= min) min = q;
    if (q >= max) max = q;
  }
  for (int i = 0; i < pods.len; ++i) {                              /* filter.go:116-120 */
    int64_t q = pods.v[i]->m.WaitingQueueSize;
    /* Go wraps on signed overflow; do the arithmetic in uint64 and cast back to match. */
    int64_t thr = (int64_t)((uint64_t)min +
                            (uint64_t)((int64_t)((uint64_t)max - (uint64_t)min) / (int64_t)pods.len));
    if (q >= min && q <= thr) slice_append(&filtered, pods.v[i]);
  }
  gc_track(g, filtered.v);
  *err = NULL;
  return filtered;
}

/* filter.go:134-154 — leastKVCacheFilterFunc.  Three separately rounded binary64 operations
 * (compile with -ffp-contract=off; see oracle/Makefile). */
static podslice leastKVCacheFilterFunc(const filter* self, const LLMRequest* req, podslice pods,
                                       const lig_err** err, garbage* g) {
  (void)self; (void)req;
  double min = DBL_MAX;                                             /* math.MaxFloat64 */
  double max = 0;
  podslice filtered = {NULL, 0, 0};
  for (int i = 0; i < pods.len; ++i) {                              /* filter.go:139-146 */
    double kv = pods.v[i]->m.KVCacheUsagePercent;
    if (kv <= min) min = kv;
    if (kv >= max) max = kv;
  }
  for (int i = 0; i < pods.len; ++i) {                              /* filter.go:148-152 */
    double kv = pods.v[i]->m.KVCacheUsagePercent;
    volatile double range = max - min;
    volatile double step = range / (double)pods.len;
    volatile double thr = min + step;
    if (kv >= min && kv <= thr) slice_append(&filtered, pods.v[i]);
  }
  gc_track(g, filtered.v);
  *err = NULL;
  return filtered;
}

/* filter.go:124-126 */
static int lowQueueingPodPredicate(const filter* s, const LLMRequest* r, const PodMetrics* pod) {
  (void)s; (void)r;
  return pod->m.WaitingQueueSize < queueingThresholdLoRA;
}
/* filter.go:163-166 */
static int lowLoRACostPredicate(const filter* s, const LLMRequest* req, const PodMetrics* pod) {
  (void)s;
  int ok = strmap_has(&pod->m.ActiveModels, req->ResolvedTargetModel);
  return ok || (int64_t)pod->m.ActiveModels.len < pod->m.MaxActiveModels;
}
/* filter.go:169-172 */
static int loRAAffinityPredicate(const filter* s, const LLMRequest* req, const PodMetrics* pod) {
  (void)s;
  return strmap_has(&pod->m.ActiveModels, req->ResolvedTargetModel);
}
/* filter.go:175-177 */
static int canAcceptNewLoraPredicate(const filter* s, const LLMRequest* r, const PodMetrics* pod) {
  (void)s; (void)r;
  return (int64_t)pod->m.ActiveModels.len < pod->m.MaxActiveModels;
}
/* filter.go:179-181 */
static int criticalRequestPredicate(const filter* s, const LLMRequest* req, const PodMetrics* pod) {
  (void)s; (void)pod;
  return req->Critical;
}
/* filter.go:183-187 — a closure over (queueThreshold, kvCacheThreshold) */
static int noQueueAndLessThanKVCacheThresholdPredicate(const filter* s, const LLMRequest* r,
                                                       const PodMetrics* pod) {
  (void)r;
  int64_t qt = s->use_closure_thresholds ? s->queueThreshold : queueThresholdCritical;
  double kt = s->use_closure_thresholds ? s->kvThreshold : kvCacheThreshold;
  return pod->m.WaitingQueueSize <= qt && pod->m.KVCacheUsagePercent <= kt;
}

/* scheduler.go:83-89 — the "drop request" leaf */
static podslice dropRequestFilterFunc(const filter* self, const LLMRequest* req, podslice pods,
                                      const lig_err** err, garbage* g) {
  (void)self; (void)req; (void)pods; (void)g;
  podslice empty = {NULL, 0, 0};                                    /* []*backend.PodMetrics{} */
  *err = &ERR_RESOURCE_EXHAUSTED;
  return empty;
}

/* filter_test.go:21-27 — a filterFunc that returns (nil, error) */
static podslice alwaysErrorFilterFunc(const filter* self, const LLMRequest* req, podslice pods,
                                      const lig_err** err, garbage* g) {
  (void)self; (void)req; (void)pods; (void)g;
  podslice nil = {NULL, 0, 0};
  *err = &ERR_FILTER_ERROR;
  return nil;
}

/* ------------------------------------------------------------------------------------------ */
/* scheduler.go:26-91 — the tree (same node names)                                              */


static const filter leastKV_leaf_a = {.name = "least KV cache percent", .filter = leastKVCacheFilterFunc};
static const filter lowCostLoRA = {                                 /* scheduler.go:38-45 */
    .name = "low cost LoRA", .filter = predicate_filter, .pp = lowLoRACostPredicate,
    .nextOnSuccessOrFailure = &leastKV_leaf_a};
static const filter queueLoRAAndKVCacheFilter = {                   /* scheduler.go:35-46 */
    .name = "least queuing", .filter = leastQueuingFilterFunc,
    .nextOnSuccessOrFailure = &lowCostLoRA};

static const filter leastKV_leaf_b = {.name = "least KV cache percent", .filter = leastKVCacheFilterFunc};
static const filter queueAndKVCacheFilter = {                       /* scheduler.go:49-56 */
    .name = "least queuing", .filter = leastQueuingFilterFunc,
    .nextOnSuccessOrFailure = &leastKV_leaf_b};

static const filter canAcceptLoRA = {                               /* scheduler.go:65-69 */
    .name = "can accept LoRA Adapter", .filter = predicate_filter, .pp = canAcceptNewLoraPredicate,
    .nextOnSuccessOrFailure = &queueAndKVCacheFilter};
static const filter affinityLoRA = {                                /* scheduler.go:61-70 */
    .name = "affinity LoRA", .filter = predicate_filter, .pp = loRAAffinityPredicate,
    .nextOnSuccess = &queueAndKVCacheFilter, .nextOnFailure = &canAcceptLoRA};
static const filter lowLatencyFilter = {                            /* scheduler.go:58-72 */
    .name = "low queueing filter", .filter = predicate_filter, .pp = lowQueueingPodPredicate,
    .nextOnSuccess = &affinityLoRA, .nextOnFailure = &queueLoRAAndKVCacheFilter};

static const filter dropRequest = {.name = "drop request", .filter = dropRequestFilterFunc};
static const filter sheddableRequestFilter = {                      /* scheduler.go:74-90 */
    .name = "has capacity for sheddable requests", .filter = predicate_filter,
    .pp = noQueueAndLessThanKVCacheThresholdPredicate,
    .nextOnSuccess = &queueLoRAAndKVCacheFilter, .nextOnFailure = &dropRequest};

static const filter defaultFilter = {                               /* scheduler.go:26-31 */
    .name = "critical request", .filter = predicate_filter, .pp = criticalRequestPredicate,
    .nextOnSuccess = &lowLatencyFilter, .nextOnFailure = &sheddableRequestFilter};

/* ------------------------------------------------------------------------------------------ */
/* pool = a fake PodMetricsProvider holding an ordered slice (scheduler.go:108-110)             */

lig_oracle_pool* lig_oracle_pool_new(int n) {
  if (n < 0) return NULL;
  lig_oracle_pool* p = (lig_oracle_pool*)calloc(1, sizeof(*p));
  p->n = n;
  p->pods = (PodMetrics*)calloc((size_t)(n > 0 ? n : 1), sizeof(PodMetrics));
  return p;
}

void lig_oracle_pool_free(lig_oracle_pool* p) {
  if (!p) return;
  for (int i = 0; i < p->n; ++i) {
    free(p->pods[i].pod.Name);
    free(p->pods[i].pod.Address);
    strmap_free(&p->pods[i].m.ActiveModels);
  }
  free(p->pods);
  free(p);
}

int lig_oracle_pool_size(const lig_oracle_pool* p) { return p ? p->n : -1; }

int lig_oracle_pool_set_pod(lig_oracle_pool* p, int i, const char* name, const char* address,
                            int64_t q, double kv, int64_t max_active,
                            const char* const* active, int n_active) {
  if (!p || i < 0 || i >= p->n || n_active < 0) return -1;
  PodMetrics* pm = &p->pods[i];
  free(pm->pod.Name);
  free(pm->pod.Address);
  strmap_free(&pm->m.ActiveModels);
  pm->pod.Name = strdup(name ? name : "");
  pm->pod.Address = strdup(address ? address : "");
  pm->m.WaitingQueueSize = q;
  pm->m.KVCacheUsagePercent = kv;
  pm->m.MaxActiveModels = max_active;
  strmap_init(&pm->m.ActiveModels, (uint32_t)n_active);
  for (int k = 0; k < n_active; ++k) strmap_put(&pm->m.ActiveModels, active[k]);
  return 0;
}

/* backend/provider.go:38-46 — AllPodMetrics(): a fresh slice grown by append. */
static podslice AllPodMetrics(const lig_oracle_pool* p, garbage* g) {
  podslice res = {NULL, 0, 0};
  for (int i = 0; i < p->n; ++i) slice_append(&res, &p->pods[i]);
  gc_track(g, res.v);
  return res;
}

static void emit_indices(const lig_oracle_pool* pool, podslice s, int32_t* out_idx, int* n_out) {
  for (int i = 0; i < s.len; ++i) out_idx[i] = (int32_t)(s.v[i] - pool->pods);
  *n_out = s.len;
}

static int classify(podslice s, const lig_err* e) {
  if (e) return e->code;          /* LIGO_DROP for ResourceExhausted, LIGO_ERROR otherwise */
  return s.len == 0 ? LIGO_EMPTY : LIGO_OK;
}

int lig_oracle_filter(const lig_oracle_pool* pool, const char* model, int critical,
                      int32_t* out_idx, int* n_out) {
  garbage g = {{0}, 0};
  LLMRequest req = {model, model, critical};
  const lig_err* e = NULL;
  podslice in = AllPodMetrics(pool, &g);
  podslice s = filter_Filter(&defaultFilter, &req, in, &e, &g);
  emit_indices(pool, s, out_idx, n_out);
  int rc = classify(s, e);
  gc_release(&g);
  return rc;
}

int lig_oracle_filter_func(const lig_oracle_pool* pool, int which, const char* model, int critical,
                           int64_t q_thr, double kv_thr, int32_t* out_idx, int* n_out) {
  garbage g = {{0}, 0};
  LLMRequest req = {model ? model : "", model ? model : "", critical};
  filter node;
  memset(&node, 0, sizeof(node));
  node.name = "single";
  switch (which) {
    case 0: node.filter = leastQueuingFilterFunc; break;
    case 1: node.filter = leastKVCacheFilterFunc; break;
    case 2: node.filter = predicate_filter; node.pp = lowLoRACostPredicate; break;
    case 3: node.filter = predicate_filter; node.pp = loRAAffinityPredicate; break;
    case 4: node.filter = predicate_filter; node.pp = canAcceptNewLoraPredicate; break;
    case 5: node.filter = predicate_filter; node.pp = lowQueueingPodPredicate; break;
    case 6: node.filter = predicate_filter; node.pp = criticalRequestPredicate; break;
    case 7:
      node.filter = predicate_filter;
      node.pp = noQueueAndLessThanKVCacheThresholdPredicate;
      node.use_closure_thresholds = 1;
      node.queueThreshold = q_thr;
      node.kvThreshold = kv_thr;
      break;
    default: return -1;
  }
  const lig_err* e = NULL;
  podslice in = AllPodMetrics(pool, &g);
  podslice s = node.filter(&node, &req, in, &e, &g);   /* test.f(test.req, test.input) */
  emit_indices(pool, s, out_idx, n_out);
  int rc = e ? LIGO_ERROR : 0;
  gc_release(&g);
  return rc;
}

int lig_oracle_filter_error_leaf(const lig_oracle_pool* pool, int32_t* out_idx, int* n_out) {
  garbage g = {{0}, 0};
  LLMRequest req = {"", "", 0};
  filter node;
  memset(&node, 0, sizeof(node));
  node.filter = alwaysErrorFilterFunc;
  const lig_err* e = NULL;
  podslice in = AllPodMetrics(pool, &g);
  podslice s = filter_Filter(&node, &req, in, &e, &g);
  emit_indices(pool, s, out_idx, n_out);
  int rc = classify(s, e);
  gc_release(&g);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* the pick: include/lig.h's definition of rand.New(src).Intn(n)                                */

uint64_t lig_oracle_splitmix64_next(uint64_t* state) {
  uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static int32_t src_Int31(uint64_t* st) {          /* Int31() = int32(Int63() >> 32) */
  uint64_t int63 = lig_oracle_splitmix64_next(st) >> 1;
  return (int32_t)(int63 >> 32);
}

int32_t lig_oracle_int31n(uint64_t* st, int32_t n) { /* math/rand (Go 1.22) Rand.Int31n */
  if ((n & (n - 1)) == 0) return src_Int31(st) & (n - 1);
  int32_t max = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
  int32_t v = src_Int31(st);
  while (v > max) v = src_Int31(st);
  return v % n;
}

/* scheduler.go:113-122 — Scheduler.Schedule */
static int Schedule(const lig_oracle_pool* pool, const LLMRequest* req, uint64_t seed,
                    uint64_t rand_key, int32_t* pod_idx, int* n_survivors, uint32_t* mask_row,
                    int W) {
  garbage g = {{0}, 0};
  podslice logged = AllPodMetrics(pool, &g);      /* scheduler.go:114: klog argument, evaluated eagerly */
  (void)logged;
  const lig_err* e = NULL;
  podslice pods = filter_Filter(&defaultFilter, req, AllPodMetrics(pool, &g), &e, &g); /* :115 */
  int rc = classify(pods, e);
  *n_survivors = pods.len;
  *pod_idx = -1;
  if (mask_row) {
    memset(mask_row, 0, (size_t)W * sizeof(uint32_t));
    for (int i = 0; i < pods.len; ++i) {
      int p = (int)(pods.v[i] - pool->pods);
      mask_row[p >> 5] |= 1u << (p & 31);
    }
  }
  if (!(e != NULL || pods.len == 0)) {            /* scheduler.go:116 */
    uint64_t st = seed ^ rand_key;
    int32_t i = lig_oracle_int31n(&st, (int32_t)pods.len);   /* scheduler.go:120 rand.Intn */
    *pod_idx = (int32_t)(pods.v[i] - pool->pods); /* scheduler.go:121 pods[i].Pod */
  } else if (rc == LIGO_DROP || rc == LIGO_ERROR) {
    *n_survivors = 0;
  }
  gc_release(&g);
  return rc;
}

int lig_oracle_schedule(const lig_oracle_pool* pool, const char* model, int critical,
                        uint64_t seed, uint64_t rand_key, int32_t* pod_idx, int* n_survivors) {
  LLMRequest req = {model, model, critical};
  return Schedule(pool, &req, seed, rand_key, pod_idx, n_survivors, NULL, 0);
}

typedef struct {
  const lig_oracle_pool* pool;
  const char* const* names;
  int n_adapters;
  const char* unknown;
  const lig_oracle_req* reqs;
  int lo, hi;
  uint64_t seed;
  lig_oracle_pick* out;
  uint32_t* masks;
  int W;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  for (int i = j->lo; i < j->hi; ++i) {
    const lig_oracle_req* r = &j->reqs[i];
    const char* model = (r->adapter_id >= 0 && r->adapter_id < j->n_adapters)
                            ? j->names[r->adapter_id] : j->unknown;
    LLMRequest req = {model, model, (int)(r->flags & 1u)};
    int32_t pod = -1;
    int n = 0;
    int rc = Schedule(j->pool, &req, j->seed, r->rand_key, &pod, &n,
                      j->masks ? j->masks + (size_t)i * j->W : NULL, j->W);
    j->out[i].pod_idx = pod;
    j->out[i].status = (uint16_t)rc;
    j->out[i].n_survivors = (uint16_t)n;
  }
  return NULL;
}

int lig_oracle_schedule_batch(const lig_oracle_pool* pool, const char* const* names,
                              int n_adapters, const char* unknown, const lig_oracle_req* reqs,
                              int R, uint64_t seed, lig_oracle_pick* out, uint32_t* masks,
                              int nthreads) {
  if (!pool || !reqs || !out || R < 0) return -1;
  int W = (pool->n + 31) / 32;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > R) nthreads = R > 0 ? R : 1;
  batch_job* jobs = (batch_job*)calloc((size_t)nthreads, sizeof(batch_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    batch_job j = {pool, names, n_adapters, unknown ? unknown : "", reqs,
                   (int)((int64_t)R * t / nthreads), (int)((int64_t)R * (t + 1) / nthreads),
                   seed, out, masks, W};
    jobs[t] = j;
  }
  if (nthreads == 1) {
    batch_worker(&jobs[0]);
  } else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  free(jobs);
  free(th);
  return 0;
}

int lig_oracle_hardware_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
