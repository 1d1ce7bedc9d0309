/*
 * lig_oracle_classtab.c — "class-table CPU" fairness datapoint and fast whole-shard checker.
 * TEST INFRASTRUCTURE ONLY, like everything under oracle/: never linked, imported or called by the
 * product.
 *
 * The survivor set of Scheduler.Schedule (pkg/ext-proc/scheduling/scheduler.go:113-122) depends on
 * the request only through (Critical, ResolvedTargetModel): filter.go:163-181 read nothing else of
 * the LLMRequest.  So a CPU can do what the GPU path does: walk the defaultFilter tree
 * (scheduler.go:26-91, filter.go:44-187) ONCE per request class per snapshot — 2(A+1) walks,
 * done here with the SoA walker of lig_oracle_soa.c, all host threads — and then answer each
 * request with a table lookup + Go's Int31n + one list read.  bench.py reports this arm as
 * cpu_baseline.class_table_cpu so that the GPU/CPU ratios can be read against a CPU
 * implementation with the SAME algorithmic restructuring; tests/test_oracle_cross.py checks it
 * against the structure-preserving port (lig_oracle.c) request by request.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lig_oracle.h"
#include "lig_oracle_soa_internal.h"

struct lig_oracle_classtab {
  int P, A, n_classes;
  int32_t* n;        /* survivors per class                               */
  uint8_t* status;   /* LIGO_OK / LIGO_DROP / LIGO_EMPTY per class         */
  uint16_t* lists;   /* n_classes rows of P entries: survivors, ascending  */
};

typedef struct {
  const soa_view* v;
  lig_oracle_classtab* t;
  int lo, hi;
} build_job;

static void* build_worker(void* arg) {
  build_job* j = (build_job*)arg;
  const soa_view* v = j->v;
  lig_oracle_classtab* t = j->t;
  const int W = v->W64 > 0 ? v->W64 : 1;
  uint64_t* x = (uint64_t*)calloc((size_t)W * 2, sizeof(uint64_t));
  uint64_t* tmp = x + W;
  for (int c = j->lo; c < j->hi; ++c) {
    const int critical = c >= v->A + 1;
    const int a = critical ? c - (v->A + 1) : c;     /* a == A: adapter active nowhere */
    int n = 0;
    const int st = ligo_soa_schedule_one(v, a, critical, x, tmp, &n);
    t->status[c] = (uint8_t)st;
    t->n[c] = (st == LIGO_OK) ? n : 0;
    if (st == LIGO_OK) {
      uint16_t* row = t->lists + (size_t)c * (size_t)(v->P > 0 ? v->P : 1);
      int k = 0;
      for (int w = 0; w < v->W64; ++w) {
        uint64_t bits = x[w];
        while (bits) {
          row[k++] = (uint16_t)(w * 64 + __builtin_ctzll(bits));
          bits &= bits - 1;
        }
      }
    }
  }
  free(x);
  return NULL;
}

lig_oracle_classtab* lig_oracle_classtab_build(int P, int A, const double* kv, const int32_t* q,
                                               const uint16_t* n_active, const uint16_t* max_active,
                                               const uint32_t* bitmap, double kv_thr, int64_t q_crit,
                                               int64_t q_lora, int nthreads) {
  if (P < 0 || A < 0 || P > 65535) return NULL;
  lig_oracle_classtab* t = (lig_oracle_classtab*)calloc(1, sizeof(*t));
  t->P = P;
  t->A = A;
  t->n_classes = 2 * (A + 1);
  t->n = (int32_t*)calloc((size_t)t->n_classes, sizeof(int32_t));
  t->status = (uint8_t*)calloc((size_t)t->n_classes, 1);
  t->lists = (uint16_t*)calloc((size_t)t->n_classes * (size_t)(P > 0 ? P : 1), sizeof(uint16_t));
  soa_view v;
  ligo_soa_view_init(&v, P, A, kv, q, n_active, max_active, bitmap, kv_thr, q_crit, q_lora);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > t->n_classes) nthreads = t->n_classes;
  build_job* jobs = (build_job*)calloc((size_t)nthreads, sizeof(build_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int i = 0; i < nthreads; ++i) {
    build_job j = {&v, t, (int)((int64_t)t->n_classes * i / nthreads),
                   (int)((int64_t)t->n_classes * (i + 1) / nthreads)};
    jobs[i] = j;
  }
  if (nthreads == 1) {
    build_worker(&jobs[0]);
  } else {
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, build_worker, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
  }
  free(jobs);
  free(th);
  ligo_soa_view_free(&v);
  return t;
}

void lig_oracle_classtab_free(lig_oracle_classtab* t) {
  if (!t) return;
  free(t->n);
  free(t->status);
  free(t->lists);
  free(t);
}

typedef struct {
  const lig_oracle_classtab* t;
  const lig_oracle_req* reqs;
  lig_oracle_pick* out;
  int lo, hi;
  uint64_t seed;
} pick_job;

static void* pick_worker(void* arg) {
  pick_job* j = (pick_job*)arg;
  const lig_oracle_classtab* t = j->t;
  const uint32_t A = (uint32_t)t->A;
  const size_t stride = (size_t)(t->P > 0 ? t->P : 1);
  for (int i = j->lo; i < j->hi; ++i) {
    const lig_oracle_req r = j->reqs[i];
    const uint32_t a = (uint32_t)r.adapter_id < A ? (uint32_t)r.adapter_id : A;   /* outside [0, A): active nowhere */
    const uint32_t c = (r.flags & 1u) * (A + 1u) + a;
    const int32_t n = t->n[c];
    int32_t pod = -1;
    if (n > 0) {
      uint64_t state = j->seed ^ r.rand_key;
      pod = t->lists[(size_t)c * stride + (size_t)lig_oracle_int31n(&state, n)];   /* scheduler.go:120-121 */
    }
    j->out[i].pod_idx = pod;
    j->out[i].status = t->status[c];
    j->out[i].n_survivors = (uint16_t)n;
  }
  return NULL;
}

int lig_oracle_classtab_schedule_batch(const lig_oracle_classtab* t, const lig_oracle_req* reqs, int R,
                                       uint64_t seed, lig_oracle_pick* out, int nthreads) {
  if (!t || R < 0 || (R > 0 && (!reqs || !out))) return -1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > R) nthreads = R > 0 ? R : 1;
  pick_job* jobs = (pick_job*)calloc((size_t)nthreads, sizeof(pick_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int i = 0; i < nthreads; ++i) {
    pick_job j = {t, reqs, out, (int)((int64_t)R * i / nthreads), (int)((int64_t)R * (i + 1) / nthreads), seed};
    jobs[i] = j;
  }
  if (nthreads == 1) {
    pick_worker(&jobs[0]);
  } else {
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, pick_worker, &jobs[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
  }
  free(jobs);
  free(th);
  return 0;
}

int lig_oracle_classtab_class(const lig_oracle_classtab* t, int critical, int adapter_id, int* status,
                              int* n, uint16_t* list) {
  if (!t || !status || !n) return -1;
  const int a = (adapter_id >= 0 && adapter_id < t->A) ? adapter_id : t->A;
  const int c = (critical ? 1 : 0) * (t->A + 1) + a;
  *status = t->status[c];
  *n = t->n[c];
  if (list && t->n[c] > 0)
    memcpy(list, t->lists + (size_t)c * (size_t)(t->P > 0 ? t->P : 1), (size_t)t->n[c] * sizeof(uint16_t));
  return 0;
}
