/*
 * lig_oracle_models.c — CPU restatement of the step BEFORE Scheduler.Schedule (SURVEY.md 8f, row
 * f2).  TEST INFRASTRUCTURE ONLY (see lig_oracle.h).
 *
 * What it restates (paths relative to the reference repo, pkg/ext-proc/...):
 *   handlers/request.go:42-56     HandleRequestBody's resolve step: FetchModelData, the weighted
 *                                 target draw when TargetModels is non-empty, Critical
 *   backend/datastore.go:70-76    FetchModelData (map lookup; nil when absent)
 *   backend/datastore.go:78-98    RandomWeightedDraw: weights summed as int32, r.Int31n(weights),
 *                                 first target with randomVal < Weight, else randomVal -= Weight
 *   backend/datastore.go:100-105  IsCritical
 *
 * Parity pinning.  The reference draws from rand.NewSource(rand.Int63()) (seed == 0 in production,
 * request.go:48) — unseeded, not reproducible.  Its own test (backend/datastore_test.go:9-90) pins
 * three weight tables with seed 420, but reproducing Go's seeded additive-lagged-Fibonacci source
 * needs its 607-word rngCooked table, which is not in this image (no Go toolchain, no stdlib
 * source): DRAW PARITY IS UNPINNED.  As for the final pick (include/lig.h) the draw is therefore
 * DEFINED on an injected SplitMix64 source private to the request, with Go's published Int31n on
 * top of it:  state = seed ^ rand_key ^ LIGO_DRAW_DOMAIN.  What IS pinned: the loop semantics
 * (tests replay datastore_test.go's three tables and check that a draw value v selects exactly the
 * target the Go loop selects for randomVal = v, for every v in [0, sum)), and the distribution.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lig_oracle.h"

typedef struct {            /* v1alpha1.TargetModel */
  char* Name;
  int32_t Weight;
} TargetModel;

typedef struct {            /* v1alpha1.InferenceModel, the fields the path reads */
  char* Name;               /* Spec.ModelName: the key of the InferenceModels map */
  int Critical;             /* Spec.Criticality != nil && *Spec.Criticality == Critical */
  TargetModel* TargetModels;
  int n_targets;
  int present;              /* 0 = FetchModelData returns nil for this id */
} InferenceModel;

struct lig_oracle_models {
  InferenceModel* m;
  int n;
};

lig_oracle_models* lig_oracle_models_new(int n_models) {
  lig_oracle_models* s = (lig_oracle_models*)calloc(1, sizeof(*s));
  s->n = n_models;
  s->m = (InferenceModel*)calloc((size_t)(n_models > 0 ? n_models : 1), sizeof(InferenceModel));
  return s;
}

void lig_oracle_models_free(lig_oracle_models* s) {
  if (!s) return;
  for (int i = 0; i < s->n; ++i) {
    free(s->m[i].Name);
    for (int k = 0; k < s->m[i].n_targets; ++k) free(s->m[i].TargetModels[k].Name);
    free(s->m[i].TargetModels);
  }
  free(s->m);
  free(s);
}

int lig_oracle_models_set(lig_oracle_models* s, int i, const char* name, int critical,
                          const char* const* target_names, const int32_t* weights, int n_targets) {
  if (!s || i < 0 || i >= s->n || n_targets < 0) return -1;
  InferenceModel* m = &s->m[i];
  m->Name = strdup(name ? name : "");
  m->Critical = critical != 0;
  m->n_targets = n_targets;
  m->TargetModels = (TargetModel*)calloc((size_t)(n_targets > 0 ? n_targets : 1), sizeof(TargetModel));
  for (int k = 0; k < n_targets; ++k) {
    m->TargetModels[k].Name = strdup(target_names[k]);
    m->TargetModels[k].Weight = weights[k];
  }
  m->present = 1;
  return 0;
}

/* The selection loop of RandomWeightedDraw for a given randomVal.            datastore.go:91-97 */
int lig_oracle_weighted_select(const lig_oracle_models* s, int model, int32_t randomVal) {
  const InferenceModel* m = &s->m[model];
  for (int k = 0; k < m->n_targets; ++k) {
    if (randomVal < m->TargetModels[k].Weight) return k;
    randomVal -= m->TargetModels[k].Weight;
  }
  return -1;                                                                /* return "" */
}

/* RandomWeightedDraw(model, seed) on the injected source.                   datastore.go:78-98
 * Returns the index of the drawn target, -1 for "", -2 where Go would panic (Int31n(n <= 0)). */
int lig_oracle_random_weighted_draw(const lig_oracle_models* s, int model, uint64_t state) {
  const InferenceModel* m = &s->m[model];
  int32_t weights = 0;
  for (int k = 0; k < m->n_targets; ++k)
    weights = (int32_t)((uint32_t)weights + (uint32_t)m->TargetModels[k].Weight);   /* int32 wrap like Go */
  if (weights <= 0) return -2;
  const int32_t randomVal = lig_oracle_int31n(&state, weights);
  return lig_oracle_weighted_select(s, model, randomVal);
}

/* request.go:42-56.  Returns 0 and fills (*resolved, *critical, *target_idx) — target_idx is 255
 * when the model has no TargetModels (the request's own model name passes through) —, 3 when
 * FetchModelData finds nothing ("error finding a model object in InferenceModel"), 4 when the
 * draw returns "" ("error getting target model name"). */
int lig_oracle_resolve(const lig_oracle_models* s, int model, uint64_t seed, uint64_t rand_key,
                       const char** resolved, int* critical, int* target_idx) {
  if (!s || model < 0 || model >= s->n || !s->m[model].present) return 3;
  const InferenceModel* m = &s->m[model];
  *resolved = m->Name;
  *target_idx = 255;
  if (m->n_targets > 0) {
    const int k = lig_oracle_random_weighted_draw(s, model, seed ^ rand_key ^ LIGO_DRAW_DOMAIN);
    if (k < 0) return 4;
    *resolved = m->TargetModels[k].Name;
    *target_idx = k;
  }
  *critical = m->Critical;
  return 0;
}

/* R requests given as model ids; request i has rand_key = first_index + i.  Resolve, then
 * Scheduler.Schedule on the port (lig_oracle_schedule), one request at a time; nthreads > 1
 * partitions the requests statically over that many pthreads (one goroutine per request in the
 * reference, handlers/server.go:51). */
typedef struct {
  const lig_oracle_pool* pool;
  const lig_oracle_models* s;
  const uint32_t* model_ids;
  lig_oracle_mpick* out;
  int lo, hi;
  uint64_t seed, first_index;
} models_job;

static void* models_worker(void* arg) {
  models_job* j = (models_job*)arg;
  for (int i = j->lo; i < j->hi; ++i) {
    const uint64_t key = j->first_index + (uint64_t)i;
    const char* name = NULL;
    int critical = 0, target = 255;
    const int rs = j->model_ids[i] <= 0x7fffffffu
                       ? lig_oracle_resolve(j->s, (int)j->model_ids[i], j->seed, key, &name, &critical, &target) : 3;
    j->out[i].pod_idx = -1;
    j->out[i].target_idx = (uint8_t)target;
    if (rs != 0) {
      j->out[i].status = (uint8_t)rs;
      j->out[i].target_idx = 255;
      continue;
    }
    int32_t pod = -1;
    int n = 0;
    const int st = lig_oracle_schedule(j->pool, name, critical, j->seed, key, &pod, &n);
    j->out[i].status = (uint8_t)st;
    j->out[i].pod_idx = (int16_t)(st == LIGO_OK ? pod : -1);
  }
  return NULL;
}

int lig_oracle_schedule_models_batch(const lig_oracle_pool* pool, const lig_oracle_models* s,
                                     const uint32_t* model_ids, int R, uint64_t seed,
                                     uint64_t first_index, lig_oracle_mpick* out, int nthreads) {
  if (!pool || !s || R < 0 || (R > 0 && (!model_ids || !out))) return -1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > R) nthreads = R > 0 ? R : 1;
  models_job* jobs = (models_job*)calloc((size_t)nthreads, sizeof(models_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    models_job j = {pool, s, model_ids, out, (int)((int64_t)R * t / nthreads),
                    (int)((int64_t)R * (t + 1) / nthreads), seed, first_index};
    jobs[t] = j;
  }
  if (nthreads == 1) {
    models_worker(&jobs[0]);
  } else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, models_worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  free(jobs);
  free(th);
  return 0;
}
