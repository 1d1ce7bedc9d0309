/*
 * lig_oracle_soa.c — "optimised CPU" fairness datapoint.  TEST INFRASTRUCTURE
 * ONLY, like everything under oracle/: never linked, imported or called by the product.
 *
 * Same algorithm as the reference (one walk of the defaultFilter tree per request —
 * pkg/ext-proc/scheduling/scheduler.go:26-91, filter.go:44-187 — then Int31n), but on the data
 * layout the GPU uses instead of the reference's: pod metrics as columns, ActiveModels as an
 * adapter-major bitmap, candidate sets as 64-bit mask words, no allocation per request.  It is a
 * third, independently written implementation of the tree; tests/test_oracle_cross.py checks it
 * against lig_oracle.c, and bench.py times it beside the structure-preserving port so that the
 * GPU numbers can be read against a CPU implementation that is not handicapped by pointer
 * chasing and string hashing.
 */
#define _GNU_SOURCE
#include <float.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lig_oracle.h"
#include "lig_oracle_soa_internal.h"


static inline int popc64(uint64_t x) { return __builtin_popcountll(x); }

static int mask_count(const uint64_t* m, int W) {
  int n = 0;
  for (int w = 0; w < W; ++w) n += popc64(m[w]);
  return n;
}

/* leastQueuingFilterFunc on a mask, in place; returns the new count.       filter.go:102-122 */
static int least_queuing(const soa_view* v, uint64_t* x, int n) {
  if (n == 0) return 0;
  int64_t mn = INT64_MAX, mx = 0;
  for (int w = 0; w < v->W64; ++w) {
    uint64_t bits = x[w];
    while (bits) {
      int p = w * 64 + __builtin_ctzll(bits);
      bits &= bits - 1;
      int64_t qq = v->q[p];
      if (qq <= mn) mn = qq;
      if (qq >= mx) mx = qq;
    }
  }
  int64_t thr = mn + (mx - mn) / n;   /* int32 inputs: no overflow; C division truncates like Go */
  int cnt = 0;
  for (int w = 0; w < v->W64; ++w) {
    uint64_t bits = x[w], keep = 0;
    while (bits) {
      int b = __builtin_ctzll(bits);
      bits &= bits - 1;
      int64_t qq = v->q[w * 64 + b];
      if (qq >= mn && qq <= thr) keep |= 1ull << b;
    }
    x[w] = keep;
    cnt += popc64(keep);
  }
  return cnt;
}

/* leastKVCacheFilterFunc on a mask, in place.                               filter.go:134-154 */
static int least_kv(const soa_view* v, uint64_t* x, int n) {
  if (n == 0) return 0;
  double mn = DBL_MAX, mx = 0.0;
  for (int w = 0; w < v->W64; ++w) {
    uint64_t bits = x[w];
    while (bits) {
      int p = w * 64 + __builtin_ctzll(bits);
      bits &= bits - 1;
      double k = v->kv[p];
      if (k <= mn) mn = k;
      if (k >= mx) mx = k;
    }
  }
  volatile double range = mx - mn;
  volatile double step = range / (double)n;
  volatile double thr = mn + step;
  int cnt = 0;
  for (int w = 0; w < v->W64; ++w) {
    uint64_t bits = x[w], keep = 0;
    while (bits) {
      int b = __builtin_ctzll(bits);
      bits &= bits - 1;
      double k = v->kv[w * 64 + b];
      if (k >= mn && k <= thr) keep |= 1ull << b;
    }
    x[w] = keep;
    cnt += popc64(keep);
  }
  return cnt;
}

/* adapter row as 64-bit words (the bitmap is stored in 32-bit words) */
static inline uint64_t row64(const soa_view* v, const uint32_t* row, int w) {
  uint64_t lo = (2 * w < v->W32) ? row[2 * w] : 0;
  uint64_t hi = (2 * w + 1 < v->W32) ? row[2 * w + 1] : 0;
  return lo | (hi << 32);
}

/* queueLoRAAndKVCacheFilter on x (count n): least queuing -> low cost LoRA -> least KV.
 *                                                                           scheduler.go:35-46 */
static int queue_lora_kv(const soa_view* v, const uint32_t* row, uint64_t* x, uint64_t* t, int n) {
  n = least_queuing(v, x, n);
  int nz = 0;
  for (int w = 0; w < v->W64; ++w) {
    uint64_t h = row ? row64(v, row, w) : 0;
    t[w] = x[w] & (h | v->m_room[w]);                                   /* filter.go:163-166 */
    nz += popc64(t[w]);
  }
  if (nz > 0) {
    memcpy(x, t, (size_t)v->W64 * sizeof(uint64_t));
    n = nz;
  }
  return least_kv(v, x, n);
}

/* One request; x receives the survivor mask.  Returns the status. */
int ligo_soa_schedule_one(const soa_view* v, int adapter, int critical, uint64_t* x, uint64_t* t,
                          int* n_out) {
  const uint32_t* row = (adapter >= 0 && adapter < v->A) ? v->bitmap + (size_t)adapter * v->W32 : NULL;
  const size_t bytes = (size_t)v->W64 * sizeof(uint64_t);
  int n;
  if (critical && v->P > 0) {                                            /* scheduler.go:26-31 */
    if (v->n_low > 0) {                                                  /* scheduler.go:58-60 */
      int nb = 0;
      for (int w = 0; w < v->W64; ++w) {                                 /* affinity  :61-64 */
        t[w] = row ? (v->m_low[w] & row64(v, row, w)) : 0;
        nb += popc64(t[w]);
      }
      if (nb > 0) {
        memcpy(x, t, bytes);
        n = nb;
      } else {
        int nc = 0;
        for (int w = 0; w < v->W64; ++w) {                               /* can accept :65-69 */
          t[w] = v->m_low[w] & v->m_room[w];
          nc += popc64(t[w]);
        }
        if (nc > 0) { memcpy(x, t, bytes); n = nc; }
        else { memcpy(x, v->m_low, bytes); n = v->n_low; }               /* failure forwards the input */
      }
      n = least_queuing(v, x, n);                                        /* scheduler.go:49-56 */
      n = least_kv(v, x, n);
      *n_out = n;
      return n ? LIGO_OK : LIGO_EMPTY;
    }
    memcpy(x, v->m_all, bytes);                                          /* scheduler.go:71 */
    n = v->P;
  } else {
    if (v->n_shed == 0) {                                                /* scheduler.go:83-89 */
      memset(x, 0, bytes);
      *n_out = 0;
      return LIGO_DROP;
    }
    memcpy(x, v->m_shed, bytes);
    n = v->n_shed;
  }
  n = queue_lora_kv(v, row, x, t, n);
  *n_out = n;
  return n ? LIGO_OK : LIGO_EMPTY;
}

static int kth_set_bit(const uint64_t* x, int W, int k) {
  for (int w = 0; w < W; ++w) {
    int c = popc64(x[w]);
    if (k < c) {
      uint64_t bits = x[w];
      while (k--) bits &= bits - 1;
      return w * 64 + __builtin_ctzll(bits);
    }
    k -= c;
  }
  return -1;
}

typedef struct {
  const soa_view* v;
  const lig_oracle_req* reqs;
  lig_oracle_pick* out;
  uint32_t* masks;
  int lo, hi;
  uint64_t seed;
} soa_job;

static void* soa_worker(void* arg) {
  soa_job* j = (soa_job*)arg;
  const soa_view* v = j->v;
  uint64_t* x = (uint64_t*)calloc((size_t)(v->W64 > 0 ? v->W64 : 1) * 2, sizeof(uint64_t));
  uint64_t* t = x + (v->W64 > 0 ? v->W64 : 1);
  for (int i = j->lo; i < j->hi; ++i) {
    int n = 0;
    int st = ligo_soa_schedule_one(v, j->reqs[i].adapter_id, (int)(j->reqs[i].flags & 1u), x, t, &n);
    int pod = -1;
    if (st == LIGO_OK) {
      uint64_t state = j->seed ^ j->reqs[i].rand_key;
      pod = kth_set_bit(x, v->W64, lig_oracle_int31n(&state, n));        /* scheduler.go:120-121 */
    }
    j->out[i].pod_idx = pod;
    j->out[i].status = (uint16_t)st;
    j->out[i].n_survivors = (uint16_t)n;
    if (j->masks) {
      uint32_t* m = j->masks + (size_t)i * v->W32;
      for (int w = 0; w < v->W32; ++w)
        m[w] = (st == LIGO_OK) ? (uint32_t)(x[w >> 1] >> ((w & 1) * 32)) : 0u;
    }
  }
  free(x);
  return NULL;
}

void ligo_soa_view_init(soa_view* v, int P, int A, const double* kv, const int32_t* q,
                        const uint16_t* n_active, const uint16_t* max_active, const uint32_t* bitmap,
                        double kv_thr, int64_t q_crit, int64_t q_lora) {
  memset(v, 0, sizeof(*v));
  v->P = P; v->A = A; v->W64 = (P + 63) / 64; v->W32 = (P + 31) / 32;
  v->kv = kv; v->q = q; v->n_active = n_active; v->max_active = max_active; v->bitmap = bitmap;
  v->kv_thr = kv_thr; v->q_crit = q_crit; v->q_lora = q_lora;
  size_t words = (size_t)(v->W64 > 0 ? v->W64 : 1);
  uint64_t* pool = (uint64_t*)calloc(words * 4, sizeof(uint64_t));
  v->m_low = pool; v->m_room = pool + words; v->m_shed = pool + 2 * words; v->m_all = pool + 3 * words;
  for (int p = 0; p < P; ++p) {
    uint64_t bit = 1ull << (p & 63);
    v->m_all[p >> 6] |= bit;
    if ((int64_t)q[p] < q_lora) v->m_low[p >> 6] |= bit;
    if (n_active[p] < max_active[p]) v->m_room[p >> 6] |= bit;
    if ((int64_t)q[p] <= q_crit && kv[p] <= kv_thr) v->m_shed[p >> 6] |= bit;
  }
  v->n_low = mask_count(v->m_low, v->W64);
  v->n_shed = mask_count(v->m_shed, v->W64);
}

void ligo_soa_view_free(soa_view* v) { free(v->m_low); v->m_low = NULL; }

int lig_oracle_soa_schedule_batch(int P, int A, const double* kv, const int32_t* q,
                                  const uint16_t* n_active, const uint16_t* max_active,
                                  const uint32_t* bitmap, double kv_thr, int64_t q_crit,
                                  int64_t q_lora, const lig_oracle_req* reqs, int R, uint64_t seed,
                                  lig_oracle_pick* out, uint32_t* masks, int nthreads) {
  if (P < 0 || A < 0 || R < 0 || (R > 0 && (!reqs || !out))) return -1;
  soa_view v;
  ligo_soa_view_init(&v, P, A, kv, q, n_active, max_active, bitmap, kv_thr, q_crit, q_lora);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > R) nthreads = R > 0 ? R : 1;
  soa_job* jobs = (soa_job*)calloc((size_t)nthreads, sizeof(soa_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    soa_job j = {&v, reqs, out, masks, (int)((int64_t)R * t / nthreads),
                 (int)((int64_t)R * (t + 1) / nthreads), seed};
    jobs[t] = j;
  }
  if (nthreads == 1) {
    soa_worker(&jobs[0]);
  } else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, soa_worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  free(jobs);
  free(th);
  ligo_soa_view_free(&v);
  return 0;
}
