/* selftest.c — sanitizer harness for the oracle (ASan + UBSan), run by tests/test_oracle_sanitizers.py.
 * Exercises pool construction, the three TestFilter vectors of the reference
 * (pkg/ext-proc/scheduling/filter_test.go:28-200), degenerate inputs and the threaded batch. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lig_oracle.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(void) {
  const char* a1[] = {"foo", "bar"};
  const char* a2[] = {"foo", "critical"};
  const char* a3[] = {"foo"};
  lig_oracle_pool* p = lig_oracle_pool_new(3);
  lig_oracle_pool_set_pod(p, 0, "pod1", "", 0, 0.2, 2, a1, 2);
  lig_oracle_pool_set_pod(p, 1, "pod2", "", 3, 0.1, 2, a2, 2);
  lig_oracle_pool_set_pod(p, 2, "pod3", "", 10, 0.2, 2, a3, 1);
  int32_t idx[8];
  int n = 0;
  CHECK(lig_oracle_filter(p, "critical", 1, idx, &n) == LIGO_OK && n == 1 && idx[0] == 1);
  CHECK(lig_oracle_filter(p, "sheddable", 0, idx, &n) == LIGO_OK && n == 1 && idx[0] == 0);
  lig_oracle_pool_set_pod(p, 0, "pod1", "", 10, 0.9, 2, a1, 2);
  lig_oracle_pool_set_pod(p, 1, "pod2", "", 3, 0.85, 2, a2, 2);
  lig_oracle_pool_set_pod(p, 2, "pod3", "", 10, 0.85, 2, a3, 1);
  CHECK(lig_oracle_filter(p, "sheddable", 0, idx, &n) == LIGO_DROP && n == 0);
  /* degenerate values: NaN, inf, int64 extremes (Go wraps on overflow; must not trip UBSan) */
  lig_oracle_pool_set_pod(p, 0, "a", "", INT64_MAX, NAN, -5, NULL, 0);
  lig_oracle_pool_set_pod(p, 1, "b", "", INT64_MIN, INFINITY, INT64_MAX, a1, 2);
  lig_oracle_pool_set_pod(p, 2, "c", "", 0, -INFINITY, 0, a3, 1);
  for (int crit = 0; crit < 2; ++crit) {
    int rc = lig_oracle_filter(p, "foo", crit, idx, &n);
    CHECK(rc >= 0 && rc <= 3 && n >= 0 && n <= 3);
  }
  lig_oracle_pool* empty = lig_oracle_pool_new(0);
  CHECK(lig_oracle_filter(empty, "x", 1, idx, &n) == LIGO_DROP && n == 0);
  /* threaded batch */
  enum { R = 5000, P = 40, A = 6 };
  lig_oracle_pool* big = lig_oracle_pool_new(P);
  const char* names[A] = {"m0", "m1", "m2", "m3", "m4", "m5"};
  unsigned s = 12345;
  for (int i = 0; i < P; ++i) {
    const char* act[3];
    int na = 0;
    for (int k = 0; k < 3; ++k) { s = s * 1103515245u + 12345u; if ((s >> 16) & 1) act[na++] = names[(s >> 20) % A]; }
    s = s * 1103515245u + 12345u;
    lig_oracle_pool_set_pod(big, i, "p", "a", (s >> 16) % 70, ((s >> 8) % 1000) / 1000.0, (s >> 4) % 4, act, na);
  }
  lig_oracle_req* reqs = (lig_oracle_req*)calloc(R, sizeof(*reqs));
  lig_oracle_pick* o1 = (lig_oracle_pick*)calloc(R, sizeof(*o1));
  lig_oracle_pick* o4 = (lig_oracle_pick*)calloc(R, sizeof(*o4));
  uint32_t* masks = (uint32_t*)calloc((size_t)R * 2, sizeof(uint32_t));
  for (int i = 0; i < R; ++i) {
    s = s * 1103515245u + 12345u;
    reqs[i].adapter_id = (int32_t)((s >> 16) % (A + 2)) - 1;
    reqs[i].flags = (s >> 8) & 1;
    reqs[i].rand_key = ((uint64_t)s << 32) | (uint64_t)i;
  }
  CHECK(lig_oracle_schedule_batch(big, names, A, "base", reqs, R, 7, o1, masks, 1) == 0);
  CHECK(lig_oracle_schedule_batch(big, names, A, "base", reqs, R, 7, o4, NULL, 4) == 0);
  CHECK(memcmp(o1, o4, R * sizeof(*o1)) == 0);
  uint64_t st = 1234567;
  CHECK(lig_oracle_splitmix64_next(&st) == 6457827717110365317ull);
  free(reqs); free(o1); free(o4); free(masks);
  lig_oracle_pool_free(big);
  lig_oracle_pool_free(empty);
  lig_oracle_pool_free(p);
  printf("oracle selftest ok\n");
  return 0;
}
