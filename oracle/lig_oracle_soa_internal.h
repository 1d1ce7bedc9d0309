/* lig_oracle_soa_internal.h — shared between lig_oracle_soa.c and lig_oracle_classtab.c.
 * TEST INFRASTRUCTURE ONLY (see lig_oracle.h). */
#ifndef LIG_ORACLE_SOA_INTERNAL_H_
#define LIG_ORACLE_SOA_INTERNAL_H_

#include <stdint.h>

typedef struct {
  int P, A, W64;            /* W64 = ceil(P / 64) */
  const double* kv;
  const int32_t* q;
  const uint16_t* n_active;
  const uint16_t* max_active;
  const uint32_t* bitmap;   /* adapter-major, A x ceil(P/32) 32-bit words */
  int W32;
  double kv_thr;
  int64_t q_crit, q_lora;
  /* request-independent masks, computed once per view */
  uint64_t* m_low;          /* q < q_lora                      filter.go:124-126 */
  uint64_t* m_room;         /* n_active < max_active           filter.go:175-177 */
  uint64_t* m_shed;         /* q <= q_crit && kv <= kv_thr     filter.go:183-187 */
  uint64_t* m_all;
  int n_low, n_shed;
} soa_view;

void ligo_soa_view_init(soa_view* v, int P, int A, const double* kv, const int32_t* q,
                        const uint16_t* n_active, const uint16_t* max_active, const uint32_t* bitmap,
                        double kv_thr, int64_t q_crit, int64_t q_lora);
void ligo_soa_view_free(soa_view* v);
/* One walk of the defaultFilter tree for (critical, adapter); x (W64 words) receives the survivor
 * mask, t is W64 words of scratch.  Returns the LIGO_* status. */
int ligo_soa_schedule_one(const soa_view* v, int adapter, int critical, uint64_t* x, uint64_t* t,
                          int* n_out);

#endif
