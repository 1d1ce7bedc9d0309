// lig_device.cuh — sm_100a device code of the endpoint picker.
//
// Three kernels, all integer / FP64-compare work (no tensor cores; HBM + issue bound):
//   lig_class_build_kernel  per snapshot: the adapter-independent stages of the reference's filter
//                           tree once per CTA, then one warp per request class (critical?,
//                           adapter) finishes the walk and writes the class's survivor list.
//   lig_pick_stream_kernel  one thread per request: 16 B descriptor in, class lookup, Go Int31n
//                           draw, one 2 B gather from the class list, 8 B result out.  The
//                           bandwidth-bound stream the roofline is quoted on.
//   lig_scan_kernel         one warp per request: the same tree walk done per request with no
//                           class tables (the direct formulation; also returns survivor masks).
//
// The tree walked by tree_eval_warp() is the reference's defaultFilter, flattened:
//   pkg/ext-proc/scheduling/scheduler.go:26-91 (tree), filter.go:44-73 (success/failure routing),
//   filter.go:79-93 (predicate nodes), :102-122 (least queuing), :134-154 (least KV cache),
//   :124-126, :163-187 (predicates).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lig.h"

namespace lig {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerCta = 8;
constexpr int kCtaThreads = kWarpsPerCta * 32;

// Device view of one resident snapshot (pointers into the packed blob, see include/lig.h).
struct SnapView {
  const double* kv;
  const int* q;
  const uint16_t* n_active;
  const uint16_t* max_active;
  const uint32_t* bitmap;  // [A][W]
  int P, A, W;
};

struct Thr {
  double kv_thr;      // kvCacheThreshold        scheduler.go:17
  long long q_crit;   // queueThresholdCritical  scheduler.go:19
  long long q_lora;   // queueingThresholdLoRA   scheduler.go:23
};

// One entry per request class c = critical * (A + 1) + min(adapter, A); 16 bytes (one LDG.128).
// Classes whose survivor set does not depend on the adapter share one of two default lists
// (rows 2(A+1) and 2(A+1)+1 of the list pool); the others own row c.  List offsets are in list
// entries and fit 32 bits: (2 * 65535 + 3) rows x 32768 entries < 2^32.
//
//   info    : bits 0-1 status | 4-8 shift | 16-31 n
//             (info & 0xffff0003 is the second word of lig_pick as is: status | n_survivors << 16)
//   magic   : M = ceil(2^(32+shift) / n), shift = ceil(log2 n) - 1, so that for every v < 2^31
//             floor(v / n) == umulhi(v, M) >> shift   (Granlund-Montgomery, N = 31 bits; n >= 2)
//   q_limit : floor(2^31 / n).  With q = floor(v / n), Go's Int31n is k = v - q * n, resampling
//             while v > 2^31-1-(2^31 % n), i.e. while q >= floor(2^31 / n) (never for a power of
//             two, where v < 2^31 already implies q < 2^31 / n).
//   list_off: first entry of the class's survivor list in the list pool.  A class without
//             survivors points at the pool's sentinel entry 0xffff, which sign-extends to
//             pod_idx = -1: the pick needs no branch on n.
struct ClassEntry {
  uint32_t info;
  uint32_t magic;
  uint32_t q_limit;
  uint32_t list_off;
};

__host__ __device__ inline uint32_t entry_n(uint32_t info) { return info >> 16; }
__host__ __device__ inline uint32_t entry_status(uint32_t info) { return info & 3u; }
// Entries of the list pool: one row per class, two shared default rows, then the sentinel.
__host__ __device__ inline size_t list_pool_entries(size_t n_classes, size_t stride) {
  return (n_classes + 2) * stride + 8;
}

// Pod metric columns as the tree walk reads them: either the snapshot in global memory (read
// through the read-only path) or a copy the CTA staged into shared memory.
struct Fields {
  const double* kv;
  const int* q;
  const uint16_t* na;
  const uint16_t* ma;
};

template <bool kStaged>
__device__ __forceinline__ double ld_kv(const Fields& f, int p) {
  if constexpr (kStaged) return f.kv[p]; else return __ldg(f.kv + p);
}
template <bool kStaged>
__device__ __forceinline__ int ld_q(const Fields& f, int p) {
  if constexpr (kStaged) return f.q[p]; else return __ldg(f.q + p);
}
template <bool kStaged>
__device__ __forceinline__ bool has_room(const Fields& f, int p) {
  // canAcceptNewLoraPredicate: len(ActiveModels) < MaxActiveModels          filter.go:175-177
  if constexpr (kStaged) return f.na[p] < f.ma[p];
  else return __ldg(f.na + p) < __ldg(f.ma + p);
}

__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_int2(int2* p, int2 v) {
  asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1, %2};" :: "l"(p), "r"(v.x), "r"(v.y)
               : "memory");
}

// ---- the request's private random stream (include/lig.h) ---------------------------------------
__device__ __forceinline__ uint32_t splitmix_int31(uint64_t& state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 33);  // Int31() = Int63() >> 32 = next() >> 33
}

__device__ __forceinline__ uint32_t max_accept_for(uint32_t n) {
  return (n & (n - 1)) == 0 ? 0xffffffffu : 0x7fffffffu - (0x80000000u % n);
}

// rand.Intn(n) for 0 < n <= 2^31-1  ->  Int31n(n)          scheduler.go:120, math/rand Go 1.22
// (plain form, used by the direct-scan kernel)
__device__ __forceinline__ uint32_t int31n(uint64_t state, uint32_t n, uint32_t max_accept) {
  uint32_t v = splitmix_int31(state);
  if (max_accept == 0xffffffffu) return v & (n - 1);
  while (v > max_accept) v = splitmix_int31(state);
  return v % n;
}

// The same draw with the class entry's precomputed magic and rejection limit (no integer division
// and no branch on n on the hot path).  n <= 1: magic 0 gives q = 0 < q_limit, k = 0.
__device__ __forceinline__ uint32_t int31n_entry(uint64_t state, const uint4 e) {
  const uint32_t n = e.x >> 16;
  const uint32_t shift = (e.x >> 4) & 31u;
  uint32_t v = splitmix_int31(state);
  uint32_t q = __umulhi(v, e.y) >> shift;
  while (q >= e.z) {   // probability < n / 2^31 per draw
    v = splitmix_int31(state);
    q = __umulhi(v, e.y) >> shift;
  }
  return n > 1u ? v - q * n : 0u;
}

// ---- one pass helpers; X is a per-warp mask of W words in shared memory -----------------------
// "member" = bit `lane` of word w.  Every lane reads the same word (broadcast), evaluates its own
// pod, and the ballot is the new word: survivor order is pod order by construction.

// leastQueuingFilterFunc                                                    filter.go:102-122
template <bool kStaged>
__device__ __forceinline__ uint32_t stage_least_queuing(const Fields& f, uint32_t* X, int W,
                                                         int lane, uint32_t n) {
  if (n == 0) return 0;  // empty in => empty out, no division                filter_test.go:226-231
  int mn = 0x7fffffff;   // math.MaxInt: lowered by the first member           filter.go:103
  int mx = 0;            // max starts at 0, not MinInt                        filter.go:104
  for (int w = 0; w < W; ++w) {
    uint32_t word = X[w];
    if (word == 0) continue;
    if ((word >> lane) & 1u) {
      int v = ld_q<kStaged>(f, w * 32 + lane);
      mn = min(mn, v);
      mx = max(mx, v);
    }
  }
  mn = __reduce_min_sync(kFull, mn);
  mx = __reduce_max_sync(kFull, mx);
  // min + (max-min)/len(pods), Go int64 truncated division.  mx >= mn always (mx >= every
  // member, or 0 >= all-negative members), so the range is a non-negative value < 2^32.
  uint32_t range = (uint32_t)mx - (uint32_t)mn;
  long long thr = (long long)mn + (long long)(range / n);
  uint32_t cnt = 0;
  for (int w = 0; w < W; ++w) {
    uint32_t word = X[w];
    if (word == 0) continue;
    bool keep = false;
    if ((word >> lane) & 1u) {
      long long v = ld_q<kStaged>(f, w * 32 + lane);
      keep = v >= (long long)mn && v <= thr;                                 // filter.go:117
    }
    uint32_t nw = __ballot_sync(kFull, keep);
    __syncwarp();  // every lane's read of X[w] is ordered before lane 0 overwrites it
    if (lane == 0) X[w] = nw;
    cnt += __popc(nw);
  }
  __syncwarp();
  return cnt;
}

// leastKVCacheFilterFunc                                                    filter.go:134-154
template <bool kStaged>
__device__ __forceinline__ uint32_t stage_least_kv(const Fields& f, uint32_t* X, int W, int lane,
                                                   uint32_t n) {
  if (n == 0) return 0;                                                      // filter_test.go:265-270
  double mn = 1.7976931348623157e308;  // math.MaxFloat64                     filter.go:135
  double mx = 0.0;                     //                                      filter.go:136
  for (int w = 0; w < W; ++w) {
    uint32_t word = X[w];
    if (word == 0) continue;
    if ((word >> lane) & 1u) {
      double v = ld_kv<kStaged>(f, w * 32 + lane);
      if (v <= mn) mn = v;  // NaN compares false: never updates, as in Go    filter.go:140-145
      if (v >= mx) mx = v;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {  // mn/mx are never NaN here
    double o = __shfl_xor_sync(kFull, mn, off);
    if (o < mn) mn = o;
    o = __shfl_xor_sync(kFull, mx, off);
    if (o > mx) mx = o;
  }
  // min + (max-min)/float64(len(pods)): three separately rounded binary64 ops, no FMA.
  double thr = __dadd_rn(mn, __ddiv_rn(__dsub_rn(mx, mn), (double)n));       // filter.go:149
  uint32_t cnt = 0;
  for (int w = 0; w < W; ++w) {
    uint32_t word = X[w];
    if (word == 0) continue;
    bool keep = false;
    if ((word >> lane) & 1u) {
      double v = ld_kv<kStaged>(f, w * 32 + lane);
      keep = v >= mn && v <= thr;
    }
    uint32_t nw = __ballot_sync(kFull, keep);
    __syncwarp();  // every lane's read of X[w] is ordered before lane 0 overwrites it
    if (lane == 0) X[w] = nw;
    cnt += __popc(nw);
  }
  __syncwarp();
  return cnt;
}

struct EvalResult {
  uint32_t* mask;   // X or T: the W survivor words (shared memory, per warp)
  uint32_t n;
  uint32_t status;  // LIG_OK / LIG_DROP / LIG_EMPTY
};

// Walk defaultFilter for one (critical, adapter row) over all P pods, one warp.
//   X, T : per-warp scratch masks (W words each);  H : the adapter's bitmap row staged in shared
//   memory (W words), or nullptr when the adapter is in no pod's ActiveModels.
template <bool kStaged>
__device__ __forceinline__ EvalResult tree_eval_warp(const Fields& f, const uint32_t* H,
                                                     bool critical, const Thr thr, int P, int W,
                                                     uint32_t* X, uint32_t* T, int lane) {
  uint32_t n = 0;
  // criticalRequestPredicate keeps every pod or none; with P == 0 the predicate node yields
  // "no pods left" and the sheddable branch runs.            scheduler.go:26-31, filter.go:179-181
  if (critical && P > 0) {
    // "low queueing filter": q < queueingThresholdLoRA        scheduler.go:58-60, filter.go:124-126
    for (int w = 0; w < W; ++w) {
      int p = w * 32 + lane;
      bool keep = p < P && (long long)ld_q<kStaged>(f, p) < thr.q_lora;
      uint32_t nw = __ballot_sync(kFull, keep);
      if (lane == 0) X[w] = nw;
      n += __popc(nw);
    }
    __syncwarp();
    if (n > 0) {
      // "affinity LoRA": ResolvedTargetModel in ActiveModels   scheduler.go:61-64, filter.go:169-172
      uint32_t nb = 0;
      for (int w = lane; w < W; w += 32) {
        uint32_t t = H ? (X[w] & H[w]) : 0u;
        T[w] = t;
        nb += __popc(t);
      }
      nb = __reduce_add_sync(kFull, nb);
      __syncwarp();
      if (nb > 0) {
        uint32_t* s = X; X = T; T = s;
        n = nb;
      } else {
        // "can accept LoRA Adapter"                            scheduler.go:65-69, filter.go:175-177
        uint32_t nc = 0;
        for (int w = 0; w < W; ++w) {
          uint32_t word = X[w];
          bool keep = ((word >> lane) & 1u) && has_room<kStaged>(f, w * 32 + lane);
          uint32_t nw = __ballot_sync(kFull, keep);
          if (lane == 0) T[w] = nw;
          nc += __popc(nw);
        }
        __syncwarp();
        if (nc > 0) {  // on failure the node's INPUT (the low-queue set) is forwarded  filter.go:71
          uint32_t* s = X; X = T; T = s;
          n = nc;
        }
      }
      // queueAndKVCacheFilter                                   scheduler.go:49-56
      n = stage_least_queuing<kStaged>(f, X, W, lane, n);
      n = stage_least_kv<kStaged>(f, X, W, lane, n);
      return {X, n, n ? (uint32_t)LIG_OK : (uint32_t)LIG_EMPTY};
    }
    // low-queueing failed: its input (all pods) goes to queueLoRAAndKVCacheFilter   scheduler.go:71
    for (int w = lane; w < W; w += 32) {
      int rem = P - w * 32;
      X[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    __syncwarp();
    n = (uint32_t)P;
  } else {
    // "has capacity for sheddable requests"                     scheduler.go:74-79, filter.go:183-187
    for (int w = 0; w < W; ++w) {
      int p = w * 32 + lane;
      bool keep = false;
      if (p < P) {
        keep = (long long)ld_q<kStaged>(f, p) <= thr.q_crit && ld_kv<kStaged>(f, p) <= thr.kv_thr;
      }
      uint32_t nw = __ballot_sync(kFull, keep);
      if (lane == 0) X[w] = nw;
      n += __popc(nw);
    }
    __syncwarp();
    if (n == 0) return {X, 0u, (uint32_t)LIG_DROP};              // "drop request"  scheduler.go:83-89
  }
  // queueLoRAAndKVCacheFilter: least queuing -> low cost LoRA -> least KV    scheduler.go:35-46
  n = stage_least_queuing<kStaged>(f, X, W, lane, n);
  uint32_t nz = 0;
  for (int w = 0; w < W; ++w) {
    uint32_t word = X[w];
    if (word == 0) { if (lane == 0) T[w] = 0; continue; }
    uint32_t hw = H ? H[w] : 0u;
    // lowLoRACostPredicate: affinity OR room                               filter.go:163-166
    bool keep = ((word >> lane) & 1u) && (((hw >> lane) & 1u) || has_room<kStaged>(f, w * 32 + lane));
    uint32_t nw = __ballot_sync(kFull, keep);
    if (lane == 0) T[w] = nw;
    nz += __popc(nw);
  }
  __syncwarp();
  if (nz > 0) {
    uint32_t* s = X; X = T; T = s;
    n = nz;
  }
  n = stage_least_kv<kStaged>(f, X, W, lane, n);
  return {X, n, n ? (uint32_t)LIG_OK : (uint32_t)LIG_EMPTY};
}

// Shared memory carve-up common to the two tree-walking kernels:
//   [ per-warp scratch: kWarpsPerCta x 3 x W words ][ staged pod columns (kStaged only) ]
__host__ __device__ inline size_t scratch_bytes(int W) {
  return (size_t)kWarpsPerCta * 3u * (size_t)W * sizeof(uint32_t);
}
__host__ __device__ inline size_t staged_bytes(int W) {  // kv f64 + q i32 + na u16 + ma u16 = 16 B/pod
  return (size_t)W * 32u * 16u;
}

// Coalesced 16-byte copies of the four pod columns into shared memory (every column of the
// packed blob is 16-byte aligned and padded to 32 pods).
__device__ __forceinline__ Fields stage_fields(const SnapView& s, unsigned char* smem) {
  const int Ppad = s.W * 32;
  double* kv = reinterpret_cast<double*>(smem);
  int* q = reinterpret_cast<int*>(kv + Ppad);
  uint16_t* na = reinterpret_cast<uint16_t*>(q + Ppad);
  uint16_t* ma = na + Ppad;
  const int4* src_kv = reinterpret_cast<const int4*>(s.kv);
  const int4* src_q = reinterpret_cast<const int4*>(s.q);
  const int4* src_na = reinterpret_cast<const int4*>(s.n_active);
  const int4* src_ma = reinterpret_cast<const int4*>(s.max_active);
  int4* dkv = reinterpret_cast<int4*>(kv);
  int4* dq = reinterpret_cast<int4*>(q);
  int4* dna = reinterpret_cast<int4*>(na);
  int4* dma = reinterpret_cast<int4*>(ma);
  for (int i = threadIdx.x; i < Ppad / 2; i += blockDim.x) dkv[i] = __ldg(src_kv + i);
  for (int i = threadIdx.x; i < Ppad / 4; i += blockDim.x) dq[i] = __ldg(src_q + i);
  for (int i = threadIdx.x; i < Ppad / 8; i += blockDim.x) {
    dna[i] = __ldg(src_na + i);
    dma[i] = __ldg(src_ma + i);
  }
  return Fields{kv, q, na, ma};
}

// Stage one adapter's bitmap row through shared memory; nullptr for an adapter outside [0, A).
__device__ __forceinline__ const uint32_t* stage_adapter_row(const SnapView& s, int adapter,
                                                             uint32_t* H, int lane) {
  if (adapter < 0 || adapter >= s.A) return nullptr;  // Go map miss on every pod  filter.go:170
  const uint32_t* row = s.bitmap + (size_t)adapter * s.W;
  for (int w = lane; w < s.W; w += 32) H[w] = __ldg(row + w);
  __syncwarp();
  return H;
}

// ---- K2a: class tables ---------------------------------------------------------------------------
// Class c = critical * (A + 1) + a, a in [0, A] (a == A: adapter active nowhere).
//
// Most of the tree does not depend on the adapter: the low-queue set, the
// "has room" set, the sheddable-capacity set and the least-queuing stage that follows them are
// the same for every class.  Each CTA therefore first walks those shared stages once with all its
// warps (dense, pod-parallel, ballot words), and a class then only costs one AND of its bitmap
// row with a shared mask; only classes whose adapter actually intersects the mask run their own
// range filters, on the (sparse) set bits.  Classes that do not intersect share one of two
// default survivor lists.  lig_scan_kernel keeps the plain per-request walk (tree_eval_warp), so
// the GPU holds two independent formulations of the tree (the parity tests check both).

constexpr int kBuildWarps = 16;   // the class build runs 512-thread CTAs, at most one per SM
constexpr int kBuildThreads = kBuildWarps * 32;

struct BuildShared {        // block-wide scalars of the shared stages (shared memory)
  uint32_t n_shed;          // |S|, S = {q <= q_crit && kv <= kv_thr}              scheduler.go:74-79
  uint32_t crit_mode;       // 0: low-queue set non-empty; 1: empty (all pods go to queueLoRAAndKV)
  uint32_t rc_n, rc_status; // default result of critical classes
  uint32_t rs_n, rs_status; // default result of sheddable classes
  uint32_t red_u[kBuildWarps];
  int red_i[2 * kBuildWarps];
  double red_d[2 * kBuildWarps];
};

__device__ __forceinline__ uint32_t blk_sum(uint32_t warp_value, BuildShared* sh, int warp, int lane) {
  __syncthreads();
  if (lane == 0) sh->red_u[warp] = warp_value;
  __syncthreads();
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < kBuildWarps; ++i) t += sh->red_u[i];
  return t;
}

// out[w] = ballot(member(in, w) && pred(p)) for all words, block-wide; returns the block count.
template <class Pred>
__device__ __forceinline__ uint32_t blk_pred_pass(uint32_t* out, const uint32_t* in, int P, int W,
                                                  BuildShared* sh, int warp, int lane, Pred pred) {
  uint32_t cnt = 0;
  for (int w = warp; w < W; w += kBuildWarps) {
    const int p = w * 32 + lane;
    const bool member = in ? ((in[w] >> lane) & 1u) : (p < P);
    const bool keep = member && pred(p);
    const uint32_t nw = __ballot_sync(kFull, keep);
    if (lane == 0) out[w] = nw;
    cnt += __popc(nw);
  }
  return blk_sum(cnt, sh, warp, lane);
}

// leastQueuingFilterFunc over mask X (in place), block-wide.                 filter.go:102-122
template <bool kStaged>
__device__ __forceinline__ uint32_t blk_least_queuing(const Fields& f, uint32_t* X, int W, uint32_t n,
                                                      BuildShared* sh, int warp, int lane) {
  if (n == 0) return 0;
  int mn = 0x7fffffff, mx = 0;
  for (int w = warp; w < W; w += kBuildWarps) {
    if ((X[w] >> lane) & 1u) {
      const int v = ld_q<kStaged>(f, w * 32 + lane);
      mn = min(mn, v);
      mx = max(mx, v);
    }
  }
  mn = __reduce_min_sync(kFull, mn);
  mx = __reduce_max_sync(kFull, mx);
  __syncthreads();
  if (lane == 0) { sh->red_i[warp] = mn; sh->red_i[kBuildWarps + warp] = mx; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kBuildWarps; ++i) {
    mn = min(mn, sh->red_i[i]);
    mx = max(mx, sh->red_i[kBuildWarps + i]);
  }
  const uint32_t range = (uint32_t)mx - (uint32_t)mn;   // see stage_least_queuing
  const long long thr = (long long)mn + (long long)(range / n);
  uint32_t cnt = 0;
  for (int w = warp; w < W; w += kBuildWarps) {
    const uint32_t word = X[w];
    bool keep = false;
    if ((word >> lane) & 1u) {
      const long long v = ld_q<kStaged>(f, w * 32 + lane);
      keep = v >= (long long)mn && v <= thr;
    }
    const uint32_t nw = __ballot_sync(kFull, keep);
    __syncwarp();  // every lane's read of X[w] is ordered before lane 0 overwrites it
    if (lane == 0) X[w] = nw;
    cnt += __popc(nw);
  }
  return blk_sum(cnt, sh, warp, lane);
}

// leastKVCacheFilterFunc over mask X (in place), block-wide.                 filter.go:134-154
template <bool kStaged>
__device__ __forceinline__ uint32_t blk_least_kv(const Fields& f, uint32_t* X, int W, uint32_t n,
                                                 BuildShared* sh, int warp, int lane) {
  if (n == 0) return 0;
  double mn = 1.7976931348623157e308, mx = 0.0;
  for (int w = warp; w < W; w += kBuildWarps) {
    if ((X[w] >> lane) & 1u) {
      const double v = ld_kv<kStaged>(f, w * 32 + lane);
      if (v <= mn) mn = v;
      if (v >= mx) mx = v;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double o = __shfl_xor_sync(kFull, mn, off);
    if (o < mn) mn = o;
    o = __shfl_xor_sync(kFull, mx, off);
    if (o > mx) mx = o;
  }
  __syncthreads();
  if (lane == 0) { sh->red_d[warp] = mn; sh->red_d[kBuildWarps + warp] = mx; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kBuildWarps; ++i) {
    const double a = sh->red_d[i], b = sh->red_d[kBuildWarps + i];
    if (a < mn) mn = a;
    if (b > mx) mx = b;
  }
  const double thr = __dadd_rn(mn, __ddiv_rn(__dsub_rn(mx, mn), (double)n));
  uint32_t cnt = 0;
  for (int w = warp; w < W; w += kBuildWarps) {
    const uint32_t word = X[w];
    bool keep = false;
    if ((word >> lane) & 1u) {
      const double v = ld_kv<kStaged>(f, w * 32 + lane);
      keep = v >= mn && v <= thr;
    }
    const uint32_t nw = __ballot_sync(kFull, keep);
    __syncwarp();  // every lane's read of X[w] is ordered before lane 0 overwrites it
    if (lane == 0) X[w] = nw;
    cnt += __popc(nw);
  }
  return blk_sum(cnt, sh, warp, lane);
}

// The same two range filters for ONE warp on a sparse mask: lane l owns words l, l+32, ... and
// walks their set bits.
template <bool kStaged>
__device__ __forceinline__ uint32_t sparse_least_queuing(const Fields& f, uint32_t* X, int W, int lane,
                                                         uint32_t n) {
  if (n == 0) return 0;
  int mn = 0x7fffffff, mx = 0;
  for (int w = lane; w < W; w += 32) {
    uint32_t bits = X[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int v = ld_q<kStaged>(f, w * 32 + b);
      mn = min(mn, v);
      mx = max(mx, v);
    }
  }
  mn = __reduce_min_sync(kFull, mn);
  mx = __reduce_max_sync(kFull, mx);
  const uint32_t range = (uint32_t)mx - (uint32_t)mn;
  const long long thr = (long long)mn + (long long)(range / n);
  uint32_t cnt = 0;
  for (int w = lane; w < W; w += 32) {
    uint32_t bits = X[w], nw = 0;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const long long v = ld_q<kStaged>(f, w * 32 + b);
      if (v >= (long long)mn && v <= thr) nw |= 1u << b;
    }
    X[w] = nw;
    cnt += __popc(nw);
  }
  cnt = __reduce_add_sync(kFull, cnt);
  __syncwarp();
  return cnt;
}

template <bool kStaged>
__device__ __forceinline__ uint32_t sparse_least_kv(const Fields& f, uint32_t* X, int W, int lane,
                                                    uint32_t n) {
  if (n == 0) return 0;
  double mn = 1.7976931348623157e308, mx = 0.0;
  for (int w = lane; w < W; w += 32) {
    uint32_t bits = X[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const double v = ld_kv<kStaged>(f, w * 32 + b);
      if (v <= mn) mn = v;
      if (v >= mx) mx = v;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double o = __shfl_xor_sync(kFull, mn, off);
    if (o < mn) mn = o;
    o = __shfl_xor_sync(kFull, mx, off);
    if (o > mx) mx = o;
  }
  const double thr = __dadd_rn(mn, __ddiv_rn(__dsub_rn(mx, mn), (double)n));
  uint32_t cnt = 0;
  for (int w = lane; w < W; w += 32) {
    uint32_t bits = X[w], nw = 0;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const double v = ld_kv<kStaged>(f, w * 32 + b);
      if (v >= mn && v <= thr) nw |= 1u << b;
    }
    X[w] = nw;
    cnt += __popc(nw);
  }
  cnt = __reduce_add_sync(kFull, cnt);
  __syncwarp();
  return cnt;
}

// Write the set bits of X, ascending, to list[0..n): lane-strided words, warp prefix per 32 words.
__device__ __forceinline__ void compact_mask_to_list(const uint32_t* X, int W, int lane,
                                                     uint16_t* __restrict__ list) {
  uint32_t base = 0;
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = w < W ? X[w] : 0u;
    const uint32_t c = __popc(bits);
    uint32_t incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += o;
    }
    uint32_t pos = base + incl - c;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      list[pos++] = (uint16_t)(w * 32 + b);
    }
    base += __shfl_sync(kFull, incl, 31);
  }
}

// Shared-memory carve-up of the class build:
//   [ BuildShared ][ 7 block masks x W ][ per-warp scratch kBuildWarps x W ][ staged columns ]
__host__ __device__ inline size_t build_fixed_bytes(int W) {
  const size_t b = ((sizeof(BuildShared) + 15) & ~(size_t)15) +
                   (size_t)(7 + kBuildWarps) * (size_t)W * sizeof(uint32_t);
  return (b + 15) & ~(size_t)15;   // the staged columns behind it are written with 16-byte stores
}

__device__ __forceinline__ ClassEntry make_entry(uint32_t n, uint32_t status, uint32_t list_off) {
  ClassEntry e;
  // n <= 1: magic 0 gives q = 0 for every draw and the limit 1 accepts the first one: k = 0.
  uint32_t shift = 0, magic = 0, q_limit = 1;
  if (n >= 2) {
    const uint32_t l = 32u - (uint32_t)__clz(n - 1u);          // ceil(log2 n), 1..15
    shift = l - 1u;
    // ceil(2^(32+shift) / n) < 2^32 because n > 2^(l-1)
    magic = (uint32_t)((((unsigned long long)1 << (32u + shift)) + n - 1u) / n);
    q_limit = 0x80000000u / n;
  }
  e.info = status | (shift << 4) | (n << 16);
  e.magic = magic;
  e.q_limit = q_limit;
  e.list_off = list_off;
  return e;
}

// ---- compact tables -----------------------------------------------------------------------------
// The survivor lists are tiny in practice (a handful of pods per class after the least-KV stage).
// Besides the strided rows (one row of P entries per class: every warp works alone, and
// lig_read_class can address a class directly) the build packs them behind a copy of the class
// entries into ONE contiguous blob that the persistent pick kernels pull into shared memory with a
// single TMA bulk copy:
//
//   [ header 16 B ][ entries: n_classes x 16 B, list_off relative to the pool ][ pool: u16[] ]
//
// pool[0] is the 0xffff sentinel (pod_idx -1) every class without survivors points at, then the
// two default lists (stored once), then the own lists in the order their warps finished (pool
// space is handed out by an atomic cursor).  If the pool does not fit `pool_capacity` entries the
// header says so (bytes = 0) and the pick kernels keep to the strided tables in global memory.
struct CompactHeader {
  uint32_t bytes;         // header + entries + pool, rounded up to 16; 0 = not available
  uint32_t n_classes;
  uint32_t pool_entries;
  uint32_t reserved;
};
struct CompactOut {       // where the build writes the blob; counters[0] = pool cursor of the own
  unsigned char* blob;    // lists, counters[1] = finished CTAs (zero at allocation, reset by the sealing CTA)
  uint32_t* counters;
  uint32_t pool_capacity;
  CompactHeader* host_hdr;   // page-locked, device-mapped copy of the header for the host's launch
                             // decision: written by the sealing thread, no D2H copy on the stream
  unsigned long long* dbg;   // LIG_BUILD_DEBUG: clock64() stamps of CTA 0 / thread 0 (nullable)
};
#define LIG_STAMP(k) do { if (co.dbg && blockIdx.x == 0 && threadIdx.x == 0) co.dbg[k] = clock64(); } while (0)

template <bool kStaged>
__global__ void __launch_bounds__(kBuildThreads)
lig_class_build_kernel(SnapView s, Thr thr, ClassEntry* __restrict__ cls,
                       uint16_t* __restrict__ lists, int list_stride, CompactOut co) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = s.P, W = s.W, A = s.A;
  BuildShared* sh = reinterpret_cast<BuildShared*>(smem);
  uint32_t* masks = reinterpret_cast<uint32_t*>(smem + ((sizeof(BuildShared) + 15) & ~(size_t)15));
  uint32_t* M_room = masks;            // n_active < max_active                       filter.go:175-177
  uint32_t* CB = masks + 1 * W;        // critical: mask the adapter row is ANDed with
  uint32_t* CZ = masks + 2 * W;        // critical mode 1: (least-queuing set) & room
  uint32_t* SB = masks + 3 * W;        // sheddable: (least-queuing set of S) & ~room
  uint32_t* SZ = masks + 4 * W;        // sheddable: (least-queuing set of S) & room
  uint32_t* TMP = masks + 5 * W;
  uint32_t* SHED = masks + 6 * W;      // q <= q_crit && kv <= kv_thr                 filter.go:183-187
  uint32_t* X = masks + (size_t)(7 + warp) * W;   // per-warp scratch
  Fields f{s.kv, s.q, s.n_active, s.max_active};
  if constexpr (kStaged) {
    f = stage_fields(s, smem + build_fixed_bytes(W));
  }
  __syncthreads();
  const int n_classes = 2 * (A + 1);
  const uint32_t rc_row = (uint32_t)n_classes * (uint32_t)list_stride;        // default list rows
  const uint32_t rs_row = rc_row + (uint32_t)list_stride;
  const uint32_t none = rs_row + (uint32_t)list_stride;                       // sentinel entry
  if (blockIdx.x == 0 && threadIdx.x == 0) lists[none] = 0xffffu;             // reads back as pod_idx -1
  CompactHeader* chdr = reinterpret_cast<CompactHeader*>(co.blob);
  uint4* centries = reinterpret_cast<uint4*>(co.blob + sizeof(CompactHeader));
  uint16_t* cpool = reinterpret_cast<uint16_t*>(co.blob + sizeof(CompactHeader) + (size_t)n_classes * sizeof(ClassEntry));
  // the last CTA to finish seals the blob: total size, or 0 when the pool overflowed
  auto seal = [&](uint32_t fixed_entries) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(co.counters + 1, 1u) == gridDim.x - 1) {
        __threadfence();
        const uint32_t total = fixed_entries + *reinterpret_cast<volatile uint32_t*>(co.counters);
        chdr->n_classes = (uint32_t)n_classes;
        chdr->pool_entries = total;
        chdr->reserved = 0;
        chdr->bytes = total <= co.pool_capacity
                          ? (uint32_t)((sizeof(CompactHeader) + (size_t)n_classes * sizeof(ClassEntry) +
                                        (size_t)total * sizeof(uint16_t) + 15) & ~(size_t)15) : 0u;
        if (co.host_hdr) {
          co.host_hdr->n_classes = chdr->n_classes;
          co.host_hdr->pool_entries = total;
          co.host_hdr->reserved = 0;
          co.host_hdr->bytes = chdr->bytes;
        }
        co.counters[0] = 0;   // ready for the next build of this slot
        co.counters[1] = 0;
      }
    }
  };

  if (P == 0) {   // critical: predicate node errs on an empty pool -> sheddable branch -> drop
    for (int c = blockIdx.x * kBuildThreads + threadIdx.x; c < n_classes; c += gridDim.x * kBuildThreads) {
      const ClassEntry e = make_entry(0u, (uint32_t)LIG_DROP, none);
      cls[c] = e;
      centries[c] = make_uint4(e.info, e.magic, e.q_limit, 0u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cpool[0] = 0xffffu;
    seal(1u);
    return;
  }

  // ---- shared stages (every CTA, all its warps) -----------------------------------------------
  auto ldq = [&](int p) { return (long long)ld_q<kStaged>(f, p); };
  // one pass over the pods for the three request-independent predicate sets
  uint32_t n_low = 0, n_shed_w = 0;
  for (int w = warp; w < W; w += kBuildWarps) {
    const int p = w * 32 + lane;
    bool room = false, low = false, shed = false;
    if (p < P) {
      const long long qq = ldq(p);
      room = has_room<kStaged>(f, p);                                     // filter.go:175-177
      low = qq < thr.q_lora;                                              // scheduler.go:58-60
      shed = qq <= thr.q_crit && ld_kv<kStaged>(f, p) <= thr.kv_thr;      // scheduler.go:74-79
    }
    const uint32_t wr = __ballot_sync(kFull, room), wl = __ballot_sync(kFull, low),
                   ws = __ballot_sync(kFull, shed);
    if (lane == 0) { M_room[w] = wr; CB[w] = wl; SHED[w] = ws; }
    n_low += __popc(wl);
    n_shed_w += __popc(ws);
  }
  n_low = blk_sum(n_low, sh, warp, lane);
  const uint32_t n_shed = blk_sum(n_shed_w, sh, warp, lane);
  uint32_t rc_n, rc_status;
  if (n_low > 0) {
    // default (adapter active in none of the low-queue pods): "can accept LoRA Adapter" on the
    // low-queue set, falling back to that set, then queueAndKVCacheFilter   scheduler.go:65-69,49-56
    uint32_t nc = 0;
    for (int w = warp; w < W; w += kBuildWarps) {
      const uint32_t t = CB[w] & M_room[w];
      if (lane == 0) TMP[w] = t;
      nc += __popc(t);
    }
    nc = blk_sum(nc, sh, warp, lane);
    if (nc == 0) {
      for (int w = threadIdx.x; w < W; w += kBuildThreads) TMP[w] = CB[w];
      nc = n_low;
    }
    __syncthreads();
    nc = blk_least_queuing<kStaged>(f, TMP, W, nc, sh, warp, lane);
    rc_n = blk_least_kv<kStaged>(f, TMP, W, nc, sh, warp, lane);
    rc_status = rc_n ? LIG_OK : LIG_EMPTY;
    if (blockIdx.x == 0 && warp == 0 && rc_n) {
      compact_mask_to_list(TMP, W, lane, lists + rc_row);
      compact_mask_to_list(TMP, W, lane, cpool + 1);
    }
    __syncthreads();
  } else {
    // low-queue filter failed: all pods -> least queuing -> low cost LoRA -> least KV  scheduler.go:71,35-46
    uint32_t ny = blk_pred_pass(TMP, nullptr, P, W, sh, warp, lane, [&](int) { return true; });
    ny = blk_least_queuing<kStaged>(f, TMP, W, ny, sh, warp, lane);
    uint32_t nz = 0;
    for (int w = warp; w < W; w += kBuildWarps) {
      const uint32_t y = TMP[w], r = M_room[w];
      if (lane == 0) { CZ[w] = y & r; CB[w] = y & ~r; }
      nz += __popc(y & r);
    }
    nz = blk_sum(nz, sh, warp, lane);
    if (nz > 0) {
      for (int w = threadIdx.x; w < W; w += kBuildThreads) TMP[w] = CZ[w];
      ny = nz;
    }
    __syncthreads();
    rc_n = blk_least_kv<kStaged>(f, TMP, W, ny, sh, warp, lane);
    rc_status = rc_n ? LIG_OK : LIG_EMPTY;
    if (blockIdx.x == 0 && warp == 0 && rc_n) {
      compact_mask_to_list(TMP, W, lane, lists + rc_row);
      compact_mask_to_list(TMP, W, lane, cpool + 1);
    }
    __syncthreads();
  }
  // sheddable side: "has capacity for sheddable requests"                 scheduler.go:74-79
  uint32_t rs_n = 0, rs_status = LIG_DROP;
  for (int w = threadIdx.x; w < W; w += kBuildThreads) TMP[w] = SHED[w];
  __syncthreads();
  if (n_shed > 0) {
    uint32_t ny = blk_least_queuing<kStaged>(f, TMP, W, n_shed, sh, warp, lane);
    uint32_t nz = 0;
    for (int w = warp; w < W; w += kBuildWarps) {
      const uint32_t y = TMP[w], r = M_room[w];
      if (lane == 0) { SZ[w] = y & r; SB[w] = y & ~r; }
      nz += __popc(y & r);
    }
    nz = blk_sum(nz, sh, warp, lane);
    if (nz > 0) {
      for (int w = threadIdx.x; w < W; w += kBuildThreads) TMP[w] = SZ[w];
      ny = nz;
    }
    __syncthreads();
    rs_n = blk_least_kv<kStaged>(f, TMP, W, ny, sh, warp, lane);
    rs_status = rs_n ? LIG_OK : LIG_EMPTY;
    if (blockIdx.x == 0 && warp == 0 && rs_n) {
      compact_mask_to_list(TMP, W, lane, lists + rs_row);
      compact_mask_to_list(TMP, W, lane, cpool + 1 + rc_n);
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) cpool[0] = 0xffffu;
  const uint32_t fixed_entries = 1u + rc_n + rs_n;   // sentinel + the two default lists

  // ---- per class (one warp each): AND the adapter row with the shared mask --------------------
  for (int c = blockIdx.x * kBuildWarps + warp; c < n_classes; c += gridDim.x * kBuildWarps) {
    const bool critical = c >= A + 1;
    const int a = critical ? c - (A + 1) : c;
    const uint32_t* row = a < A ? s.bitmap + (size_t)a * W : nullptr;
    ClassEntry e;
    uint32_t coff = 0;                   // the class's list in the compact pool (0 = the sentinel)
    if (!critical && n_shed == 0) {
      e = make_entry(0u, (uint32_t)LIG_DROP, none);                        // scheduler.go:83-89
    } else {
      const uint32_t* Bm = critical ? CB : SB;
      const uint32_t* Zm = critical ? (n_low > 0 ? nullptr : CZ) : SZ;
      uint32_t hit = 0;
      if (row) {
        for (int w = lane; w < W; w += 32) {
          const uint32_t t = Bm[w] & __ldg(row + w);
          X[w] = t;
          hit += __popc(t);
        }
        hit = __reduce_add_sync(kFull, hit);
      }
      if (hit == 0) {
        e = critical ? make_entry(rc_n, rc_status, rc_n ? rc_row : none)
                     : make_entry(rs_n, rs_status, rs_n ? rs_row : none);
        coff = critical ? (rc_n ? 1u : 0u) : (rs_n ? 1u + rc_n : 0u);
      } else {
        uint32_t n = hit;
        if (Zm) {   // low cost LoRA: (affinity | room) on the least-queuing set     filter.go:163-166
          n = 0;
          for (int w = lane; w < W; w += 32) {
            const uint32_t t = X[w] | Zm[w];
            X[w] = t;
            n += __popc(t);
          }
          n = __reduce_add_sync(kFull, n);
        }
        __syncwarp();
        if (critical && n_low > 0)      // "affinity LoRA" succeeded -> queueAndKVCacheFilter
          n = sparse_least_queuing<kStaged>(f, X, W, lane, n);
        n = sparse_least_kv<kStaged>(f, X, W, lane, n);
        const uint32_t off = (uint32_t)c * (uint32_t)list_stride;
        if (n) {
          compact_mask_to_list(X, W, lane, lists + off);
          uint32_t at = 0;
          if (lane == 0) at = atomicAdd(co.counters, n);           // pool space for this list
          at = __shfl_sync(kFull, at, 0) + fixed_entries;
          if (at + n <= co.pool_capacity) compact_mask_to_list(X, W, lane, cpool + at);
          coff = at;
        }
        e = make_entry(n, n ? (uint32_t)LIG_OK : (uint32_t)LIG_EMPTY, n ? off : none);
      }
    }
    if (lane == 0) {
      *reinterpret_cast<uint4*>(cls + c) = make_uint4(e.info, e.magic, e.q_limit, e.list_off);
      centries[c] = make_uint4(e.info, e.magic, e.q_limit, coff);
    }
    __syncwarp();
  }
  seal(fixed_entries);
}

// ---- K2a (fast form): pools of up to 4096 pods ------------------------------------------------------
// Same tables, bit for bit, as lig_class_build_kernel; the differences are all about latency (the
// build is a chain of a dozen tiny block-wide stages, so it is bound by synchronisation and
// dependent shared-memory round trips, not by work):
//   * the shared stages keep each thread's 8 pods (word warp + 16 j, bit lane) in REGISTERS, sets
//     are 8-bit masks per thread, a stage is a few compares + one warp reduction + ONE
//     __syncthreads (double-buffered scratch) instead of ballot passes over shared-memory masks;
//   * one class per warp (twice the CTAs), its bitmap row requested from HBM/L2 before the shared
//     stages start, pool space reserved (atomic) before the compaction, one compaction pass that
//     writes the strided row and the pool entry together, lanes owning contiguous words.
constexpr int kOwn = 8;                                 // words per warp: W <= kBuildWarps * kOwn
constexpr int kFastMaxWords = kBuildWarps * kOwn;       // 128 words = 4096 pods

struct FastShared {          // double-buffered scratch of the block-wide stages
  uint32_t red_u[2][4][kBuildWarps];
  int red_i[2][4][kBuildWarps];
  double red_d[2][4][kBuildWarps];
};

__host__ __device__ inline size_t fast_fixed_bytes(int W) {
  const size_t b = ((sizeof(FastShared) + 15) & ~(size_t)15) + (size_t)(7 + kBuildWarps) * (size_t)W * sizeof(uint32_t);
  return (b + 15) & ~(size_t)15;
}

// Set bits of X (ascending) -> out_a[0..n) and out_b[0..n); lane l owns the contiguous words
// [l * per, (l + 1) * per): one warp scan for the whole mask.
__device__ __forceinline__ void compact_mask_to_two_lists(const uint32_t* X, int W, int lane,
                                                          uint16_t* __restrict__ out_a,
                                                          uint16_t* __restrict__ out_b, bool write_b) {
  const int per = (W + 31) / 32;
  const int w0 = lane * per;
  uint32_t c = 0;
  for (int i = 0; i < per; ++i) c += (w0 + i < W) ? __popc(X[w0 + i]) : 0;
  uint32_t incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(kFull, incl, off);
    if (lane >= off) incl += o;
  }
  uint32_t pos = incl - c;
  for (int i = 0; i < per; ++i) {
    if (w0 + i >= W) break;
    uint32_t bits = X[w0 + i];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const uint16_t pod = (uint16_t)((w0 + i) * 32 + b);
      out_a[pos] = pod;
      if (write_b) out_b[pos] = pod;
      ++pos;
    }
  }
}

__global__ void __launch_bounds__(kBuildThreads)
lig_class_build_fast_kernel(SnapView s, Thr thr, ClassEntry* __restrict__ cls, uint16_t* __restrict__ lists,
                            int list_stride, CompactOut co) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = s.P, W = s.W, A = s.A;
  FastShared* sh = reinterpret_cast<FastShared*>(smem);
  uint32_t* masks = reinterpret_cast<uint32_t*>(smem + ((sizeof(FastShared) + 15) & ~(size_t)15));
  uint32_t* M_room = masks;            // n_active < max_active                       filter.go:175-177
  uint32_t* CB = masks + 1 * W;        // critical: mask the adapter row is ANDed with
  uint32_t* CZ = masks + 2 * W;        // critical mode 1: (least-queuing set) & room
  uint32_t* SB = masks + 3 * W;        // sheddable: (least-queuing set of S) & ~room
  uint32_t* SZ = masks + 4 * W;        // sheddable: (least-queuing set of S) & room
  uint32_t* TMPC = masks + 5 * W;      // the default critical survivor set on its way to a list
  uint32_t* TMPS = masks + 6 * W;      // the default sheddable one
  uint32_t* X = masks + (size_t)(7 + warp) * W;   // per-warp scratch
  const int n_classes = 2 * (A + 1);
  const uint32_t rc_row = (uint32_t)n_classes * (uint32_t)list_stride;
  const uint32_t rs_row = rc_row + (uint32_t)list_stride;
  const uint32_t none = rs_row + (uint32_t)list_stride;
  CompactHeader* chdr = reinterpret_cast<CompactHeader*>(co.blob);
  uint4* centries = reinterpret_cast<uint4*>(co.blob + sizeof(CompactHeader));
  uint16_t* cpool = reinterpret_cast<uint16_t*>(co.blob + sizeof(CompactHeader) + (size_t)n_classes * sizeof(ClassEntry));

  // this warp's first class: ask for its bitmap row now, use it after the shared stages
  // classes are dealt to the CTAs round-robin: the popular adapters (low ids, long sparse walks)
  // end up on different SMs instead of sharing one
  const int c0 = warp * (int)gridDim.x + (int)blockIdx.x;
  uint32_t rowreg[kFastMaxWords / 32] = {0, 0, 0, 0};
  {
    const int a0 = c0 >= A + 1 ? c0 - (A + 1) : c0;
    if (c0 < n_classes && a0 < A) {
      const uint32_t* row = s.bitmap + (size_t)a0 * W;
#pragma unroll
      for (int i = 0; i < kFastMaxWords / 32; ++i)
        if (lane + 32 * i < W) rowreg[i] = __ldg(row + lane + 32 * i);
    }
  }
  LIG_STAMP(0);
  const Fields f = stage_fields(s, smem + fast_fixed_bytes(W));
  __syncthreads();
  LIG_STAMP(1);

  // ---- this thread's pods ----
  int q[kOwn];
  double kv[kOwn];
  uint32_t valid = 0, room = 0, low = 0, shed = 0;
#pragma unroll
  for (int j = 0; j < kOwn; ++j) {
    const int w = warp + kBuildWarps * j;
    const int p = w * 32 + lane;
    q[j] = 0;
    kv[j] = 0.0;
    if (w < W && p < P) {
      q[j] = f.q[p];
      kv[j] = f.kv[p];
      valid |= 1u << j;
      if (f.na[p] < f.ma[p]) room |= 1u << j;                                       // filter.go:175-177
      if ((long long)q[j] < thr.q_lora) low |= 1u << j;                             // scheduler.go:58-60
      if ((long long)q[j] <= thr.q_crit && kv[j] <= thr.kv_thr) shed |= 1u << j;    // scheduler.go:74-79
    }
  }
  int rb = 0;   // which half of the reduction scratch the next block-wide stage uses
  // Block-wide stages.  The critical and the sheddable side of the tree are independent, so every
  // stage reduces BOTH at once: 5 barriers for the whole shared part of the tree.
  auto sum4 = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t* out) {
    a = __reduce_add_sync(kFull, a);
    b = __reduce_add_sync(kFull, b);
    c = __reduce_add_sync(kFull, c);
    d = __reduce_add_sync(kFull, d);
    if (lane == 0) { sh->red_u[rb][0][warp] = a; sh->red_u[rb][1][warp] = b; sh->red_u[rb][2][warp] = c; sh->red_u[rb][3][warp] = d; }
    __syncthreads();
    out[0] = out[1] = out[2] = out[3] = 0;
#pragma unroll
    for (int i = 0; i < kBuildWarps; ++i) {
      out[0] += sh->red_u[rb][0][i]; out[1] += sh->red_u[rb][1][i];
      out[2] += sh->red_u[rb][2][i]; out[3] += sh->red_u[rb][3][i];
    }
    rb ^= 1;
  };
  auto seal = [&](uint32_t fixed_entries) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(co.counters + 1, 1u) == gridDim.x - 1) {
        __threadfence();
        const uint32_t total = fixed_entries + *reinterpret_cast<volatile uint32_t*>(co.counters);
        chdr->n_classes = (uint32_t)n_classes;
        chdr->pool_entries = total;
        chdr->reserved = 0;
        chdr->bytes = total <= co.pool_capacity
                          ? (uint32_t)((sizeof(CompactHeader) + (size_t)n_classes * sizeof(ClassEntry) +
                                        (size_t)total * sizeof(uint16_t) + 15) & ~(size_t)15) : 0u;
        if (co.host_hdr) {
          co.host_hdr->n_classes = chdr->n_classes;
          co.host_hdr->pool_entries = total;
          co.host_hdr->reserved = 0;
          co.host_hdr->bytes = chdr->bytes;
        }
        co.counters[0] = 0;   // ready for the next build of this slot
        co.counters[1] = 0;
      }
    }
  };

  // ---- shared stages ----
  LIG_STAMP(2);
  uint32_t cnt[4];
  sum4(__popc(low), __popc(shed), __popc(low & room), 0u, cnt);                       // barrier 1
  const uint32_t n_low = cnt[0], n_shed = cnt[1], n_low_room = cnt[2];
  LIG_STAMP(3);
  // critical side, start set:  low-queue pods that can accept the adapter, else all low-queue pods
  // (scheduler.go:58-69); no low-queue pod at all: every pod (scheduler.go:71)
  uint32_t tc = n_low > 0 ? (n_low_room > 0 ? (low & room) : low) : valid;
  uint32_t n_tc = n_low > 0 ? (n_low_room > 0 ? n_low_room : n_low) : (uint32_t)P;
  // sheddable side, start set: pods with capacity (scheduler.go:74-79); empty -> drop
  uint32_t ts = shed;
  uint32_t n_ts = n_shed;
  // leastQueuingFilterFunc on both                                             filter.go:102-122
  {
    int mnc = 0x7fffffff, mxc = 0, mns = 0x7fffffff, mxs = 0;
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
      if ((tc >> j) & 1u) { mnc = min(mnc, q[j]); mxc = max(mxc, q[j]); }
      if ((ts >> j) & 1u) { mns = min(mns, q[j]); mxs = max(mxs, q[j]); }
    }
    mnc = __reduce_min_sync(kFull, mnc); mxc = __reduce_max_sync(kFull, mxc);
    mns = __reduce_min_sync(kFull, mns); mxs = __reduce_max_sync(kFull, mxs);
    if (lane == 0) { sh->red_i[rb][0][warp] = mnc; sh->red_i[rb][1][warp] = mxc; sh->red_i[rb][2][warp] = mns; sh->red_i[rb][3][warp] = mxs; }
    __syncthreads();                                                                  // barrier 2
#pragma unroll
    for (int i = 0; i < kBuildWarps; ++i) {
      mnc = min(mnc, sh->red_i[rb][0][i]); mxc = max(mxc, sh->red_i[rb][1][i]);
      mns = min(mns, sh->red_i[rb][2][i]); mxs = max(mxs, sh->red_i[rb][3][i]);
    }
    rb ^= 1;
    // min + (max - min) / len, Go int64 truncated division (see stage_least_queuing)
    const long long thc = (long long)mnc + (long long)(((uint32_t)mxc - (uint32_t)mnc) / (n_tc ? n_tc : 1u));
    const long long ths = (long long)mns + (long long)(((uint32_t)mxs - (uint32_t)mns) / (n_ts ? n_ts : 1u));
    uint32_t kc = 0, ks = 0;
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
      if (((tc >> j) & 1u) && (long long)q[j] >= (long long)mnc && (long long)q[j] <= thc) kc |= 1u << j;
      if (((ts >> j) & 1u) && (long long)q[j] >= (long long)mns && (long long)q[j] <= ths) ks |= 1u << j;
    }
    tc = n_tc ? kc : 0u;
    ts = n_ts ? ks : 0u;
  }
  sum4(__popc(tc), __popc(ts), __popc(tc & room), __popc(ts & room), cnt);            // barrier 3
  n_tc = cnt[0];
  n_ts = cnt[1];
  // "low cost LoRA" on the least-queuing set: (affinity | room); for the default lists (no
  // affinity) that is the room part, falling back to the whole set           scheduler.go:35-46
  uint32_t cb_set, cz_set = 0, sb_set = 0, sz_set = 0;
  if (n_low > 0) {
    cb_set = low;                      // rows are ANDed with the low-queue set ("affinity LoRA")
  } else {
    cz_set = tc & room;
    cb_set = tc & ~room;
    if (cnt[2] > 0) { tc = cz_set; n_tc = cnt[2]; }
  }
  if (n_shed > 0) {
    sz_set = ts & room;
    sb_set = ts & ~room;
    if (cnt[3] > 0) { ts = sz_set; n_ts = cnt[3]; }
  }
  // leastKVCacheFilterFunc on both                                             filter.go:134-154
  {
    double mnc = 1.7976931348623157e308, mxc = 0.0, mns = 1.7976931348623157e308, mxs = 0.0;
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
      if ((tc >> j) & 1u) { if (kv[j] <= mnc) mnc = kv[j]; if (kv[j] >= mxc) mxc = kv[j]; }
      if ((ts >> j) & 1u) { if (kv[j] <= mns) mns = kv[j]; if (kv[j] >= mxs) mxs = kv[j]; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double o = __shfl_xor_sync(kFull, mnc, off); if (o < mnc) mnc = o;
      o = __shfl_xor_sync(kFull, mxc, off);        if (o > mxc) mxc = o;
      o = __shfl_xor_sync(kFull, mns, off);        if (o < mns) mns = o;
      o = __shfl_xor_sync(kFull, mxs, off);        if (o > mxs) mxs = o;
    }
    if (lane == 0) { sh->red_d[rb][0][warp] = mnc; sh->red_d[rb][1][warp] = mxc; sh->red_d[rb][2][warp] = mns; sh->red_d[rb][3][warp] = mxs; }
    __syncthreads();                                                                  // barrier 4
#pragma unroll
    for (int i = 0; i < kBuildWarps; ++i) {
      double v = sh->red_d[rb][0][i]; if (v < mnc) mnc = v;
      v = sh->red_d[rb][1][i];        if (v > mxc) mxc = v;
      v = sh->red_d[rb][2][i];        if (v < mns) mns = v;
      v = sh->red_d[rb][3][i];        if (v > mxs) mxs = v;
    }
    rb ^= 1;
    const double thc = __dadd_rn(mnc, __ddiv_rn(__dsub_rn(mxc, mnc), (double)(n_tc ? n_tc : 1u)));
    const double ths = __dadd_rn(mns, __ddiv_rn(__dsub_rn(mxs, mns), (double)(n_ts ? n_ts : 1u)));
    uint32_t kc = 0, ks = 0;
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
      if (((tc >> j) & 1u) && kv[j] >= mnc && kv[j] <= thc) kc |= 1u << j;
      if (((ts >> j) & 1u) && kv[j] >= mns && kv[j] <= ths) ks |= 1u << j;
    }
    tc = n_tc ? kc : 0u;
    ts = n_ts ? ks : 0u;
  }
  sum4(__popc(tc), __popc(ts), 0u, 0u, cnt);                                          // barrier 5
  const uint32_t rc_n = cnt[0], rs_n = cnt[1];
  const uint32_t rc_status = rc_n ? LIG_OK : LIG_EMPTY;
  const uint32_t rs_status = n_shed == 0 ? (uint32_t)LIG_DROP : (rs_n ? (uint32_t)LIG_OK : (uint32_t)LIG_EMPTY);
  LIG_STAMP(4);
  // sets -> mask words in shared memory (what the per-class part ANDs the bitmap rows with)
#pragma unroll
  for (int j = 0; j < kOwn; ++j) {
    const int w = warp + kBuildWarps * j;
    const uint32_t w_room = __ballot_sync(kFull, (room >> j) & 1u), w_cb = __ballot_sync(kFull, (cb_set >> j) & 1u),
                   w_cz = __ballot_sync(kFull, (cz_set >> j) & 1u), w_sb = __ballot_sync(kFull, (sb_set >> j) & 1u),
                   w_sz = __ballot_sync(kFull, (sz_set >> j) & 1u), w_tc = __ballot_sync(kFull, (tc >> j) & 1u),
                   w_ts = __ballot_sync(kFull, (ts >> j) & 1u);
    if (lane == 0 && w < W) {
      M_room[w] = w_room; CB[w] = w_cb; CZ[w] = w_cz; SB[w] = w_sb; SZ[w] = w_sz; TMPC[w] = w_tc; TMPS[w] = w_ts;
    }
  }
  __syncthreads();
  // the two default lists, by the last two warps of CTA 0 (its first warps hold the most popular adapters)
  if (blockIdx.x == 0 && warp == kBuildWarps - 1 && rc_n) compact_mask_to_two_lists(TMPC, W, lane, lists + rc_row, cpool + 1, true);
  if (blockIdx.x == 0 && warp == kBuildWarps - 2 && rs_n)
    compact_mask_to_two_lists(TMPS, W, lane, lists + rs_row, cpool + 1 + rc_n, true);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cpool[0] = 0xffffu;
    lists[none] = 0xffffu;
  }
  const uint32_t fixed_entries = 1u + rc_n + rs_n;
  LIG_STAMP(5);

  // ---- per class (one warp each) ----
  for (int c = c0; c < n_classes; c += gridDim.x * kBuildWarps) {
    const bool critical = c >= A + 1;
    const int a = critical ? c - (A + 1) : c;
    ClassEntry e;
    uint32_t coff = 0;
    if (!critical && n_shed == 0) {
      e = make_entry(0u, (uint32_t)LIG_DROP, none);                        // scheduler.go:83-89
    } else {
      const uint32_t* Bm = critical ? CB : SB;
      const uint32_t* Zm = critical ? (n_low > 0 ? nullptr : CZ) : SZ;
      uint32_t hit = 0;
      if (a < A) {
        if (c != c0) {   // later classes of this warp (A > 1183): fetch the row now
          const uint32_t* row = s.bitmap + (size_t)a * W;
#pragma unroll
          for (int i = 0; i < kFastMaxWords / 32; ++i) rowreg[i] = (lane + 32 * i < W) ? __ldg(row + lane + 32 * i) : 0u;
        }
#pragma unroll
        for (int i = 0; i < kFastMaxWords / 32; ++i) {
          const int w = lane + 32 * i;
          if (w < W) {
            const uint32_t t = Bm[w] & rowreg[i];
            X[w] = t;
            hit += __popc(t);
          }
        }
        hit = __reduce_add_sync(kFull, hit);
      }
      if (hit == 0) {
        e = critical ? make_entry(rc_n, rc_status, rc_n ? rc_row : none)
                     : make_entry(rs_n, rs_status, rs_n ? rs_row : none);
        coff = critical ? (rc_n ? 1u : 0u) : (rs_n ? 1u + rc_n : 0u);
      } else {
        uint32_t n = hit;
        if (Zm) {   // low cost LoRA: (affinity | room) on the least-queuing set     filter.go:163-166
          n = 0;
          for (int w = lane; w < W; w += 32) {
            const uint32_t t = X[w] | Zm[w];
            X[w] = t;
            n += __popc(t);
          }
          n = __reduce_add_sync(kFull, n);
        }
        __syncwarp();
        if (critical && n_low > 0)      // "affinity LoRA" succeeded -> queueAndKVCacheFilter
          n = sparse_least_queuing<true>(f, X, W, lane, n);
        n = sparse_least_kv<true>(f, X, W, lane, n);
        const uint32_t off = (uint32_t)c * (uint32_t)list_stride;
        if (n) {
          uint32_t at = 0;
          if (lane == 0) at = atomicAdd(co.counters, n);           // pool space for this list
          at = __shfl_sync(kFull, at, 0) + fixed_entries;
          compact_mask_to_two_lists(X, W, lane, lists + off, cpool + at, at + n <= co.pool_capacity);
          coff = at;
        }
        e = make_entry(n, n ? (uint32_t)LIG_OK : (uint32_t)LIG_EMPTY, n ? off : none);
      }
    }
    if (lane == 0) {
      *reinterpret_cast<uint4*>(cls + c) = make_uint4(e.info, e.magic, e.q_limit, e.list_off);
      centries[c] = make_uint4(e.info, e.magic, e.q_limit, coff);
    }
    __syncwarp();
  }
  LIG_STAMP(6);
  seal(fixed_entries);
  LIG_STAMP(7);
}

// ---- snapshot delta ---------------------------------------------------------------------------------
// One warp per dirty pod: overwrite its four column values, clear its bit in every adapter row,
// set it again in the rows of its new ActiveModels.  Different pods may share a bitmap word, hence
// the atomics; a pod appears at most once in a delta.
struct DeltaView {
  const int* pod_idx;
  const double* kv;
  const int* q;
  const uint16_t* n_active;
  const uint16_t* max_active;
  const int* adapter_offsets;   // n_dirty + 1
  const int* adapter_ids;
  int n_dirty;
};

__global__ void lig_apply_delta_kernel(DeltaView d, double* __restrict__ kv, int* __restrict__ q,
                                       uint16_t* __restrict__ n_active, uint16_t* __restrict__ max_active,
                                       uint32_t* __restrict__ bitmap, int P, int A, int W) {
  const int lane = threadIdx.x & 31;
  const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= d.n_dirty) return;
  const int p = d.pod_idx[i];
  if (p < 0 || p >= P) return;
  if (lane == 0) {
    kv[p] = d.kv[i];
    q[p] = d.q[i];
    n_active[p] = d.n_active[i];
    max_active[p] = d.max_active[i];
  }
  const uint32_t bit = 1u << (p & 31);
  uint32_t* col = bitmap + (p >> 5);
  for (int a = lane; a < A; a += 32)
    if (col[(size_t)a * W] & bit) atomicAnd(col + (size_t)a * W, ~bit);
  __syncwarp();
  for (int k = d.adapter_offsets[i] + lane; k < d.adapter_offsets[i + 1]; k += 32) {
    const int a = d.adapter_ids[k];
    if (a >= 0 && a < A) atomicOr(col + (size_t)a * W, bit);
  }
}

// ---- K2b: the streaming pick -----------------------------------------------------------------------
// kPerThread requests per thread, strided by the CTA size so that every warp-level load is 512
// contiguous bytes and every store 256.  All loads of a thread are issued before the first use.
constexpr int kPickThreads = 256;

__device__ __forceinline__ int2 pick_one(const int4 r, const uint4* __restrict__ cls,
                                         const uint16_t* __restrict__ lists, uint32_t A,
                                         uint64_t seed) {
  const uint32_t critical = (uint32_t)r.y & LIG_REQ_CRITICAL;
  const uint64_t key = ((uint64_t)(uint32_t)r.w << 32) | (uint32_t)r.z;
  const uint32_t a = min((uint32_t)r.x, A);          // ids outside [0, A) (negative too) -> A
  const uint32_t c = critical * (A + 1u) + a;
  const uint4 e = __ldg(cls + c);                    // {info, magic, q_limit, list_off}
  const uint32_t k = int31n_entry(seed ^ key, e);
  // n == 0: list_off is the pool's 0xffff sentinel -> -1
  const int pod = (int)(short)__ldg(lists + (e.w + k));
  return make_int2(pod, (int)(e.x & 0xffff0003u));   // {pod_idx, status | n_survivors << 16}
}

// The same pick against the compact tables in shared memory (entries + pool, see CompactHeader).
__device__ __forceinline__ int2 pick_one_smem(const int4 r, const uint4* tab, const uint16_t* pool,
                                              uint32_t A, uint64_t seed) {
  const uint32_t critical = (uint32_t)r.y & LIG_REQ_CRITICAL;
  const uint64_t key = ((uint64_t)(uint32_t)r.w << 32) | (uint32_t)r.z;
  const uint32_t a = min((uint32_t)r.x, A);
  const uint32_t c = critical * (A + 1u) + a;
  const uint4 e = tab[c];
  const uint32_t k = int31n_entry(seed ^ key, e);
  const int pod = (int)(short)pool[e.w + k];
  return make_int2(pod, (int)(e.x & 0xffff0003u));
}

// One CTA's share of one batch: kPerThread requests per thread.
template <int kPerThread>
__device__ __forceinline__ void pick_cta(const int4* __restrict__ reqs, int2* __restrict__ out, int R,
                                         int cta, const uint4* __restrict__ cls,
                                         const uint16_t* __restrict__ lists, int A, uint64_t seed) {
  constexpr int kPerCta = kPickThreads * kPerThread;
  const int first = cta * kPerCta;
  const int4* src = reqs + first + threadIdx.x;
  int2* dst = out + first + threadIdx.x;
  if constexpr (kPerThread > 4) {
    // Software-pipelined form: kPerThread = kIter x kWidth; the loads of iteration i+1 are in
    // flight while iteration i is being scheduled, so one batch needs only R / kPerThread
    // threads and two consecutive batches of a queue are co-resident on the SMs.
    constexpr int kWidth = kPerThread / 4;
    constexpr int kIter = 4;
    if (first + kPerCta <= R) {
      int4 cur[kWidth], nxt[kWidth];
#pragma unroll
      for (int j = 0; j < kWidth; ++j) cur[j] = ld_stream_int4(src + j * kPickThreads);
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        if (it + 1 < kIter) {
#pragma unroll
          for (int j = 0; j < kWidth; ++j)
            nxt[j] = ld_stream_int4(src + ((it + 1) * kWidth + j) * kPickThreads);
        }
#pragma unroll
        for (int j = 0; j < kWidth; ++j)
          st_stream_int2(dst + (it * kWidth + j) * kPickThreads,
                         pick_one(cur[j], cls, lists, (uint32_t)A, seed));
#pragma unroll
        for (int j = 0; j < kWidth; ++j) cur[j] = nxt[j];
      }
      return;
    }
#pragma unroll 1
    for (int j = 0; j < kPerThread; ++j) {   // ragged tail CTA
      const int i = first + threadIdx.x + j * kPickThreads;
      if (i < R)
        st_stream_int2(dst + j * kPickThreads,
                       pick_one(ld_stream_int4(src + j * kPickThreads), cls, lists,
                                (uint32_t)A, seed));
    }
    return;
  }
  int4 r[kPerThread > 4 ? 1 : kPerThread];
  if (first + kPerCta <= R) {              // full CTA: no per-request bounds checks
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) r[j] = ld_stream_int4(src + j * kPickThreads);
#pragma unroll
    for (int j = 0; j < kPerThread; ++j)
      st_stream_int2(dst + j * kPickThreads,
                     pick_one(r[j], cls, lists, (uint32_t)A, seed));
  } else {                                 // ragged tail CTA
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int i = first + threadIdx.x + j * kPickThreads;
      if (i < R)
        st_stream_int2(dst + j * kPickThreads,
                       pick_one(ld_stream_int4(src + j * kPickThreads), cls, lists,
                                (uint32_t)A, seed));
    }
  }
}

// One batch, one short-lived CTA per kPickThreads * kPerThread requests.  This is the kernel of the
// host-buffer path: `reqs` / `out` may be page-locked host memory, read and written over PCIe in
// place.  8 CTAs/SM (<= 32 registers).
template <int kPerThread>
__global__ void __launch_bounds__(kPickThreads, kPerThread <= 8 ? 8 : 6)
lig_pick_stream_kernel(const int4* __restrict__ reqs, int2* __restrict__ out, int R,
                       const uint4* __restrict__ cls, const uint16_t* __restrict__ lists,
                       int A, uint64_t seed) {
  pick_cta<kPerThread>(reqs, out, R, blockIdx.x, cls, lists, A, seed);
}

// A whole queue of batches in one launch: blockIdx.y selects the batch.  A batch of a few thousand
// requests is far below one kernel launch's worth of work (C2: 1024 requests = one CTA), and even
// at 2^20 requests per batch one launch per batch only matches the merged launch on an otherwise
// idle host (see lig_ctx::merge_max_requests).
struct QueueItem {
  const int4* reqs;
  int2* out;
  uint64_t seed;
};

template <int kPerThread>
__global__ void __launch_bounds__(kPickThreads, kPerThread <= 8 ? 8 : 6)
lig_pick_queue_kernel(const QueueItem* __restrict__ items, int R, const uint4* __restrict__ cls,
                      const uint16_t* __restrict__ lists, int A) {
  const QueueItem* it = items + blockIdx.y;
  const int4* reqs = reinterpret_cast<const int4*>(__ldg(reinterpret_cast<const unsigned long long*>(&it->reqs)));
  int2* out = reinterpret_cast<int2*>(__ldg(reinterpret_cast<const unsigned long long*>(&it->out)));
  const uint64_t seed = __ldg(reinterpret_cast<const unsigned long long*>(&it->seed));
  pick_cta<kPerThread>(reqs, out, R, blockIdx.x, cls, lists, A, seed);
}

// ---- K2c: the persistent, TMA-pipelined pick (default for HBM-resident batches) -----------------
// One launch serves a whole queue of batches with a FIXED grid of resident CTAs.  Every CTA walks
// the queue's 1024-request tiles (CTA-local tile j = global tile blockIdx.x + j * gridDim.x):
//   * a producer warp streams each tile's 16 KB of descriptors global -> shared with one TMA bulk
//     copy (cp.async.bulk, SASS UBLKCP.S.G) into a kStages-deep ring, completion signalled on an
//     mbarrier (expect_tx), so HBM reads never wait for the compute of an earlier tile;
//   * kGroups consumer groups of 8 warps take the CTA's tiles in turn (group g: tiles g,
//     g + kGroups, ...).  A group waits for its tile, takes 4 descriptors per thread out of shared
//     memory (conflict-free LDS.128) and hands the stage straight back (mbarrier arrive): the ring
//     only holds bytes in flight, never bytes being worked on, so a few 16 KB stages feed many
//     warps.  kStages must be a multiple of kGroups: a group then always returns to the stages
//     whose previous phase it consumed itself, which is what makes the parity wait unambiguous.  Then class lookup + Int31n + list gather from the L1-resident tables;
//   * picks leave either as plain coalesced 8-byte stores or (kBulkStore) staged in shared memory
//     and written with one TMA bulk store per tile (UBLKCP.G.S), tracked by bulk async-groups.
// The CTAs live for the whole queue, so L1 keeps the class tables across tiles and across
// batches (L1 is invalidated at launch boundaries only) and there is no per-CTA launch/drain.
constexpr int kTile = 1024;                     // requests per tile: 16 KB in, 8 KB out
constexpr int kGroupThreads = 256;              // one consumer group: 8 warps x 4 requests per thread
constexpr int kMaxInlineItems = 96;             // queue items carried in kernel parameters

struct QueueParams {
  int n_batches;
  int R;                    // requests per batch
  int tiles_per_batch;      // ceil(R / kTile)
  int total_tiles;          // n_batches * tiles_per_batch (< 2^31, the host splits longer queues)
  const QueueItem* dev_items;             // used when n_batches > kMaxInlineItems
  QueueItem items[kMaxInlineItems];       // else: the per-batch pointers and seeds, inline
};

// Shared-memory budget of the compact tables (CompactHeader blob) inside the persistent kernels:
// 72 KB hold C4's 2050 class entries (32 KB), its 5.5 K list entries (11 KB) and a ~1100-model
// table (18 KB) of the model-request kernels.
constexpr uint32_t kTabBudget = 72 * 1024;

__host__ __device__ constexpr int persist_threads(int groups) { return groups * kGroupThreads + 32; }
__host__ __device__ constexpr size_t persist_smem_bytes(int groups, int stages, bool bulk_store, bool tab_smem) {
  return (tab_smem ? (size_t)kTabBudget : 0) + (size_t)stages * kTile * 16 +
         (bulk_store ? (size_t)groups * kTile * 8 : 0) + 2u * (size_t)stages * 8 + 64;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done;
  do {   // try_wait suspends the thread in hardware up to a time limit; loop until the phase flips
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
// TMA bulk copy global -> shared, completing `bytes` transaction bytes on `bar` (16 B granular).
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// TMA bulk copy shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void group_barrier(int group) {   // named barrier 1 + group: that group's 8 warps
  asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(kGroupThreads) : "memory");
}

__device__ __forceinline__ QueueItem queue_item(const QueueParams& qp, int b) {
  if (qp.n_batches <= kMaxInlineItems) return qp.items[b];
  QueueItem it;
  it.reqs = reinterpret_cast<const int4*>(__ldg(reinterpret_cast<const unsigned long long*>(&qp.dev_items[b].reqs)));
  it.out = reinterpret_cast<int2*>(__ldg(reinterpret_cast<const unsigned long long*>(&qp.dev_items[b].out)));
  it.seed = __ldg(reinterpret_cast<const unsigned long long*>(&qp.dev_items[b].seed));
  return it;
}

// kTabSmem: the compact tables (ctab, ctab_bytes <= kTabBudget) are pulled into shared memory by
// one TMA bulk copy at CTA start and every lookup is an LDS; otherwise lookups go to the strided
// tables (cls, lists) through L1.  The host picks the variant (see launch_persistent).
template <int kGroups, int kStages, bool kBulkStore, bool kTabSmem>
__global__ void __launch_bounds__(persist_threads(kGroups), (kTabSmem ? 2 : 6) / kGroups > 0 ? (kTabSmem ? 2 : 6) / kGroups : 1)
lig_pick_persistent_kernel(const __grid_constant__ QueueParams qp, const uint4* __restrict__ cls,
                           const uint16_t* __restrict__ lists, int A,
                           const unsigned char* __restrict__ ctab, uint32_t ctab_bytes) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* base = smem + (kTabSmem ? kTabBudget : 0u);
  int4* in_ring = reinterpret_cast<int4*>(base);                                   // [kStages][kTile]
  int2* out_stage = reinterpret_cast<int2*>(base + (size_t)kStages * kTile * 16);  // [kGroups][kTile] (kBulkStore)
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)kStages * kTile * 16 +
                                               (kBulkStore ? (size_t)kGroups * kTile * 8 : 0));
  uint64_t* empty = full + kStages;
  uint64_t* tab_bar = empty + kStages;
  const uint4* tab = reinterpret_cast<const uint4*>(smem + sizeof(CompactHeader));
  const uint16_t* pool = reinterpret_cast<const uint16_t*>(smem + sizeof(CompactHeader) +
                                                           (size_t)2 * ((size_t)A + 1) * sizeof(ClassEntry));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);                          // the producer's arrive.expect_tx
      mbar_init(empty + s, kGroupThreads / 32);        // one arrive per warp of the consuming group
    }
    mbar_init(tab_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tpb = qp.tiles_per_batch, total = qp.total_tiles, R = qp.R;

  if (warp == kGroups * (kGroupThreads / 32)) {
    // ---- producer: one lane feeds the ring, CTA-local tiles in order ----
    if (lane == 0) {
      if (kTabSmem) {                                  // the tables first: one bulk copy, L2-resident
        mbar_arrive_expect_tx(tab_bar, ctab_bytes);
        bulk_load(smem, ctab, ctab_bytes, tab_bar);
      }
      const int step = (int)gridDim.x;
      int b = (int)blockIdx.x / tpb;
      int t = (int)blockIdx.x - b * tpb;
      int stage = 0;
      uint32_t phase = 0;
      for (int g = (int)blockIdx.x; g < total; g += step) {
        mbar_wait(empty + stage, phase ^ 1u);          // free (passes at once on the first lap)
        const QueueItem it = queue_item(qp, b);
        const uint32_t n = (uint32_t)min(kTile, R - t * kTile);
        mbar_arrive_expect_tx(full + stage, n * 16u);
        bulk_load(in_ring + (size_t)stage * kTile, it.reqs + (size_t)t * kTile, n * 16u, full + stage);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
        t += step;
        while (t >= tpb) { t -= tpb; ++b; }
      }
    }
    return;
  }

  // ---- consumer group `grp`: CTA-local tiles grp, grp + kGroups, ... ----
  const int grp = warp / (kGroupThreads / 32);
  const int tid = (int)threadIdx.x - grp * kGroupThreads;
  const int step = (int)gridDim.x * kGroups;
  int g = (int)blockIdx.x + grp * (int)gridDim.x;
  int b = g / tpb;
  int t = g - b * tpb;
  int stage = grp % kStages;
  uint32_t phase = (uint32_t)(grp / kStages) & 1u;
  if (kTabSmem) mbar_wait(tab_bar, 0);                 // the tables have landed
  for (; g < total; g += step) {
    const QueueItem it = queue_item(qp, b);
    const int n = min(kTile, R - t * kTile);
    const int4* tile = in_ring + (size_t)stage * kTile + tid;
    int4 r[4];
    mbar_wait(full + stage, phase);                    // the tile's bytes have landed
    if (n == kTile) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = tile[j * kGroupThreads];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        r[j] = (tid + j * kGroupThreads < n) ? tile[j * kGroupThreads] : make_int4(0, 0, 0, 0);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);         // descriptors are in registers: stage is free
    // this group's next tile is kGroups ring positions further
    stage += kGroups;
    while (stage >= kStages) { stage -= kStages; phase ^= 1u; }
    int2 p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      p[j] = kTabSmem ? pick_one_smem(r[j], tab, pool, (uint32_t)A, it.seed)
                      : pick_one(r[j], cls, lists, (uint32_t)A, it.seed);
    int2* dst = it.out + (size_t)t * kTile + tid;
    if (kBulkStore && n == kTile) {
      int2* ob = out_stage + (size_t)grp * kTile;
      // the group's previous bulk store has finished READING the staging buffer
      if (tid == 0) bulk_store_wait_read_all();
      group_barrier(grp);
#pragma unroll
      for (int j = 0; j < 4; ++j) ob[tid + j * kGroupThreads] = p[j];
      fence_proxy_async_smem();                        // generic-proxy writes -> visible to the TMA
      group_barrier(grp);
      if (tid == 0) bulk_store(it.out + (size_t)t * kTile, ob, kTile * 8u);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (tid + j * kGroupThreads < n) st_stream_int2(dst + j * kGroupThreads, p[j]);
    }
    t += step;
    while (t >= tpb) { t -= tpb; ++b; }
  }
  if (kBulkStore && tid == 0) bulk_store_wait_read_all();   // shared memory must outlive the reads
}

// ---- K2d: persistent grid-stride pick with register prefetch (no shared memory) ------------------
// The LDG/STG counterpart of K2c, kept as the A/B reference of the TMA ring: resident CTAs walk
// a tile sequence of their own (2048-request tiles); the 4 descriptors of a thread's NEXT tile are requested (LDG.128) before
// the current tile is computed, so every warp always has one tile's worth of loads in flight.
constexpr int kLoopThreads = 512;   // tile = 2048 requests (4 per thread)
constexpr int kLoopTile = kLoopThreads * 4;

template <bool kTabSmem>
__global__ void __launch_bounds__(kLoopThreads, 3)
lig_pick_loop_kernel(const __grid_constant__ QueueParams qp, const uint4* __restrict__ cls,
                     const uint16_t* __restrict__ lists, int A,
                     const unsigned char* __restrict__ ctab, uint32_t ctab_bytes) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* tab_bar = reinterpret_cast<uint64_t*>(smem + kTabBudget);
  const uint4* tab = reinterpret_cast<const uint4*>(smem + sizeof(CompactHeader));
  const uint16_t* pool = reinterpret_cast<const uint16_t*>(smem + sizeof(CompactHeader) +
                                                           (size_t)2 * ((size_t)A + 1) * sizeof(ClassEntry));
  if (kTabSmem) {
    if (threadIdx.x == 0) {
      mbar_init(tab_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(tab_bar, ctab_bytes);
      bulk_load(smem, ctab, ctab_bytes, tab_bar);
    }
  }
  const int tpb = qp.tiles_per_batch, total = qp.total_tiles, R = qp.R;
  const int step = (int)gridDim.x;
  int g = (int)blockIdx.x;
  int b = g / tpb;
  int t = g - b * tpb;
  if (g >= total) return;
  QueueItem it = queue_item(qp, b);
  int n = min(kLoopTile, R - t * kLoopTile);
  int4 cur[4];
  {
    const int4* src = it.reqs + (size_t)t * kLoopTile + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      cur[j] = ((int)threadIdx.x + j * kLoopThreads < n) ? ld_stream_int4(src + j * kLoopThreads) : make_int4(0, 0, 0, 0);
  }
  if (kTabSmem) mbar_wait(tab_bar, 0);                 // the tables have landed
  for (;;) {
    const int g2 = g + step;
    int b2 = b, t2 = t + step;
    while (t2 >= tpb) { t2 -= tpb; ++b2; }
    QueueItem it2 = it;
    int n2 = 0;
    int4 nxt[4];
    if (g2 < total) {
      it2 = queue_item(qp, b2);
      n2 = min(kLoopTile, R - t2 * kLoopTile);
      const int4* src = it2.reqs + (size_t)t2 * kLoopTile + threadIdx.x;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        nxt[j] = ((int)threadIdx.x + j * kLoopThreads < n2) ? ld_stream_int4(src + j * kLoopThreads) : make_int4(0, 0, 0, 0);
    }
    int2* dst = it.out + (size_t)t * kLoopTile + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int2 p = kTabSmem ? pick_one_smem(cur[j], tab, pool, (uint32_t)A, it.seed)
                              : pick_one(cur[j], cls, lists, (uint32_t)A, it.seed);
      if ((int)threadIdx.x + j * kLoopThreads < n) st_stream_int2(dst + j * kLoopThreads, p);
    }
    if (g2 >= total) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
    g = g2; b = b2; t = t2; it = it2; n = n2;
  }
}

// ---- K2e: model requests (the step before Schedule, fused into the pick) --------------------------
// Replaces per request handlers/request.go:42-56: FetchModelData (backend/datastore.go:70-76),
// RandomWeightedDraw (datastore.go:78-98), IsCritical (datastore.go:100-105).  A request is one
// 32-bit model id, a result one 32-bit lig_mpick: 8 bytes per decision.
//
// Model blob (global; the persistent kernel appends it to the class tables in shared memory):
//   [ header 16 B ][ entries: n_models x 16 B ][ targets: 8 B each ]
//   entry.x  bits 0-7 n_targets | 8 critical | 9 present | 12-16 shift (Int31n magic of `total`)
//   entry.y  magic (see ClassEntry) of total = sum of weights, when n_targets >= 2
//   entry.z  total
//   entry.w  n_targets >= 2: index of the first target record; == 1: that target's adapter id;
//            == 0: the model's own adapter id (TargetModels empty, request.go:47)
//   target   {adapter_id, inclusive cumulative weight}: the reference's loop "if randomVal <
//            Weight return; randomVal -= Weight" (datastore.go:91-97) picks the first k with
//            randomVal < cum[k] when the weights are non-negative.
struct ModelHeader {
  uint32_t bytes;        // header + entries + targets, multiple of 16
  uint32_t n_models;
  uint32_t n_target_records;
  uint32_t reserved;
};
constexpr uint32_t kModelCritical = 1u << 8, kModelPresent = 1u << 9;

struct ModelTables {         // where a kernel reads the model blob from (shared or global memory)
  const uint4* entries;
  const uint2* targets;
  uint32_t n_models;
};

__device__ __forceinline__ uint32_t pack_mpick(int pod, uint32_t status, uint32_t target) {
  return ((uint32_t)pod & 0xffffu) | (status << 16) | (target << 24);
}

// The weighted draw of a model with >= 2 targets (entry e): randomVal = r.Int31n(total), then the
// first k with randomVal < cum[k]                                                datastore.go:81-97
template <bool kSmem>
__device__ __forceinline__ void draw_target(const uint4 e, const ModelTables& mt, uint64_t seed, uint64_t key,
                                            uint32_t* adapter, uint32_t* target) {
  const uint32_t nt = e.x & 0xffu;
  const uint32_t total = e.z, shift = (e.x >> 12) & 31u;
  const uint32_t q_limit = total > 1u ? 0x80000000u / total : 1u;
  uint64_t state = seed ^ key ^ LIG_DRAW_DOMAIN;
  uint32_t v = splitmix_int31(state);
  uint32_t q = __umulhi(v, e.y) >> shift;
  while (q >= q_limit) {
    v = splitmix_int31(state);
    q = __umulhi(v, e.y) >> shift;
  }
  const uint32_t r = total > 1u ? v - q * total : 0u;
  uint32_t k = 0;
  uint2 tr = kSmem ? mt.targets[e.w] : __ldg(mt.targets + e.w);
  while (k + 1u < nt && r >= tr.y) {   // first k with randomVal < cum[k]       datastore.go:91-97
    ++k;
    tr = kSmem ? mt.targets[e.w + k] : __ldg(mt.targets + e.w + k);
  }
  *adapter = tr.x;
  *target = k;
}

// resolve: model id -> (adapter, critical, target index); false = LIG_NO_MODEL.
template <bool kSmem>
__device__ __forceinline__ bool resolve_model(uint32_t m, const ModelTables& mt, uint64_t seed, uint64_t key,
                                              uint32_t* adapter, uint32_t* critical, uint32_t* target) {
  *adapter = 0xffffffffu; *critical = 0; *target = 255u;
  if (m >= mt.n_models) return false;
  const uint4 e = kSmem ? mt.entries[m] : __ldg(mt.entries + m);
  if (!(e.x & kModelPresent)) return false;
  *critical = (e.x >> 8) & 1u;
  const uint32_t nt = e.x & 0xffu;
  if (nt <= 1u) {                      // no TargetModels, or a single one: nothing to draw
    *adapter = e.w;
    *target = nt ? 0u : 255u;
    return true;
  }
  draw_target<kSmem>(e, mt, seed, key, adapter, target);
  return true;
}

// One model request end to end; the class lookup as in pick_one / pick_one_smem.
template <bool kSmem>
__device__ __forceinline__ uint32_t pick_model(uint32_t m, uint64_t key, const ModelTables& mt,
                                               const uint4* cls, const uint16_t* lists, uint32_t A,
                                               uint64_t seed) {
  uint32_t adapter, critical, target;
  if (!resolve_model<kSmem>(m, mt, seed, key, &adapter, &critical, &target))
    return pack_mpick(-1, (uint32_t)LIG_NO_MODEL, 255u);
  const int4 r = make_int4((int)adapter, (int)critical, (int)(uint32_t)key, (int)(uint32_t)(key >> 32));
  const int2 p = kSmem ? pick_one_smem(r, cls, lists, A, seed) : pick_one(r, cls, lists, A, seed);
  return pack_mpick(p.x, (uint32_t)p.y & 3u, target);
}

struct MQueueItem {
  const uint32_t* ids;
  uint32_t* out;
  uint64_t seed;
  uint64_t first_index;
};
constexpr int kMaxInlineMItems = 64;
struct MQueueParams {
  int n_batches;
  int R;
  int tiles_per_batch;
  int total_tiles;
  const MQueueItem* dev_items;
  MQueueItem items[kMaxInlineMItems];
};

__device__ __forceinline__ MQueueItem mqueue_item(const MQueueParams& qp, int b) {
  if (qp.n_batches <= kMaxInlineMItems) return qp.items[b];
  MQueueItem it;
  const unsigned long long* p = reinterpret_cast<const unsigned long long*>(qp.dev_items + b);
  it.ids = reinterpret_cast<const uint32_t*>(__ldg(p));
  it.out = reinterpret_cast<uint32_t*>(__ldg(p + 1));
  it.seed = __ldg(p + 2);
  it.first_index = __ldg(p + 3);
  return it;
}

__host__ __device__ constexpr size_t mpersist_smem_bytes(int stages, bool tab_smem) {
  return (tab_smem ? (size_t)kTabBudget : 0) + (size_t)stages * kTile * 4 + 2u * (size_t)stages * 8 + 64;
}

// Persistent kernel for HBM-resident model-id batches: same structure as K2c (TMA ring, consumer
// groups, tables in shared memory), 4 KB tiles, every thread takes 4 consecutive requests
// (one LDS.128 in, one 16-byte store out).
template <int kGroups, int kStages, bool kTabSmem>
__global__ void __launch_bounds__(persist_threads(kGroups), kGroups == 1 ? 3 : 2)
lig_pick_models_kernel(const __grid_constant__ MQueueParams qp, const uint4* __restrict__ cls,
                       const uint16_t* __restrict__ lists, int A,
                       const unsigned char* __restrict__ ctab, uint32_t ctab_bytes,
                       const unsigned char* __restrict__ mtab, uint32_t mtab_bytes, uint32_t n_models) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* base = smem + (kTabSmem ? kTabBudget : 0u);
  uint4* in_ring = reinterpret_cast<uint4*>(base);                                  // [kStages][kTile / 4]
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)kStages * kTile * 4);
  uint64_t* empty = full + kStages;
  uint64_t* tab_bar = empty + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kGroupThreads / 32);
    }
    mbar_init(tab_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tpb = qp.tiles_per_batch, total = qp.total_tiles, R = qp.R;

  if (warp == kGroups * (kGroupThreads / 32)) {
    if (lane == 0) {
      if (kTabSmem) {      // class tables, then the model table right behind them
        mbar_arrive_expect_tx(tab_bar, ctab_bytes + mtab_bytes);
        bulk_load(smem, ctab, ctab_bytes, tab_bar);
        bulk_load(smem + ctab_bytes, mtab, mtab_bytes, tab_bar);
      }
      const int step = (int)gridDim.x;
      int b = (int)blockIdx.x / tpb;
      int t = (int)blockIdx.x - b * tpb;
      int stage = 0;
      uint32_t phase = 0;
      for (int g = (int)blockIdx.x; g < total; g += step) {
        mbar_wait(empty + stage, phase ^ 1u);
        const MQueueItem it = mqueue_item(qp, b);
        const uint32_t n = (uint32_t)min(kTile, R - t * kTile);
        const uint32_t bytes = (n * 4u + 15u) & ~15u;   // the buffers are padded to 16 bytes by the ABI
        mbar_arrive_expect_tx(full + stage, bytes);
        bulk_load(in_ring + (size_t)stage * (kTile / 4), it.ids + (size_t)t * kTile, bytes, full + stage);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
        t += step;
        while (t >= tpb) { t -= tpb; ++b; }
      }
    }
    return;
  }

  const int grp = warp / (kGroupThreads / 32);
  const int tid = (int)threadIdx.x - grp * kGroupThreads;
  const int step = (int)gridDim.x * kGroups;
  int g = (int)blockIdx.x + grp * (int)gridDim.x;
  int b = g / tpb;
  int t = g - b * tpb;
  int stage = grp % kStages;
  uint32_t phase = (uint32_t)(grp / kStages) & 1u;
  ModelTables mt;
  const uint4* tab = cls;
  const uint16_t* pool = lists;
  if (kTabSmem) {
    tab = reinterpret_cast<const uint4*>(smem + sizeof(CompactHeader));
    pool = reinterpret_cast<const uint16_t*>(smem + sizeof(CompactHeader) + (size_t)2 * ((size_t)A + 1) * sizeof(ClassEntry));
    mt.entries = reinterpret_cast<const uint4*>(smem + ctab_bytes + sizeof(ModelHeader));
  } else {
    mt.entries = reinterpret_cast<const uint4*>(mtab + sizeof(ModelHeader));
  }
  mt.targets = reinterpret_cast<const uint2*>(mt.entries + n_models);
  mt.n_models = n_models;
  if (kTabSmem) mbar_wait(tab_bar, 0);
  for (; g < total; g += step) {
    const MQueueItem it = mqueue_item(qp, b);
    const int n = min(kTile, R - t * kTile);
    mbar_wait(full + stage, phase);
    const uint4 ids = in_ring[(size_t)stage * (kTile / 4) + tid];
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    stage += kGroups;
    while (stage >= kStages) { stage -= kStages; phase ^= 1u; }
    const int i0 = tid * 4;
    const uint64_t key0 = it.first_index + (uint64_t)t * kTile + (uint64_t)i0;
    // The thread's 4 requests in three passes.  (1) entry lookups; models with at most one target
    // are resolved on the spot.  (2) the weighted draws: few requests need one (only models that
    // split traffic), so instead of a divergent branch inside each of the 4 requests every lane
    // works off its OWN pending requests — the warp runs the draw max-over-lanes(pending) times
    // (1-2) rather than 4.  (3) the 4 picks.
    const uint32_t m[4] = {ids.x, ids.y, ids.z, ids.w};
    uint32_t adapter[4], tgt[4];
    uint32_t crit = 0, okm = 0, pending = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in_range = m[j] < mt.n_models;
      const uint4 e = in_range ? (kTabSmem ? mt.entries[m[j]] : __ldg(mt.entries + m[j])) : make_uint4(0, 0, 0, 0);
      const uint32_t nt = e.x & 0xffu;
      const bool ok = (e.x & kModelPresent) != 0;
      okm |= (ok ? 1u : 0u) << j;
      crit |= ((e.x >> 8) & 1u) << j;
      adapter[j] = e.w;
      tgt[j] = nt ? 0u : 255u;
      pending |= ((ok && nt >= 2u) ? 1u : 0u) << j;
    }
    while (pending) {
      const int j = __ffs((int)pending) - 1;
      pending &= pending - 1u;
      const uint32_t mj = j == 0 ? m[0] : j == 1 ? m[1] : j == 2 ? m[2] : m[3];
      const uint4 e = kTabSmem ? mt.entries[mj] : __ldg(mt.entries + mj);
      uint32_t a, k;
      draw_target<kTabSmem>(e, mt, it.seed, key0 + (uint64_t)j, &a, &k);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        if (jj == j) { adapter[jj] = a; tgt[jj] = k; }
    }
    uint32_t ov[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t key = key0 + (uint64_t)j;
      const int4 r = make_int4((int)adapter[j], (int)((crit >> j) & 1u), (int)(uint32_t)key, (int)(uint32_t)(key >> 32));
      const int2 p = kTabSmem ? pick_one_smem(r, tab, pool, (uint32_t)A, it.seed) : pick_one(r, tab, pool, (uint32_t)A, it.seed);
      ov[j] = ((okm >> j) & 1u) ? pack_mpick(p.x, (uint32_t)p.y & 3u, tgt[j]) : pack_mpick(-1, (uint32_t)LIG_NO_MODEL, 255u);
    }
    uint4 o = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    uint32_t* dst = it.out + (size_t)t * kTile + i0;
    if (i0 + 4 <= n) {
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};"
                   :: "l"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    } else {
      if (i0 + 0 < n) dst[0] = o.x;
      if (i0 + 1 < n) dst[1] = o.y;
      if (i0 + 2 < n) dst[2] = o.z;
    }
    t += step;
    while (t >= tpb) { t -= tpb; ++b; }
  }
}

// The plain form: one short-lived CTA per 1024 requests, LDG/STG only.  Host-buffer path (`ids` /
// `out` may be page-locked host memory), unaligned device buffers, and the resolve-only test hook
// (reqs_out != nullptr: also write the 16-byte descriptor built for every request).
__global__ void __launch_bounds__(kGroupThreads)
lig_pick_models_stream_kernel(const uint32_t* __restrict__ ids, uint32_t* __restrict__ out, int R,
                              const uint4* __restrict__ cls, const uint16_t* __restrict__ lists, int A,
                              const unsigned char* __restrict__ mtab, uint32_t n_models, uint64_t seed,
                              uint64_t first_index, int4* __restrict__ reqs_out) {
  ModelTables mt;
  mt.entries = reinterpret_cast<const uint4*>(mtab + sizeof(ModelHeader));
  mt.targets = reinterpret_cast<const uint2*>(mt.entries + n_models);
  mt.n_models = n_models;
  const int first = (int)blockIdx.x * kTile + (int)threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = first + j * kGroupThreads;
    if (i >= R) break;
    const uint32_t m = __ldg(ids + i);
    const uint64_t key = first_index + (uint64_t)i;
    if (reqs_out) {
      uint32_t adapter, critical, target;
      const bool ok = resolve_model<false>(m, mt, seed, key, &adapter, &critical, &target);
      reqs_out[i] = make_int4((int)adapter, (int)critical, (int)(uint32_t)key, (int)(uint32_t)(key >> 32));
      out[i] = pack_mpick(-1, ok ? (uint32_t)LIG_OK : (uint32_t)LIG_NO_MODEL, target);
    } else {
      out[i] = pick_model<false>(m, key, mt, cls, lists, (uint32_t)A, seed);
    }
  }
}

// ---- K2f: load feedback between the windows of a batch (opt-in) -----------------------------------
// hist[pod] += 1 for every pick of the window that chose a pod.
__global__ void lig_pick_hist_kernel(const int2* __restrict__ picks, int n, int* __restrict__ hist) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int pod = picks[i].x;
    if (pod >= 0) atomicAdd(hist + pod, 1);
  }
}
// WaitingQueueSize += picks of the window (all ranks); total += window; window = 0.
__global__ void lig_apply_feedback_kernel(int* __restrict__ q, int* __restrict__ hist_window,
                                          int* __restrict__ hist_total, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int h = hist_window[p];
  // saturating: a queue size beyond int32 is not representable in the device record
  const long long nq = (long long)q[p] + h;
  q[p] = nq > 0x7fffffffLL ? 0x7fffffff : (int)nq;
  if (hist_total) hist_total[p] += h;
  hist_window[p] = 0;
}

// ---- K3: persistent doorbell kernel (streaming micro-batches) ------------------------------------
// One resident CTA polls a mailbox in page-locked host memory.  The host writes the micro-batch
// (header + descriptors), then the ticket; the CTA schedules it with the same pick_one() and
// writes the picks and the completion ticket back to host memory.  No kernel launch and no stream
// synchronisation per micro-batch: the round trip is two PCIe crossings.  Class tables are read
// with ld.global.cg (L2 only): the kernel outlives snapshot uploads, and L1 is not coherent.
constexpr int kMailboxCapacity = 4096;   // requests per doorbell
constexpr uint32_t kMailboxQuit = 0xffffffffu;

struct alignas(64) Mailbox {
  // --- written by the host ---
  uint32_t ticket;            // doorbell: last field written (release); kMailboxQuit = leave the kernel
  uint32_t count;
  uint32_t A;
  uint32_t reserved0;
  uint64_t seed;
  const uint4* cls;           // device pointers of the snapshot slot to use
  const uint16_t* lists;
  uint32_t pad0[6];
  // --- written by the device ---
  alignas(64) uint32_t done;  // completion ticket (release)
  uint32_t pad1[15];
  alignas(64) lig_req reqs[kMailboxCapacity];
  alignas(64) lig_pick picks[kMailboxCapacity];
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int4 ld_sys_int4(const void* p) {
  int4 r;
  asm volatile("ld.relaxed.sys.global.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint4 ld_cg_uint4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_cg_u16(const uint16_t* p) {
  uint16_t r;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ int2 doorbell_pick(const int4 r, const uint4* cls, const uint16_t* lists,
                                              uint32_t A, uint64_t seed) {
  const uint32_t critical = (uint32_t)r.y & LIG_REQ_CRITICAL;
  const uint64_t key = ((uint64_t)(uint32_t)r.w << 32) | (uint32_t)r.z;
  const uint32_t a = min((uint32_t)r.x, A);
  const uint32_t c = critical * (A + 1u) + a;
  const uint4 e = ld_cg_uint4(cls + c);
  const uint32_t k = int31n_entry(seed ^ key, e);
  const int pod = (int)(short)ld_cg_u16(lists + (e.w + k));
  return make_int2(pod, (int)(e.x & 0xffff0003u));
}

__global__ void __launch_bounds__(kPickThreads)
lig_doorbell_kernel(Mailbox* mb) {
  __shared__ uint32_t s_go;
  uint32_t ticket = 1;
  for (;;) {
    if (threadIdx.x == 0) {
      uint32_t go = 1;
      for (;;) {
        const uint32_t t = ld_acquire_sys_u32(&mb->ticket);
        if (t == ticket) break;
        if (t == kMailboxQuit) { go = 0; break; }
      }
      s_go = go;
    }
    __syncthreads();
    if (!s_go) return;
    // One PCIe round trip for everything a small micro-batch needs: the header (same 64-byte
    // line as the ticket) and this thread's first descriptor are requested together, the
    // descriptor speculatively (the mailbox always holds kMailboxCapacity slots).
    const int4 h0 = ld_sys_int4(&mb->ticket);          // ticket, count, A, -
    const int4 h1 = ld_sys_int4(&mb->seed);            // seed lo/hi, cls lo/hi
    const int4 h2 = ld_sys_int4(&mb->lists);           // lists lo/hi, pad
    int4 r = ld_sys_int4(&mb->reqs[threadIdx.x]);
    const uint32_t count = (uint32_t)h0.y, A = (uint32_t)h0.z;
    const uint64_t seed = ((uint64_t)(uint32_t)h1.y << 32) | (uint32_t)h1.x;
    const uint4* cls = reinterpret_cast<const uint4*>(((uint64_t)(uint32_t)h1.w << 32) | (uint32_t)h1.z);
    const uint16_t* lists = reinterpret_cast<const uint16_t*>(((uint64_t)(uint32_t)h2.y << 32) | (uint32_t)h2.x);
    for (uint32_t i = threadIdx.x; i < count; i += kPickThreads) {
      if (i != threadIdx.x) r = ld_sys_int4(&mb->reqs[i]);
      const int2 p = doorbell_pick(r, cls, lists, A, seed);
      asm volatile("st.relaxed.sys.global.v2.s32 [%0], {%1, %2};"
                   :: "l"(reinterpret_cast<int2*>(&mb->picks[i])), "r"(p.x), "r"(p.y) : "memory");
    }
    // the barrier orders every thread's pick stores before thread 0's release store (cumulative)
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys_u32(&mb->done, ticket);
    ++ticket;
  }
}

// ---- K1: direct scan -------------------------------------------------------------------------------
template <bool kStaged>
__global__ void __launch_bounds__(kCtaThreads)
lig_scan_kernel(SnapView s, Thr thr, const int4* __restrict__ reqs, int2* __restrict__ out, int R,
                uint32_t* __restrict__ masks, uint64_t seed) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* X = reinterpret_cast<uint32_t*>(smem) + (size_t)warp * 3 * s.W;
  uint32_t* T = X + s.W;
  uint32_t* H = T + s.W;
  Fields f{s.kv, s.q, s.n_active, s.max_active};
  if constexpr (kStaged) {
    f = stage_fields(s, smem + scratch_bytes(s.W));
    __syncthreads();
  }
  for (int i = blockIdx.x * kWarpsPerCta + warp; i < R; i += gridDim.x * kWarpsPerCta) {
    int4 r = __ldg(reqs + i);  // whole warp reads the same 16 bytes (broadcast)
    const bool critical = ((uint32_t)r.y & LIG_REQ_CRITICAL) != 0;
    const uint64_t key = ((uint64_t)(uint32_t)r.w << 32) | (uint32_t)r.z;
    const uint32_t* Hrow = stage_adapter_row(s, r.x, H, lane);
    EvalResult e = tree_eval_warp<kStaged>(f, Hrow, critical, thr, s.P, s.W, X, T, lane);
    int pod = -1;
    if (e.n > 0) {
      uint32_t k = int31n(seed ^ key, e.n, max_accept_for(e.n));
      for (int w = 0; w < s.W; ++w) {  // k-th survivor in pod order = pods[k]   scheduler.go:121
        uint32_t word = e.mask[w];
        uint32_t c = __popc(word);
        if (k < c) { pod = w * 32 + (int)__fns(word, 0, (int)k + 1); break; }
        k -= c;
      }
    }
    if (masks) {
      uint32_t* dst = masks + (size_t)i * s.W;
      for (int w = lane; w < s.W; w += 32) dst[w] = e.n ? e.mask[w] : 0u;
    }
    if (lane == 0) out[i] = make_int2(pod, (int)(e.status | (e.n << 16)));
    __syncwarp();
  }
}

}  // namespace lig
