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

// One entry per request class c = critical * (A + 1) + min(adapter, A); 8 bytes (one uint2 load).
// Classes whose survivor set does not depend on the adapter share one of two default lists
// (rows 2(A+1) and 2(A+1)+1 of the list pool); the others own row c.  Row offsets are in list
// entries and fit 32 bits: (2 * 65535 + 2) rows x 32768 entries < 2^32.
//
//   info : bits 0-1 status | 2-3 list row selector | 4-8 shift | 9 n is a power of two | 16-31 n
//          (info & 0xffff0003 is the second word of lig_pick as is: status | n_survivors << 16)
//   magic: M = ceil(2^(32+shift) / n), shift = ceil(log2 n) - 1, so that for every v < 2^31
//          floor(v / n) == umulhi(v, M) >> shift   (Granlund-Montgomery, N = 31 bits; n >= 2)
// With q = floor(v / n), Go's Int31n is k = v - q * n, resampling while v > 2^31-1-(2^31 % n),
// i.e. while q >= floor(2^31 / n) = floor((2^31-1) / n) + (n is a power of two ? 1 : 0).
struct ClassEntry {
  uint32_t info;
  uint32_t magic;
};
enum : uint32_t { kRowOwn = 0, kRowCriticalDefault = 1, kRowSheddableDefault = 2 };

__host__ __device__ inline uint32_t entry_n(uint32_t info) { return info >> 16; }
__host__ __device__ inline uint32_t entry_status(uint32_t info) { return info & 3u; }
__host__ __device__ inline uint32_t class_list_row(uint32_t info, uint32_t c, uint32_t n_classes) {
  const uint32_t sel = (info >> 2) & 3u;
  return sel == kRowOwn ? c : n_classes + sel - 1u;
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

// The same draw with the class entry's precomputed magic (no integer division on the hot path).
__device__ __forceinline__ uint32_t int31n_magic(uint64_t state, uint32_t info, uint32_t magic) {
  const uint32_t n = info >> 16;
  const uint32_t shift = (info >> 4) & 31u;
  const uint32_t q_limit = (__umulhi(0x7fffffffu, magic) >> shift) + ((info >> 9) & 1u);
  uint32_t v = splitmix_int31(state);
  uint32_t q = __umulhi(v, magic) >> shift;
  while (q >= q_limit) {   // probability < n / 2^31 per draw
    v = splitmix_int31(state);
    q = __umulhi(v, magic) >> shift;
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

__device__ __forceinline__ ClassEntry make_entry(uint32_t n, uint32_t status, uint32_t row_sel) {
  ClassEntry e;
  // n <= 1: magic 0 gives q = 0 for every draw; the power-of-two bit makes the rejection limit 1,
  // so the first draw is always accepted and k = 0.
  uint32_t shift = 0, magic = 0, pow2 = 1;
  if (n >= 2) {
    const uint32_t l = 32u - (uint32_t)__clz(n - 1u);          // ceil(log2 n), 1..15
    shift = l - 1u;
    pow2 = (n & (n - 1u)) == 0u;
    // ceil(2^(32+shift) / n) < 2^32 because n > 2^(l-1)
    magic = (uint32_t)((((unsigned long long)1 << (32u + shift)) + n - 1u) / n);
  }
  e.info = status | (row_sel << 2) | (shift << 4) | (pow2 << 9) | (n << 16);
  e.magic = magic;
  return e;
}

template <bool kStaged>
__global__ void __launch_bounds__(kBuildThreads)
lig_class_build_kernel(SnapView s, Thr thr, ClassEntry* __restrict__ cls,
                       uint16_t* __restrict__ lists, int list_stride) {
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

  if (P == 0) {   // critical: predicate node errs on an empty pool -> sheddable branch -> drop
    for (int c = blockIdx.x * kBuildThreads + threadIdx.x; c < n_classes; c += gridDim.x * kBuildThreads)
      cls[c] = make_entry(0u, (uint32_t)LIG_DROP, kRowOwn);
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
    if (blockIdx.x == 0 && warp == 0 && rc_n) compact_mask_to_list(TMP, W, lane, lists + rc_row);
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
    if (blockIdx.x == 0 && warp == 0 && rc_n) compact_mask_to_list(TMP, W, lane, lists + rc_row);
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
    if (blockIdx.x == 0 && warp == 0 && rs_n) compact_mask_to_list(TMP, W, lane, lists + rs_row);
  }
  __syncthreads();

  // ---- per class (one warp each): AND the adapter row with the shared mask --------------------
  for (int c = blockIdx.x * kBuildWarps + warp; c < n_classes; c += gridDim.x * kBuildWarps) {
    const bool critical = c >= A + 1;
    const int a = critical ? c - (A + 1) : c;
    const uint32_t* row = a < A ? s.bitmap + (size_t)a * W : nullptr;
    ClassEntry e;
    if (!critical && n_shed == 0) {
      e = make_entry(0u, (uint32_t)LIG_DROP, kRowOwn);                     // scheduler.go:83-89
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
        e = critical ? make_entry(rc_n, rc_status, kRowCriticalDefault)
                     : make_entry(rs_n, rs_status, kRowSheddableDefault);
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
        if (n) compact_mask_to_list(X, W, lane, lists + off);
        e = make_entry(n, n ? (uint32_t)LIG_OK : (uint32_t)LIG_EMPTY, kRowOwn);
      }
    }
    if (lane == 0) cls[c] = e;
    __syncwarp();
  }
}

// ---- K2b: the streaming pick -----------------------------------------------------------------------
// kPerThread requests per thread, strided by the CTA size so that every warp-level load is 512
// contiguous bytes and every store 256.  All loads of a thread are issued before the first use.
constexpr int kPickThreads = 256;

__device__ __forceinline__ int2 pick_one(const int4 r, const uint2* __restrict__ cls,
                                         const uint16_t* __restrict__ lists, uint32_t list_stride,
                                         uint32_t A, uint64_t seed) {
  const uint32_t critical = (uint32_t)r.y & LIG_REQ_CRITICAL;
  const uint64_t key = ((uint64_t)(uint32_t)r.w << 32) | (uint32_t)r.z;
  const uint32_t a = min((uint32_t)r.x, A);          // ids outside [0, A) (negative too) -> A
  const uint32_t c = critical * (A + 1u) + a;
  const uint2 e = __ldg(cls + c);                    // {info, magic}
  int pod = -1;
  if (e.x >> 16) {
    const uint32_t k = int31n_magic(seed ^ key, e.x, e.y);
    const uint32_t row = class_list_row(e.x, c, 2u * (A + 1u));
    pod = (int)__ldg(lists + (row * list_stride + k));
  }
  return make_int2(pod, (int)(e.x & 0xffff0003u));   // {pod_idx, status | n_survivors << 16}
}

// One CTA's share of one batch: kPerThread requests per thread.
template <int kPerThread>
__device__ __forceinline__ void pick_cta(const int4* __restrict__ reqs, int2* __restrict__ out, int R,
                                         int cta, const uint2* __restrict__ cls,
                                         const uint16_t* __restrict__ lists, int list_stride, int A,
                                         uint64_t seed) {
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
                         pick_one(cur[j], cls, lists, (uint32_t)list_stride, (uint32_t)A, seed));
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
                                (uint32_t)list_stride, (uint32_t)A, seed));
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
                     pick_one(r[j], cls, lists, (uint32_t)list_stride, (uint32_t)A, seed));
  } else {                                 // ragged tail CTA
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int i = first + threadIdx.x + j * kPickThreads;
      if (i < R)
        st_stream_int2(dst + j * kPickThreads,
                       pick_one(ld_stream_int4(src + j * kPickThreads), cls, lists,
                                (uint32_t)list_stride, (uint32_t)A, seed));
    }
  }
}

// 8 CTAs/SM (<= 32 registers) so a 2^20-request batch (1024 CTAs) is a single wave on 148 SMs.
template <int kPerThread>
__global__ void __launch_bounds__(kPickThreads, kPerThread <= 8 ? 8 : 6)
lig_pick_stream_kernel(const int4* __restrict__ reqs, int2* __restrict__ out, int R,
                       const uint2* __restrict__ cls, const uint16_t* __restrict__ lists,
                       int list_stride, int A, uint64_t seed, const int4* __restrict__ prefetch,
                       const uint64_t* __restrict__ seed_cell) {
  constexpr int kPerCta = kPickThreads * kPerThread;
  // Inside a replayed CUDA graph the per-call seed lives in device memory (written by the
  // graph's root node); `seed` then only carries the batch's offset within the queue.
  if (seed_cell != nullptr) seed += __ldg(seed_cell);
  const int first = blockIdx.x * kPerCta;
  // Cross-kernel software pipeline: while this batch is being scheduled, pull the CTA's slice of
  // a LATER batch of the same queue from HBM into the 126 MB L2 with one TMA-family bulk
  // prefetch (every descriptor still crosses HBM exactly once; it just does so while this
  // kernel is busy with its class lookups and stores, so HBM never idles at kernel boundaries).
  if (prefetch != nullptr && threadIdx.x == 0) {
    const int n = min(kPerCta, R - first);
    if (n > 0) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                   :: "l"(prefetch + first), "r"(n * 16) : "memory");
    }
  }
  // Batches of one queue are independent: let the next batch's grid (launched with programmatic
  // stream serialization, see launch_pick) start as soon as SM slots free up.  A no-op for a
  // normally launched successor.
  asm volatile("griddepcontrol.launch_dependents;");
  pick_cta<kPerThread>(reqs, out, R, blockIdx.x, cls, lists, list_stride, A, seed);
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
lig_pick_queue_kernel(const QueueItem* __restrict__ items, int R, const uint2* __restrict__ cls,
                      const uint16_t* __restrict__ lists, int list_stride, int A) {
  const QueueItem* it = items + blockIdx.y;
  const int4* reqs = reinterpret_cast<const int4*>(__ldg(reinterpret_cast<const unsigned long long*>(&it->reqs)));
  int2* out = reinterpret_cast<int2*>(__ldg(reinterpret_cast<const unsigned long long*>(&it->out)));
  const uint64_t seed = __ldg(reinterpret_cast<const unsigned long long*>(&it->seed));
  pick_cta<kPerThread>(reqs, out, R, blockIdx.x, cls, lists, list_stride, A, seed);
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
  uint32_t list_stride;
  uint64_t seed;
  const uint2* cls;           // device pointers of the snapshot slot to use
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
__device__ __forceinline__ uint2 ld_cg_uint2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_cg_u16(const uint16_t* p) {
  uint16_t r;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ int2 doorbell_pick(const int4 r, const uint2* cls, const uint16_t* lists,
                                              uint32_t stride, uint32_t A, uint64_t seed) {
  const uint32_t critical = (uint32_t)r.y & LIG_REQ_CRITICAL;
  const uint64_t key = ((uint64_t)(uint32_t)r.w << 32) | (uint32_t)r.z;
  const uint32_t a = min((uint32_t)r.x, A);
  const uint32_t c = critical * (A + 1u) + a;
  const uint2 e = ld_cg_uint2(cls + c);
  int pod = -1;
  if (e.x >> 16) {
    const uint32_t k = int31n_magic(seed ^ key, e.x, e.y);
    const uint32_t row = class_list_row(e.x, c, 2u * (A + 1u));
    pod = (int)ld_cg_u16(lists + (row * stride + k));
  }
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
    const int4 h0 = ld_sys_int4(&mb->ticket);          // ticket, count, A, list_stride
    const int4 h1 = ld_sys_int4(&mb->seed);            // seed lo/hi, cls lo/hi
    const int4 h2 = ld_sys_int4(&mb->lists);           // lists lo/hi, pad
    int4 r = ld_sys_int4(&mb->reqs[threadIdx.x]);
    const uint32_t count = (uint32_t)h0.y, A = (uint32_t)h0.z, stride = (uint32_t)h0.w;
    const uint64_t seed = ((uint64_t)(uint32_t)h1.y << 32) | (uint32_t)h1.x;
    const uint2* cls = reinterpret_cast<const uint2*>(((uint64_t)(uint32_t)h1.w << 32) | (uint32_t)h1.z);
    const uint16_t* lists = reinterpret_cast<const uint16_t*>(((uint64_t)(uint32_t)h2.y << 32) | (uint32_t)h2.x);
    for (uint32_t i = threadIdx.x; i < count; i += kPickThreads) {
      if (i != threadIdx.x) r = ld_sys_int4(&mb->reqs[i]);
      const int2 p = doorbell_pick(r, cls, lists, stride, A, seed);
      asm volatile("st.relaxed.sys.global.v2.s32 [%0], {%1, %2};"
                   :: "l"(reinterpret_cast<int2*>(&mb->picks[i])), "r"(p.x), "r"(p.y) : "memory");
    }
    // the barrier orders every thread's pick stores before thread 0's release store (cumulative)
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys_u32(&mb->done, ticket);
    ++ticket;
  }
}

// Root node of a cached queue graph: publishes the call's seed to the graph's kernels.
__global__ void lig_set_seed_kernel(uint64_t* cell, uint64_t value) { *cell = value; }

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
