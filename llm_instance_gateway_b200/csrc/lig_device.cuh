// lig_device.cuh — sm_100a device code of the endpoint picker.
//
// Three kernels, all integer / FP64-compare work (no tensor cores; HBM + issue bound):
//   lig_class_build_kernel  one warp per request class (critical?, adapter): walks the reference's
//                           filter tree over all P pods and writes the class's survivor list.
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

// One entry per request class c = critical * (A + 1) + min(adapter, A); 8 bytes.
struct ClassEntry {
  uint32_t n_status;    // n_survivors | status << 16
  uint32_t max_accept;  // Int31n's rejection bound 2^31 - 1 - (2^31 % n); 0xffffffff if n is 2^k
};

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
__device__ __forceinline__ uint32_t int31n(uint64_t state, uint32_t n, uint32_t max_accept) {
  uint32_t v = splitmix_int31(state);
  if (max_accept == 0xffffffffu) return v & (n - 1);
  while (v > max_accept) v = splitmix_int31(state);
  return v % n;
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
// Class c = critical * (A + 1) + a, a in [0, A] (a == A: adapter active nowhere).  Writes
// cls[c] and the survivors, ascending pod index, to lists[c * list_stride ...].
template <bool kStaged>
__global__ void __launch_bounds__(kCtaThreads)
lig_class_build_kernel(SnapView s, Thr thr, ClassEntry* __restrict__ cls,
                       uint16_t* __restrict__ lists, int list_stride) {
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
  const int n_classes = 2 * (s.A + 1);
  for (int c = blockIdx.x * kWarpsPerCta + warp; c < n_classes; c += gridDim.x * kWarpsPerCta) {
    const bool critical = c >= s.A + 1;
    const int a = critical ? c - (s.A + 1) : c;
    const uint32_t* Hrow = stage_adapter_row(s, a, H, lane);
    EvalResult r = tree_eval_warp<kStaged>(f, Hrow, critical, thr, s.P, s.W, X, T, lane);
    uint16_t* list = lists + (size_t)c * list_stride;
    uint32_t base = 0;
    if (r.n > 0) {
      for (int w = 0; w < s.W; ++w) {
        uint32_t word = r.mask[w];
        if (word == 0) continue;
        if ((word >> lane) & 1u) {
          list[base + __popc(word & ((1u << lane) - 1u))] = (uint16_t)(w * 32 + lane);
        }
        base += __popc(word);
      }
    }
    if (lane == 0) {
      ClassEntry e;
      e.n_status = r.n | (r.status << 16);
      e.max_accept = r.n ? max_accept_for(r.n) : 0u;
      cls[c] = e;
    }
    __syncwarp();
  }
}

// ---- K2b: the streaming pick -----------------------------------------------------------------------
// kPerThread requests per thread, strided by the CTA size so that every warp-level load is 512
// contiguous bytes and every store 256.  All loads of a thread are issued before the first use.
constexpr int kPickThreads = 256;

template <int kPerThread>
__global__ void __launch_bounds__(kPickThreads)
lig_pick_stream_kernel(const int4* __restrict__ reqs, int2* __restrict__ out, int R,
                       const uint2* __restrict__ cls, const uint16_t* __restrict__ lists,
                       int list_stride, int A, uint64_t seed) {
  const int base = blockIdx.x * (kPickThreads * kPerThread) + threadIdx.x;
  int4 r[kPerThread];
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    int i = base + j * kPickThreads;
    if (i < R) r[j] = ld_stream_int4(reqs + i);
  }
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    int i = base + j * kPickThreads;
    if (i >= R) continue;
    const int adapter = r[j].x;
    const uint32_t critical = (uint32_t)r[j].y & LIG_REQ_CRITICAL;
    const uint64_t key = ((uint64_t)(uint32_t)r[j].w << 32) | (uint32_t)r[j].z;
    const int a = ((unsigned)adapter < (unsigned)A) ? adapter : A;
    const int c = (int)critical * (A + 1) + a;
    const uint2 e = __ldg(cls + c);
    const uint32_t n = e.x & 0xffffu;
    int pod = -1;
    if (n > 0) {
      uint32_t k = int31n(seed ^ key, n, e.y);
      pod = (int)__ldg(lists + (size_t)((uint32_t)c * (uint32_t)list_stride + k));
    }
    // lig_pick {int32 pod_idx; uint16 status; uint16 n_survivors}
    st_stream_int2(out + i, make_int2(pod, (int)((e.x >> 16) | (n << 16))));
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
