// lig.cu — the C ABI of include/lig.h over the sm_100a kernels of lig_device.cuh.
//
// Host-side responsibilities only: context and HBM/pinned allocation, snapshot slots (two resident
// epochs), stream/event ordering, the chunk-pipelined host-buffer path, error reporting.  There is
// deliberately no CPU implementation of the scheduling path in this library: if CUDA is missing
// every entry point fails with LIG_ERR_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "lig_device.cuh"

using namespace lig;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY_RC(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      *rc = fail(LIG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                 __FILE__, __LINE__);                                                       \
      return nullptr;                                                                       \
    }                                                                                       \
  } while (0)

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(LIG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                  __FILE__, __LINE__);                                                      \
  } while (0)

inline int words_for(int P) { return (P + 31) / 32; }

// Run STMT with the compile-time constant kPpt bound to the run-time requests-per-thread knob.
#define LIG_DISPATCH_PPT(ppt, STMT)                           \
  switch (ppt) {                                              \
    case 1:  { constexpr int kPpt = 1;  STMT; } break;        \
    case 2:  { constexpr int kPpt = 2;  STMT; } break;        \
    case 8:  { constexpr int kPpt = 8;  STMT; } break;        \
    case 16: { constexpr int kPpt = 16; STMT; } break;        \
    default: { constexpr int kPpt = 4;  STMT; } break;        \
  }

struct Layout {  // offsets into the packed blob
  size_t kv, q, na, ma, bitmap, total;
};

Layout layout_for(int P, int A) {
  const size_t Ppad = (size_t)words_for(P) * 32;
  Layout l;
  l.kv = 0;
  l.q = l.kv + Ppad * sizeof(double);
  l.na = l.q + Ppad * sizeof(int32_t);
  l.ma = l.na + Ppad * sizeof(uint16_t);
  l.bitmap = l.ma + Ppad * sizeof(uint16_t);
  l.total = l.bitmap + (size_t)A * words_for(P) * sizeof(uint32_t);
  l.total = (l.total + 15) & ~(size_t)15;
  if (l.total == 0) l.total = 16;
  return l;
}

struct Slot {
  bool valid = false;
  uint64_t epoch = 0;
  int P = 0, A = 0, W = 0;
  unsigned char* d_blob = nullptr;
  ClassEntry* d_cls = nullptr;
  uint16_t* d_lists = nullptr;
  unsigned char* h_blob = nullptr;  // pinned staging for host uploads
  cudaEvent_t ready = nullptr;      // class tables built
  cudaEvent_t idle = nullptr;       // last batch that read this slot
  uint64_t stamp = 0;               // upload order, to pick the slot to overwrite
};

constexpr int kPipeStreams = 8;
constexpr int kHostPipe = 3;
constexpr int kChunk = 1 << 16;  // requests per chunk of the host-buffer pipeline (1 MiB in)

}  // namespace

namespace {
struct QueueGraph {
  int slot_index = 0, P = 0, A = 0, R = 0, ns = 0, pd = 0, ppt = 0;
  std::vector<const lig_req*> reqs;
  std::vector<lig_pick*> outs;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t seed_node = nullptr;
  uint64_t* d_seed = nullptr;
  uint64_t last_use = 0;
};
constexpr size_t kMaxQueueGraphs = 24;
}  // namespace

struct lig_ctx {
  int device = 0;
  int max_pods = 0, max_adapters = 0, max_batch = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  lig_thresholds thr{0.8, 5, 50};
  std::mutex mu;                             // guards everything below
  std::atomic<uint64_t> launches{0};

  // ---- snapshots: two resident epochs ----
  Slot slot[2];
  uint64_t stamp = 0;
  cudaStream_t s_up = nullptr;               // snapshot uploads + table builds

  // ---- host-buffer batches ----
  cudaStream_t s_pipe[kPipeStreams] = {};    // also the queue streams of lig_schedule_batches_device
  lig_req* d_reqs = nullptr;                 // device staging (scan test hook)
  lig_pick* d_out = nullptr;
  uint32_t* d_masks = nullptr;               // scan test hook staging (lazily sized)
  size_t d_masks_bytes = 0;
  lig_req* h_reqs = nullptr;                 // pinned, device-mapped bounce buffers for pageable callers
  lig_pick* h_out = nullptr;

  // ---- batch queues (lig_schedule_batches_device) ----
  cudaEvent_t fork = nullptr;
  cudaEvent_t join[kPipeStreams] = {};
  // cached CUDA graphs: a queue with the same buffers, shape and snapshot slot is replayed with
  // one cudaGraphLaunch instead of one cudaLaunchKernel per batch
  std::vector<QueueGraph*> graphs;
  uint64_t graph_clock = 0;
  // a queue normally runs as ONE launch with blockIdx.y = batch; the item table goes through a
  // small pinned ring
  static constexpr int kItemSlots = 4;
  static constexpr int kMaxItems = 65535;
  QueueItem* d_items[kItemSlots] = {};
  QueueItem* h_items[kItemSlots] = {};
  cudaEvent_t items_free[kItemSlots] = {};
  int item_slot = 0;

  // ---- streaming doorbell (lig_stream_open): a persistent kernel polling a host mailbox ----
  Mailbox* mailbox = nullptr;       // pinned, device-mapped
  Mailbox* d_mailbox = nullptr;     // its device alias
  cudaStream_t s_doorbell = nullptr;
  uint32_t next_ticket = 1;
  bool stream_open = false;

  // ---- tuning knobs, read from the environment at lig_create ----
  int pick_per_thread = 4;          // LIG_PICK_PER_THREAD = 1|2|4 (plain) | 8|16 (software-pipelined)
  int queue_streams = 4;            // LIG_QUEUE_STREAMS   = 1..8 streams a queue is forked over
  bool use_pdl = false;             // LIG_PDL=1           programmatic dependent launch inside a queue
  int prefetch_distance = 1;        // LIG_PREFETCH=d      batch b pulls batch b+d of the queue into L2 (0 = off)
  bool use_graph = true;            // LIG_GRAPH=0         disable cached graph replay of queues
  int graph_min_batches = 4;
  // LIG_MERGE_MAX: largest R whose queues run as ONE merged launch (blockIdx.y = batch).  Default:
  // always.  At R = 2^20 a merged queue and a graph of per-batch kernels are equally fast on one
  // GPU (4.39 vs 4.43 us per batch), but with several GPUs busy on one host the per-kernel dispatch
  // overhead grows (5.9 us at 8 GPUs) and only the merged launch is immune to it.
  int merge_max_requests = 0x7fffffff;
};

namespace {

SnapView view_of(const Slot& s) {
  const Layout l = layout_for(s.P, s.A);
  SnapView v;
  v.kv = reinterpret_cast<const double*>(s.d_blob + l.kv);
  v.q = reinterpret_cast<const int*>(s.d_blob + l.q);
  v.n_active = reinterpret_cast<const uint16_t*>(s.d_blob + l.na);
  v.max_active = reinterpret_cast<const uint16_t*>(s.d_blob + l.ma);
  v.bitmap = reinterpret_cast<const uint32_t*>(s.d_blob + l.bitmap);
  v.P = s.P;
  v.A = s.A;
  v.W = s.W;
  return v;
}

Thr thr_of(const lig_ctx* c) {
  return Thr{c->thr.kv_cache_threshold, (long long)c->thr.queue_threshold_critical,
             (long long)c->thr.queueing_threshold_lora};
}

Slot* find_slot(lig_ctx* c, uint64_t epoch) {
  for (auto& s : c->slot)
    if (s.valid && s.epoch == epoch) return &s;
  return nullptr;
}

int resolve_slot(lig_ctx* c, uint64_t epoch, Slot** out) {
  if (!c->slot[0].valid && !c->slot[1].valid)
    return fail(LIG_ERR_NO_SNAPSHOT, "no snapshot uploaded yet");
  Slot* s = find_slot(c, epoch);
  if (!s)
    return fail(LIG_ERR_STALE_EPOCH, "epoch %llu is not resident (resident: %llu%s, %llu%s)",
                (unsigned long long)epoch, (unsigned long long)c->slot[0].epoch,
                c->slot[0].valid ? "" : " [empty]", (unsigned long long)c->slot[1].epoch,
                c->slot[1].valid ? "" : " [empty]");
  *out = s;
  return 0;
}

// Does the tree-walking kernel stage the pod columns in shared memory for this W?
bool staged_fits(const lig_ctx* c, int W) {
  return scratch_bytes(W) + staged_bytes(W) <= c->smem_optin;
}

// Enqueue the class-table build for slot s on `stream`.
int launch_class_build(lig_ctx* c, Slot& s, cudaStream_t stream) {
  const SnapView v = view_of(s);
  const int n_classes = 2 * (s.A + 1);
  const int W = s.W > 0 ? s.W : 1;
  const bool staged = s.P > 0 && build_fixed_bytes(W) + staged_bytes(W) <= c->smem_optin;
  const size_t smem = build_fixed_bytes(W) + (staged ? staged_bytes(W) : 0);
  // every CTA repeats the shared stages, so no more CTAs than needed to give each warp a couple
  // of classes, and never more than one per SM
  int grid = (n_classes + 2 * kBuildWarps - 1) / (2 * kBuildWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  if (grid < 1) grid = 1;
  const int stride = s.P > 0 ? s.P : 1;
  if (staged) {
    lig_class_build_kernel<true><<<grid, kBuildThreads, smem, stream>>>(v, thr_of(c), s.d_cls, s.d_lists, stride);
  } else {
    lig_class_build_kernel<false><<<grid, kBuildThreads, smem, stream>>>(v, thr_of(c), s.d_cls, s.d_lists, stride);
  }
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

template <int kPerThread>
cudaError_t launch_pick_variant(int grid, cudaStream_t stream, bool overlap_prev, const int4* in,
                                int2* out, int R, const uint2* cls, const uint16_t* lists,
                                int stride, int A, uint64_t seed, const int4* prefetch) {
  const uint64_t* no_cell = nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kPickThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap_prev ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, lig_pick_stream_kernel<kPerThread>, in, out, R, cls, lists, stride, A,
                            seed, prefetch, no_cell);
}

// overlap_prev: the previous operation on `stream` is a pick kernel of the same queue call, whose
// inputs and outputs this batch does not touch, so this grid may start before that one finished
// (programmatic dependent launch; completion order on the stream is unchanged).
int launch_pick(lig_ctx* c, const Slot& s, uint64_t seed, const lig_req* d_reqs, int R,
                lig_pick* d_out, cudaStream_t stream, bool overlap_prev = false,
                const lig_req* prefetch_reqs = nullptr) {
  if (R == 0) return 0;
  const int4* in = reinterpret_cast<const int4*>(d_reqs);
  int2* out = reinterpret_cast<int2*>(d_out);
  const uint2* cls = reinterpret_cast<const uint2*>(s.d_cls);
  const int stride = s.P > 0 ? s.P : 1;
  const int per_cta = kPickThreads * c->pick_per_thread;
  const int grid = (R + per_cta - 1) / per_cta;
  overlap_prev = overlap_prev && c->use_pdl;
  const int4* pf = reinterpret_cast<const int4*>(prefetch_reqs);
  cudaError_t e = cudaSuccess;
  LIG_DISPATCH_PPT(c->pick_per_thread,
                   e = launch_pick_variant<kPpt>(grid, stream, overlap_prev, in, out, R, cls, s.d_lists,
                                                 stride, s.A, seed, pf));
  CUDA_TRY(e);
  c->launches++;
  return 0;
}

int launch_scan(lig_ctx* c, const Slot& s, uint64_t seed, const lig_req* d_reqs, int R,
                lig_pick* d_out, uint32_t* d_masks, cudaStream_t stream) {
  if (R == 0) return 0;
  const SnapView v = view_of(s);
  const bool staged = s.P > 0 && staged_fits(c, s.W);
  const size_t smem = scratch_bytes(s.W > 0 ? s.W : 1) + (staged ? staged_bytes(s.W) : 0);
  const int ctas_per_sm = staged ? (smem > 110 * 1024 ? 1 : 2) : 4;
  const int ctas_needed = (R + kWarpsPerCta - 1) / kWarpsPerCta;
  int grid = ctas_needed < c->sm_count * ctas_per_sm ? ctas_needed : c->sm_count * ctas_per_sm;
  if (staged) {
    lig_scan_kernel<true><<<grid, kCtaThreads, smem, stream>>>(
        v, thr_of(c), reinterpret_cast<const int4*>(d_reqs), reinterpret_cast<int2*>(d_out), R,
        d_masks, seed);
  } else {
    lig_scan_kernel<false><<<grid, kCtaThreads, smem, stream>>>(
        v, thr_of(c), reinterpret_cast<const int4*>(d_reqs), reinterpret_cast<int2*>(d_out), R,
        d_masks, seed);
  }
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return 0;
}

// Pick the slot a new epoch overwrites: the one already holding that epoch, else an empty one,
// else the older of the two.
Slot& victim_slot(lig_ctx* c, uint64_t epoch) {
  if (Slot* s = find_slot(c, epoch)) return *s;
  if (!c->slot[0].valid) return c->slot[0];
  if (!c->slot[1].valid) return c->slot[1];
  return c->slot[0].stamp <= c->slot[1].stamp ? c->slot[0] : c->slot[1];
}

int check_shape(const lig_ctx* c, int P, int A) {
  if (P < 0 || P > c->max_pods)
    return fail(LIG_ERR_INVALID, "P=%d outside [0, max_pods=%d]", P, c->max_pods);
  if (A < 0 || A > c->max_adapters)
    return fail(LIG_ERR_INVALID, "A=%d outside [0, max_adapters=%d]", A, c->max_adapters);
  return 0;
}

void destroy_queue_graph(QueueGraph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  if (g->d_seed) cudaFree(g->d_seed);
  delete g;
}

template <int kPerThread>
void* pick_kernel_ptr() { return reinterpret_cast<void*>(&lig_pick_stream_kernel<kPerThread>); }

// Build the graph of one queue: a root node that publishes the seed, then the batches as
// kernel nodes in `ns` independent chains (batch b after batch b - ns), mirroring the forked
// streams of the ungraphed path; batch b prefetches batch b + pd.
int build_queue_graph(const Slot& s, QueueGraph* g, int n_batches) {
  CUDA_TRY(cudaMalloc(&g->d_seed, sizeof(uint64_t)));
  CUDA_TRY(cudaGraphCreate(&g->graph, 0));
  uint64_t seed0 = 0;
  void* seed_args[2] = {&g->d_seed, &seed0};
  cudaKernelNodeParams kp = {};
  kp.func = reinterpret_cast<void*>(&lig_set_seed_kernel);
  kp.gridDim = dim3(1);
  kp.blockDim = dim3(1);
  kp.kernelParams = seed_args;
  CUDA_TRY(cudaGraphAddKernelNode(&g->seed_node, g->graph, nullptr, 0, &kp));
  void* fn = nullptr;
  LIG_DISPATCH_PPT(g->ppt, fn = pick_kernel_ptr<kPpt>());
  const uint2* cls = reinterpret_cast<const uint2*>(s.d_cls);
  const uint16_t* lists = s.d_lists;
  int stride = s.P > 0 ? s.P : 1;
  int A = s.A, R = g->R;
  const int per_cta = kPickThreads * g->ppt;
  std::vector<cudaGraphNode_t> nodes((size_t)n_batches);
  for (int b = 0; b < n_batches; ++b) {
    const int4* in = reinterpret_cast<const int4*>(g->reqs[(size_t)b]);
    int2* out = reinterpret_cast<int2*>(g->outs[(size_t)b]);
    uint64_t offset = (uint64_t)b;
    const int4* pf = (g->pd > 0 && b + g->pd < n_batches)
                         ? reinterpret_cast<const int4*>(g->reqs[(size_t)(b + g->pd)]) : nullptr;
    const uint64_t* cell = g->d_seed;
    void* args[11] = {&in, &out, &R, &cls, &lists, &stride, &A, &offset, &pf, &cell, nullptr};
    cudaKernelNodeParams np = {};
    np.func = fn;
    np.gridDim = dim3((unsigned)((R + per_cta - 1) / per_cta));
    np.blockDim = dim3(kPickThreads);
    np.kernelParams = args;
    cudaGraphNode_t dep = b < g->ns ? g->seed_node : nodes[(size_t)(b - g->ns)];
    CUDA_TRY(cudaGraphAddKernelNode(&nodes[(size_t)b], g->graph, &dep, 1, &np));
  }
  CUDA_TRY(cudaGraphInstantiate(&g->exec, g->graph, 0));
  return 0;
}

// Find or build the cached graph of this queue; nullptr (and *rc = 0) when graphs do not apply.
QueueGraph* queue_graph_for(lig_ctx* c, const Slot& s, const lig_req* const* d_reqs, int R,
                            lig_pick* const* d_out, int n_batches, int ns, int pd, int* rc) {
  *rc = 0;
  if (!c->use_graph || n_batches < c->graph_min_batches || R == 0) return nullptr;
  const int slot_index = (int)(&s - c->slot);
  for (QueueGraph* g : c->graphs) {
    if (g->slot_index != slot_index || g->P != s.P || g->A != s.A || g->R != R || g->ns != ns ||
        g->pd != pd || g->ppt != c->pick_per_thread || (int)g->reqs.size() != n_batches)
      continue;
    if (memcmp(g->reqs.data(), d_reqs, (size_t)n_batches * sizeof(void*)) != 0 ||
        memcmp(g->outs.data(), d_out, (size_t)n_batches * sizeof(void*)) != 0)
      continue;
    g->last_use = ++c->graph_clock;
    return g;
  }
  if (c->graphs.size() >= kMaxQueueGraphs && c->stream_open)
    return nullptr;   // eviction needs a device-wide synchronise, impossible under a resident kernel
  if (c->graphs.size() >= kMaxQueueGraphs) {   // evict the least recently used
    size_t victim = 0;
    for (size_t i = 1; i < c->graphs.size(); ++i)
      if (c->graphs[i]->last_use < c->graphs[victim]->last_use) victim = i;
    CUDA_TRY_RC(cudaDeviceSynchronize());
    destroy_queue_graph(c->graphs[victim]);
    c->graphs.erase(c->graphs.begin() + (long)victim);
  }
  QueueGraph* g = new QueueGraph();
  g->slot_index = slot_index;
  g->P = s.P;
  g->A = s.A;
  g->R = R;
  g->ns = ns;
  g->pd = pd;
  g->ppt = c->pick_per_thread;
  g->reqs.assign(d_reqs, d_reqs + n_batches);
  g->outs.assign(d_out, d_out + n_batches);
  g->last_use = ++c->graph_clock;
  if ((*rc = build_queue_graph(s, g, n_batches)) != 0) {
    destroy_queue_graph(g);
    return nullptr;
  }
  c->graphs.push_back(g);
  return g;
}

// Ranges handed out by lig_host_alloc (and the ctx's own staging buffers): known to be pinned and
// device-mapped, so the per-call driver query below can be skipped for them.
struct PinnedRange { const char* base; size_t bytes; char* dev; };
std::mutex g_pinned_mu;
std::vector<PinnedRange> g_pinned;

void register_pinned(void* host, size_t bytes) {
  void* dev = nullptr;
  if (cudaHostGetDevicePointer(&dev, host, 0) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned.push_back(PinnedRange{static_cast<const char*>(host), bytes, static_cast<char*>(dev)});
}

void unregister_pinned(void* host) {
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  for (size_t i = 0; i < g_pinned.size(); ++i)
    if (g_pinned[i].base == host) {
      g_pinned.erase(g_pinned.begin() + (long)i);
      return;
    }
}

// Device-visible alias of a page-locked host pointer (identical under UVA), or nullptr when p is
// ordinary pageable memory.
template <typename T>
T* mapped_device_pointer(T* p) {
  {
    const char* q = reinterpret_cast<const char*>(p);
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    for (const PinnedRange& r : g_pinned)
      if (q >= r.base && q < r.base + r.bytes)
        return reinterpret_cast<T*>(r.dev + (q - r.base));
  }
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return reinterpret_cast<T*>(at.devicePointer);
}

}  // namespace

extern "C" {

const char* lig_last_error(void) { return g_err; }
const char* lig_version(void) { return "lig-b200 0.1 (sm_100a)"; }
int lig_abi_version(void) { return LIG_ABI_VERSION; }

int lig_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

size_t lig_snapshot_bytes(int P, int A) {
  if (P < 0 || A < 0) return 0;
  return layout_for(P, A).total;
}

int lig_pack_pods(int P, const int64_t* q, const int64_t* na, const int64_t* ma, int32_t* q_out,
                  uint16_t* na_out, uint16_t* ma_out) {
  if (P < 0 || (P > 0 && (!q || !na || !ma || !q_out || !na_out || !ma_out)))
    return fail(LIG_ERR_INVALID, "lig_pack_pods: null array");
  for (int i = 0; i < P; ++i) {
    if (q[i] < INT32_MIN || q[i] > INT32_MAX)
      return fail(LIG_ERR_RANGE, "pod %d: WaitingQueueSize %lld does not fit int32", i,
                  (long long)q[i]);
    if (na[i] < 0 || na[i] > LIG_MAX_ADAPTERS)
      return fail(LIG_ERR_RANGE, "pod %d: len(ActiveModels) %lld outside [0, %d]", i,
                  (long long)na[i], LIG_MAX_ADAPTERS);
    q_out[i] = (int32_t)q[i];
    na_out[i] = (uint16_t)na[i];
    ma_out[i] = (uint16_t)(ma[i] < 0 ? 0 : (ma[i] > 65535 ? 65535 : ma[i]));
  }
  return 0;
}

int lig_pack_snapshot(void* blob, int P, int A, const double* kv, const int32_t* q,
                      const uint16_t* na, const uint16_t* ma, const uint32_t* bitmap) {
  if (!blob || P < 0 || A < 0) return fail(LIG_ERR_INVALID, "lig_pack_snapshot: bad argument");
  if (P > 0 && (!kv || !q || !na || !ma)) return fail(LIG_ERR_INVALID, "null pod column");
  if (A > 0 && P > 0 && !bitmap) return fail(LIG_ERR_INVALID, "null bitmap");
  const Layout l = layout_for(P, A);
  unsigned char* b = static_cast<unsigned char*>(blob);
  if (P == 0) memset(b, 0, l.total);
  if (P > 0) {
    // copy the P real pods and zero only the padding pods of each column (no full-blob memset)
    const size_t pad = (size_t)words_for(P) * 32 - (size_t)P;
    memcpy(b + l.kv, kv, (size_t)P * sizeof(double));
    memset(b + l.kv + (size_t)P * sizeof(double), 0, pad * sizeof(double));
    memcpy(b + l.q, q, (size_t)P * sizeof(int32_t));
    memset(b + l.q + (size_t)P * sizeof(int32_t), 0, pad * sizeof(int32_t));
    memcpy(b + l.na, na, (size_t)P * sizeof(uint16_t));
    memset(b + l.na + (size_t)P * sizeof(uint16_t), 0, pad * sizeof(uint16_t));
    memcpy(b + l.ma, ma, (size_t)P * sizeof(uint16_t));
    memset(b + l.ma + (size_t)P * sizeof(uint16_t), 0, pad * sizeof(uint16_t));
    const size_t bm_bytes = (size_t)A * words_for(P) * sizeof(uint32_t);
    if (A > 0) memcpy(b + l.bitmap, bitmap, bm_bytes);
    memset(b + l.bitmap + bm_bytes, 0, l.total - l.bitmap - bm_bytes);
    // bits of padding pods must be clear in the last word of every row
    const int W = words_for(P), rem = P & 31;
    if (rem) {
      uint32_t* bm = reinterpret_cast<uint32_t*>(b + l.bitmap);
      const uint32_t keep = (1u << rem) - 1u;
      for (int a = 0; a < A; ++a) bm[(size_t)a * W + (W - 1)] &= keep;
    }
  }
  return 0;
}

static int create_impl(lig_ctx* c, int device, int max_pods, int max_adapters, int max_batch) {
  CUDA_TRY(cudaSetDevice(device));
  c->device = device;
  c->max_pods = max_pods;
  c->max_adapters = max_adapters;
  c->max_batch = max_batch;
  int v = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
  c->sm_count = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  c->smem_optin = (size_t)v;
  CUDA_TRY(cudaFuncSetAttribute(lig_class_build_kernel<true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, v));
  CUDA_TRY(cudaFuncSetAttribute(lig_class_build_kernel<false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, v));
  CUDA_TRY(cudaFuncSetAttribute(lig_scan_kernel<true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, v));
  CUDA_TRY(cudaFuncSetAttribute(lig_scan_kernel<false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, v));
  const size_t need = scratch_bytes(words_for(max_pods)) > build_fixed_bytes(words_for(max_pods))
                          ? scratch_bytes(words_for(max_pods)) : build_fixed_bytes(words_for(max_pods));
  if (need > c->smem_optin)
    return fail(LIG_ERR_INVALID, "max_pods=%d needs %zu B of per-CTA scratch, device allows %zu",
                max_pods, need, c->smem_optin);
  const Layout l = layout_for(max_pods, max_adapters);
  const size_t n_classes = 2 * ((size_t)max_adapters + 1);
  for (auto& s : c->slot) {
    CUDA_TRY(cudaMalloc(&s.d_blob, l.total));
    CUDA_TRY(cudaMalloc(&s.d_cls, n_classes * sizeof(ClassEntry)));
    CUDA_TRY(cudaMalloc(&s.d_lists, (n_classes + 2) * (size_t)max_pods * sizeof(uint16_t)));
    CUDA_TRY(cudaHostAlloc(&s.h_blob, l.total, cudaHostAllocDefault));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.idle, cudaEventDisableTiming));
  }
  if (const char* e = getenv("LIG_PICK_PER_THREAD")) {
    int v2 = atoi(e);
    if (v2 == 1 || v2 == 2 || v2 == 4 || v2 == 8 || v2 == 16) c->pick_per_thread = v2;
  }
  if (const char* e = getenv("LIG_PDL")) c->use_pdl = atoi(e) != 0;
  if (const char* e = getenv("LIG_GRAPH")) c->use_graph = atoi(e) != 0;
  if (const char* e = getenv("LIG_MERGE_MAX")) c->merge_max_requests = atoi(e);
  for (int i = 0; i < lig_ctx::kItemSlots; ++i) {
    CUDA_TRY(cudaMalloc(&c->d_items[i], sizeof(QueueItem) * lig_ctx::kMaxItems));
    CUDA_TRY(cudaHostAlloc(&c->h_items[i], sizeof(QueueItem) * lig_ctx::kMaxItems, cudaHostAllocDefault));
    CUDA_TRY(cudaEventCreateWithFlags(&c->items_free[i], cudaEventDisableTiming));
  }
  if (const char* e = getenv("LIG_PREFETCH")) {
    int v2 = atoi(e);
    if (v2 >= 0 && v2 <= 16) c->prefetch_distance = v2;
  }
  if (const char* e = getenv("LIG_QUEUE_STREAMS")) {
    int v2 = atoi(e);
    if (v2 >= 1 && v2 <= kPipeStreams) c->queue_streams = v2;
  }
  CUDA_TRY(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
  for (auto& ev : c->join) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
  for (auto& s : c->s_pipe) CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&c->d_reqs, (size_t)max_batch * sizeof(lig_req)));
  CUDA_TRY(cudaMalloc(&c->d_out, (size_t)max_batch * sizeof(lig_pick)));
  CUDA_TRY(cudaHostAlloc(&c->h_reqs, (size_t)max_batch * sizeof(lig_req), cudaHostAllocMapped));
  CUDA_TRY(cudaHostAlloc(&c->h_out, (size_t)max_batch * sizeof(lig_pick), cudaHostAllocMapped));
  register_pinned(c->h_reqs, (size_t)max_batch * sizeof(lig_req));
  register_pinned(c->h_out, (size_t)max_batch * sizeof(lig_pick));
  return 0;
}

int lig_create(lig_ctx** out, int device, int max_pods, int max_adapters, int max_batch) {
  if (!out) return fail(LIG_ERR_INVALID, "lig_create: out is null");
  *out = nullptr;
  if (max_pods < 1 || max_pods > LIG_MAX_PODS)
    return fail(LIG_ERR_INVALID, "max_pods=%d outside [1, %d]", max_pods, LIG_MAX_PODS);
  if (max_adapters < 0 || max_adapters > LIG_MAX_ADAPTERS)
    return fail(LIG_ERR_INVALID, "max_adapters=%d outside [0, %d]", max_adapters, LIG_MAX_ADAPTERS);
  if (max_batch < 1) return fail(LIG_ERR_INVALID, "max_batch=%d must be >= 1", max_batch);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(LIG_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev)
    return fail(LIG_ERR_INVALID, "device %d outside [0, %d)", device, ndev);
  lig_ctx* c = new lig_ctx();
  if (int rc = create_impl(c, device, max_pods, max_adapters, max_batch)) {
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    lig_destroy(c);
    memcpy(g_err, keep, sizeof(keep));
    return rc;
  }
  *out = c;
  return 0;
}

void lig_destroy(lig_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  lig_stream_close(c);   // a resident doorbell kernel would make the synchronize below wait forever
  cudaDeviceSynchronize();
  for (auto& s : c->slot) {
    cudaFree(s.d_blob);
    cudaFree(s.d_cls);
    cudaFree(s.d_lists);
    cudaFreeHost(s.h_blob);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.idle) cudaEventDestroy(s.idle);
  }
  if (c->mailbox) cudaFreeHost(c->mailbox);
  if (c->s_doorbell) cudaStreamDestroy(c->s_doorbell);
  for (int i = 0; i < lig_ctx::kItemSlots; ++i) {
    cudaFree(c->d_items[i]);
    cudaFreeHost(c->h_items[i]);
    if (c->items_free[i]) cudaEventDestroy(c->items_free[i]);
  }
  for (QueueGraph* g : c->graphs) destroy_queue_graph(g);
  c->graphs.clear();
  if (c->fork) cudaEventDestroy(c->fork);
  for (auto& ev : c->join)
    if (ev) cudaEventDestroy(ev);
  if (c->s_up) cudaStreamDestroy(c->s_up);
  for (auto& s : c->s_pipe)
    if (s) cudaStreamDestroy(s);
  cudaFree(c->d_reqs);
  cudaFree(c->d_out);
  cudaFree(c->d_masks);
  if (c->h_reqs) unregister_pinned(c->h_reqs);
  if (c->h_out) unregister_pinned(c->h_out);
  cudaFreeHost(c->h_reqs);
  cudaFreeHost(c->h_out);
  delete c;
}

int lig_set_thresholds(lig_ctx* c, const lig_thresholds* t) {
  if (!c || !t) return fail(LIG_ERR_INVALID, "lig_set_thresholds: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  c->thr = *t;
  // resident class tables were built with the old thresholds: rebuild them
  CUDA_TRY(cudaSetDevice(c->device));
  for (auto& s : c->slot) {
    if (!s.valid) continue;
    CUDA_TRY(cudaStreamWaitEvent(c->s_up, s.idle, 0));
    if (int rc = launch_class_build(c, s, c->s_up)) return rc;
    CUDA_TRY(cudaEventRecord(s.ready, c->s_up));
  }
  CUDA_TRY(cudaStreamSynchronize(c->s_up));
  return 0;
}

int lig_get_thresholds(const lig_ctx* c, lig_thresholds* t) {
  if (!c || !t) return fail(LIG_ERR_INVALID, "lig_get_thresholds: null argument");
  *t = c->thr;
  return 0;
}

int lig_upload_snapshot(lig_ctx* c, uint64_t epoch, int P, int A, const double* kv,
                        const int32_t* q, const uint16_t* na, const uint16_t* ma,
                        const uint32_t* bitmap) {
  if (!c) return fail(LIG_ERR_INVALID, "lig_upload_snapshot: ctx is null");
  if (int rc = check_shape(c, P, A)) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  Slot& s = victim_slot(c, epoch);
  // the pinned staging blob of this slot may still be in flight from its previous upload
  CUDA_TRY(cudaStreamSynchronize(c->s_up));
  if (int rc = lig_pack_snapshot(s.h_blob, P, A, kv, q, na, ma, bitmap)) return rc;
  CUDA_TRY(cudaStreamWaitEvent(c->s_up, s.idle, 0));  // batches still reading the old content
  s.valid = false;
  CUDA_TRY(cudaMemcpyAsync(s.d_blob, s.h_blob, layout_for(P, A).total, cudaMemcpyHostToDevice,
                           c->s_up));
  s.P = P;
  s.A = A;
  s.W = words_for(P);
  if (int rc = launch_class_build(c, s, c->s_up)) return rc;
  CUDA_TRY(cudaEventRecord(s.ready, c->s_up));
  CUDA_TRY(cudaStreamSynchronize(c->s_up));
  s.epoch = epoch;
  s.stamp = ++c->stamp;
  s.valid = true;
  return 0;
}

int lig_upload_snapshot_device(lig_ctx* c, uint64_t epoch, int P, int A, const void* d_blob,
                               void* stream) {
  if (!c || !d_blob) return fail(LIG_ERR_INVALID, "lig_upload_snapshot_device: null argument");
  if (int rc = check_shape(c, P, A)) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Slot& s = victim_slot(c, epoch);
  CUDA_TRY(cudaStreamWaitEvent(st, s.idle, 0));
  CUDA_TRY(cudaStreamWaitEvent(st, s.ready, 0));
  CUDA_TRY(cudaMemcpyAsync(s.d_blob, d_blob, layout_for(P, A).total, cudaMemcpyDeviceToDevice, st));
  s.P = P;
  s.A = A;
  s.W = words_for(P);
  if (int rc = launch_class_build(c, s, st)) return rc;
  CUDA_TRY(cudaEventRecord(s.ready, st));
  s.epoch = epoch;
  s.stamp = ++c->stamp;
  s.valid = true;
  return 0;
}

int lig_schedule_batch_device(lig_ctx* c, uint64_t epoch, uint64_t seed, const lig_req* d_reqs,
                              int R, lig_pick* d_out, void* stream) {
  if (!c || R < 0 || (R > 0 && (!d_reqs || !d_out)))
    return fail(LIG_ERR_INVALID, "lig_schedule_batch_device: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaStreamWaitEvent(st, s->ready, 0));
  if (int rc = launch_pick(c, *s, seed, d_reqs, R, d_out, st)) return rc;
  CUDA_TRY(cudaEventRecord(s->idle, st));
  return 0;
}

int lig_schedule_batches_device(lig_ctx* c, uint64_t epoch, uint64_t seed,
                                const lig_req* const* d_reqs, int R, lig_pick* const* d_out,
                                int n_batches, void* stream) {
  if (!c || R < 0 || n_batches < 0 || (n_batches > 0 && (!d_reqs || !d_out)))
    return fail(LIG_ERR_INVALID, "lig_schedule_batches_device: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  if (n_batches == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaStreamWaitEvent(st, s->ready, 0));
  for (int b = 0; b < n_batches; ++b)
    if (R > 0 && (!d_reqs[b] || !d_out[b]))
      return fail(LIG_ERR_INVALID, "lig_schedule_batches_device: null buffer in batch %d", b);
  // Independent batches may overlap on the device: fork the caller's stream into the ctx's
  // queue streams round-robin and join back, so the tail of batch b overlaps the head of b+1
  // while everything stays ordered with respect to `stream`.
  const int ns = (n_batches > 1) ? c->queue_streams : 1;
  // L2 prefetch of the next batch only pays while a few batches fit the 126 MB L2 together;
  // beyond that the prefetched lines are evicted before use and every descriptor is read twice
  const int pd = ((size_t)R * sizeof(lig_req) <= ((size_t)32 << 20)) ? c->prefetch_distance : 0;
  if (n_batches >= 2 && R > 0 && R <= c->merge_max_requests) {
    // small batches: one launch for up to 65535 of them
    const uint2* cls = reinterpret_cast<const uint2*>(s->d_cls);
    const int stride = s->P > 0 ? s->P : 1;
    const int per_cta = kPickThreads * c->pick_per_thread;
    for (int lo = 0; lo < n_batches; lo += lig_ctx::kMaxItems) {
      const int n = (n_batches - lo) < lig_ctx::kMaxItems ? (n_batches - lo) : lig_ctx::kMaxItems;
      const int slot = c->item_slot;
      c->item_slot = (slot + 1) % lig_ctx::kItemSlots;
      CUDA_TRY(cudaEventSynchronize(c->items_free[slot]));   // table of 4 queues ago has been copied
      for (int b = 0; b < n; ++b)
        c->h_items[slot][b] = QueueItem{reinterpret_cast<const int4*>(d_reqs[lo + b]),
                                        reinterpret_cast<int2*>(d_out[lo + b]), seed + (uint64_t)(lo + b)};
      CUDA_TRY(cudaMemcpyAsync(c->d_items[slot], c->h_items[slot], sizeof(QueueItem) * (size_t)n,
                               cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaEventRecord(c->items_free[slot], st));
      const dim3 grid((unsigned)((R + per_cta - 1) / per_cta), (unsigned)n);
      LIG_DISPATCH_PPT(c->pick_per_thread,
                       (lig_pick_queue_kernel<kPpt><<<grid, kPickThreads, 0, st>>>(
                           c->d_items[slot], R, cls, s->d_lists, stride, s->A)));
      CUDA_TRY(cudaGetLastError());
      c->launches++;
    }
    CUDA_TRY(cudaEventRecord(s->idle, st));
    return 0;
  }
  int grc = 0;
  if (QueueGraph* g = queue_graph_for(c, *s, d_reqs, R, d_out, n_batches, ns, pd, &grc)) {
    uint64_t seed_value = seed;
    void* seed_args[2] = {&g->d_seed, &seed_value};
    cudaKernelNodeParams kp = {};
    kp.func = reinterpret_cast<void*>(&lig_set_seed_kernel);
    kp.gridDim = dim3(1);
    kp.blockDim = dim3(1);
    kp.kernelParams = seed_args;
    CUDA_TRY(cudaGraphExecKernelNodeSetParams(g->exec, g->seed_node, &kp));
    CUDA_TRY(cudaGraphLaunch(g->exec, st));
    c->launches += (uint64_t)n_batches + 1;
    CUDA_TRY(cudaEventRecord(s->idle, st));
    return 0;
  }
  if (grc != 0) return grc;
  if (ns == 1) {
    for (int b = 0; b < n_batches; ++b)
      if (int rc = launch_pick(c, *s, seed + (uint64_t)b, d_reqs[b], R, d_out[b], st, b > 0,
                               (pd > 0 && b + pd < n_batches) ? d_reqs[b + pd] : nullptr))
        return rc;
  } else {
    CUDA_TRY(cudaEventRecord(c->fork, st));
    for (int k = 0; k < ns; ++k) CUDA_TRY(cudaStreamWaitEvent(c->s_pipe[k], c->fork, 0));
    for (int b = 0; b < n_batches; ++b)
      if (int rc = launch_pick(c, *s, seed + (uint64_t)b, d_reqs[b], R, d_out[b], c->s_pipe[b % ns],
                               b >= ns, (pd > 0 && b + pd < n_batches) ? d_reqs[b + pd] : nullptr))
        return rc;
    for (int k = 0; k < ns; ++k) {
      CUDA_TRY(cudaEventRecord(c->join[k], c->s_pipe[k]));
      CUDA_TRY(cudaStreamWaitEvent(st, c->join[k], 0));
    }
  }
  CUDA_TRY(cudaEventRecord(s->idle, st));
  return 0;
}

int lig_schedule_scan_device(lig_ctx* c, uint64_t epoch, uint64_t seed, const lig_req* d_reqs,
                             int R, lig_pick* d_out, uint32_t* d_masks, void* stream) {
  if (!c || R < 0 || (R > 0 && (!d_reqs || !d_out)))
    return fail(LIG_ERR_INVALID, "lig_schedule_scan_device: bad argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaStreamWaitEvent(st, s->ready, 0));
  if (int rc = launch_scan(c, *s, seed, d_reqs, R, d_out, d_masks, st)) return rc;
  CUDA_TRY(cudaEventRecord(s->idle, st));
  return 0;
}

int lig_schedule_batch(lig_ctx* c, uint64_t epoch, uint64_t seed, const lig_req* reqs, int R,
                       lig_pick* out) {
  if (!c || R < 0 || (R > 0 && (!reqs || !out)))
    return fail(LIG_ERR_INVALID, "lig_schedule_batch: bad argument");
  if (R > c->max_batch)
    return fail(LIG_ERR_INVALID, "R=%d exceeds max_batch=%d", R, c->max_batch);
  std::lock_guard<std::mutex> lk(c->mu);
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  if (R == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  // Host buffers are not staged through HBM: the pick kernel reads the descriptors from, and
  // writes the picks to, page-locked host memory directly over PCIe (16 B in / 8 B out per
  // request, both directions in flight at once, one launch, no copy-engine hop).  Measured on
  // this pool: 36 GB/s in + 18 GB/s out concurrently vs 22 GB/s for a cudaMemcpyAsync H2D.
  // Pinned caller buffers (lig_host_alloc, cudaHostAlloc, cudaHostRegister) are used in place;
  // pageable ones bounce through the ctx's pinned buffers chunk by chunk, the CPU copy of chunk
  // i+1 overlapping the kernel of chunk i.
  const lig_req* dev_in = mapped_device_pointer(reqs);
  lig_pick* dev_out = mapped_device_pointer(out);
  const int chunk = (dev_in && dev_out) ? R : kChunk;
  const int n_chunks = (R + chunk - 1) / chunk;
  const lig_req* stage_in = dev_in ? dev_in : mapped_device_pointer(c->h_reqs);
  lig_pick* stage_out = dev_out ? dev_out : mapped_device_pointer(c->h_out);
  if (!stage_in || !stage_out)
    return fail(LIG_ERR_CUDA, "pinned staging buffers are not device-mapped on this platform");
  for (int k = 0; k < n_chunks; ++k) {
    const int lo = k * chunk;
    const int n = (R - lo) < chunk ? (R - lo) : chunk;
    cudaStream_t st = c->s_pipe[k % kHostPipe];
    if (k < kHostPipe) CUDA_TRY(cudaStreamWaitEvent(st, s->ready, 0));
    if (!dev_in) memcpy(c->h_reqs + lo, reqs + lo, (size_t)n * sizeof(lig_req));
    if (int rc = launch_pick(c, *s, seed, stage_in + lo, n, stage_out + lo, st)) return rc;
  }
  for (int k = 0; k < kHostPipe && k < n_chunks; ++k) {
    CUDA_TRY(cudaEventRecord(s->idle, c->s_pipe[k]));  // last record wins; all are synced below
    CUDA_TRY(cudaStreamSynchronize(c->s_pipe[k]));
  }
  if (!dev_out) memcpy(out, c->h_out, (size_t)R * sizeof(lig_pick));
  return 0;
}

int lig_schedule_scan(lig_ctx* c, uint64_t epoch, uint64_t seed, const lig_req* reqs, int R,
                      lig_pick* out, uint32_t* masks) {
  if (!c || R < 0 || (R > 0 && (!reqs || !out)))
    return fail(LIG_ERR_INVALID, "lig_schedule_scan: bad argument");
  if (R > c->max_batch)
    return fail(LIG_ERR_INVALID, "R=%d exceeds max_batch=%d", R, c->max_batch);
  std::lock_guard<std::mutex> lk(c->mu);
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  if (R == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = c->s_pipe[0];
  const size_t mask_bytes = masks ? (size_t)R * (s->W > 0 ? s->W : 1) * sizeof(uint32_t) : 0;
  if (mask_bytes > c->d_masks_bytes) {
    if (c->stream_open)
      return fail(LIG_ERR_INVALID, "lig_schedule_scan: cannot grow the mask buffer while a stream is open "
                                   "(cudaFree would wait for the resident doorbell kernel)");
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(c->d_masks);
    c->d_masks = nullptr;
    c->d_masks_bytes = 0;
    CUDA_TRY(cudaMalloc(&c->d_masks, mask_bytes));
    c->d_masks_bytes = mask_bytes;
  }
  CUDA_TRY(cudaStreamWaitEvent(st, s->ready, 0));
  CUDA_TRY(cudaMemcpyAsync(c->d_reqs, reqs, (size_t)R * sizeof(lig_req), cudaMemcpyHostToDevice, st));
  if (int rc = launch_scan(c, *s, seed, c->d_reqs, R, c->d_out, masks ? c->d_masks : nullptr, st))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(out, c->d_out, (size_t)R * sizeof(lig_pick), cudaMemcpyDeviceToHost, st));
  if (masks && s->W > 0)
    CUDA_TRY(cudaMemcpyAsync(masks, c->d_masks, mask_bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(s->idle, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int lig_read_class(lig_ctx* c, uint64_t epoch, int critical, int adapter_id, int* status,
                   int* n_survivors, uint16_t* list) {
  if (!c || !status || !n_survivors) return fail(LIG_ERR_INVALID, "lig_read_class: null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaEventSynchronize(s->ready));
  const int a = (adapter_id >= 0 && adapter_id < s->A) ? adapter_id : s->A;
  const int cls = (critical ? 1 : 0) * (s->A + 1) + a;
  ClassEntry e;
  CUDA_TRY(cudaMemcpy(&e, s->d_cls + cls, sizeof(e), cudaMemcpyDeviceToHost));
  *n_survivors = (int)entry_n(e.info);
  *status = (int)entry_status(e.info);
  if (list && *n_survivors > 0)
    CUDA_TRY(cudaMemcpy(list, s->d_lists + (size_t)class_list_row(e.info, (uint32_t)cls, 2u * (uint32_t)(s->A + 1)) *
                                               (size_t)(s->P > 0 ? s->P : 1),
                        (size_t)*n_survivors * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  return 0;
}

void* lig_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(LIG_ERR_CUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  register_pinned(p, bytes);
  return p;
}

void lig_host_free(void* p) {
  if (!p) return;
  unregister_pinned(p);
  cudaFreeHost(p);
}

int lig_stream_capacity(void) { return kMailboxCapacity; }

int lig_stream_open(lig_ctx* c) {
  if (!c) return fail(LIG_ERR_INVALID, "lig_stream_open: ctx is null");
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->stream_open) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  if (!c->mailbox) {
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->mailbox), sizeof(Mailbox), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_mailbox), c->mailbox, 0));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->s_doorbell, cudaStreamNonBlocking));
  }
  memset(c->mailbox, 0, offsetof(Mailbox, reqs));
  c->next_ticket = 1;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  lig_doorbell_kernel<<<1, kPickThreads, 0, c->s_doorbell>>>(c->d_mailbox);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  c->stream_open = true;
  return 0;
}

int lig_stream_close(lig_ctx* c) {
  if (!c) return fail(LIG_ERR_INVALID, "lig_stream_close: ctx is null");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->stream_open) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  reinterpret_cast<std::atomic<uint32_t>*>(&c->mailbox->ticket)->store(kMailboxQuit, std::memory_order_release);
  CUDA_TRY(cudaStreamSynchronize(c->s_doorbell));
  c->stream_open = false;
  return 0;
}

int lig_stream_submit(lig_ctx* c, uint64_t epoch, uint64_t seed, const lig_req* reqs, int n,
                      lig_pick* out) {
  if (!c || n < 0 || (n > 0 && (!reqs || !out)))
    return fail(LIG_ERR_INVALID, "lig_stream_submit: bad argument");
  if (n > kMailboxCapacity)
    return fail(LIG_ERR_INVALID, "n=%d exceeds the doorbell capacity %d", n, kMailboxCapacity);
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->stream_open) return fail(LIG_ERR_INVALID, "lig_stream_submit: call lig_stream_open first");
  Slot* s = nullptr;
  if (int rc = resolve_slot(c, epoch, &s)) return rc;
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  // the tables of this slot must be complete before the resident kernel reads them; uploads
  // through the host API have already synchronised, device-side uploads are waited for here
  CUDA_TRY(cudaEventSynchronize(s->ready));
  Mailbox* mb = c->mailbox;
  memcpy(mb->reqs, reqs, (size_t)n * sizeof(lig_req));
  mb->count = (uint32_t)n;
  mb->A = (uint32_t)s->A;
  mb->list_stride = (uint32_t)(s->P > 0 ? s->P : 1);
  mb->seed = seed;
  mb->cls = reinterpret_cast<const uint2*>(s->d_cls);
  mb->lists = s->d_lists;
  const uint32_t ticket = c->next_ticket++;
  reinterpret_cast<std::atomic<uint32_t>*>(&mb->ticket)->store(ticket, std::memory_order_release);
  auto* done = reinterpret_cast<std::atomic<uint32_t>*>(&mb->done);
  uint64_t spins = 0;
  while (done->load(std::memory_order_acquire) != ticket) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((++spins & 0xfffff) == 0 && cudaStreamQuery(c->s_doorbell) != cudaErrorNotReady) {
      c->stream_open = false;   // the resident kernel is gone (error or device reset)
      return fail(LIG_ERR_CUDA, "doorbell kernel is not running: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  memcpy(out, mb->picks, (size_t)n * sizeof(lig_pick));
  return 0;
}

uint64_t lig_kernel_launches(const lig_ctx* c) { return c ? c->launches.load() : 0; }
int lig_sm_count(const lig_ctx* c) { return c ? c->sm_count : 0; }

}  // extern "C"
