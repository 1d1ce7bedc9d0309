/*
 * lig.h — C ABI of the B200-native endpoint picker (drop-in for the ext-proc scheduler hot path).
 *
 * This is the boundary a cgo shim binds.  Every entry point below replaces a
 * piece of the reference's Go scheduling package; the reference location is cited per symbol as
 * <file>:<line> relative to the reference repository root (kubernetes-sigs/llm-instance-gateway
 * @ 8e96339).  Plain pointers and sizes only: no C++ types, no torch types, no Go pointers kept
 * after a call returns.
 *
 * Conventions
 *   - every function returning int returns 0 on success or a negative LIG_ERR_* code; the
 *     human-readable reason is available (thread-local) from lig_last_error();
 *   - per-request outcomes are NOT errors: they are reported in lig_pick.status
 *     (LIG_OK / LIG_DROP / LIG_EMPTY), mirroring Schedule()'s (pod, err) return;
 *   - there is no CPU fallback anywhere behind this header: without a CUDA device lig_create
 *     fails with LIG_ERR_CUDA.
 */
#ifndef LIG_H_
#define LIG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIG_ABI_VERSION 2

/* ---- per-request status (lig_pick.status) ------------------------------------------------------
 * LIG_OK    Schedule returned a pod, nil                      pkg/ext-proc/scheduling/scheduler.go:120-121
 * LIG_DROP  the "drop request" leaf fired: the Go side must return
 *           status.Errorf(codes.ResourceExhausted, "dropping request due to limited backend
 *           resources") wrapped like scheduler.go:117         scheduler.go:83-89  -> 429 at handlers/server.go:97-109
 * LIG_EMPTY the tree returned ([], nil): "failed to apply filter, resulted 0 pods, this should
 *           never happen"                                      scheduler.go:116-118
 */
enum { LIG_OK = 0, LIG_DROP = 1, LIG_EMPTY = 2,
       /* only from the model entry points (lig_schedule_models_*): FetchModelData returned nil,
        * "error finding a model object in InferenceModel for input ..."  handlers/request.go:42-45 */
       LIG_NO_MODEL = 3 };

/* ---- batch-level error codes (negative return values) ---------------------------------------- */
enum {
  LIG_ERR_INVALID     = -1, /* bad argument (null pointer, P/A/R out of the ctx capacity, ...)    */
  LIG_ERR_CUDA        = -2, /* CUDA runtime/driver failure, or no device                          */
  LIG_ERR_STALE_EPOCH = -3, /* the epoch passed to schedule is not a resident snapshot            */
  LIG_ERR_NO_SNAPSHOT = -4, /* schedule called before any snapshot upload                         */
  LIG_ERR_RANGE       = -5, /* a host value does not fit the device record (see lig_pack_pods)    */
  LIG_ERR_BUSY        = -6, /* no free ticket for an asynchronous submit (wait for one first)     */
  LIG_ERR_NCCL        = -7  /* NCCL missing (dlopen libnccl.so.2 failed) or an NCCL call failed   */
};

/* ---- request descriptor: 16 bytes, one int4 load on the device --------------------------------
 * Replaces scheduling.LLMRequest (pkg/ext-proc/scheduling/types.go:4-11).  The filter tree reads
 * only ResolvedTargetModel and Critical (filter.go:163-181); Model and TargetModels are unused on
 * the path.  ResolvedTargetModel is interned by the host layer into adapter_id: the index of the
 * adapter's row in the snapshot bitmap, or any value outside [0, A) for a model that no pod
 * lists in ActiveModels (the base model, or a never-loaded adapter) — such a request matches no
 * pod in loRAAffinityPredicate, exactly like a Go map lookup miss (filter.go:170).
 * rand_key keys the request's private random stream used for the final pick (see below).
 */
typedef struct lig_req {
  int32_t  adapter_id;
  uint32_t flags;      /* bit 0 = Critical (types.go:10); other bits must be 0 */
  uint64_t rand_key;
} lig_req;
#define LIG_REQ_CRITICAL 1u

/* ---- result: 8 bytes ---------------------------------------------------------------------------
 * pod_idx      index into the snapshot's pod order (= order of the slice the injected
 *              PodMetricsProvider returned at pack time, scheduler.go:108-115), -1 unless LIG_OK;
 *              the host maps it back to backend.Pod{Name,Address} (backend/types.go:8-11).
 * n_survivors  len(pods) after the filter tree (the argument of rand.Intn, scheduler.go:120).
 */
typedef struct lig_pick {
  int32_t  pod_idx;
  uint16_t status;
  uint16_t n_survivors;
} lig_pick;

/* ---- thresholds: the compile-time constants of scheduler.go:15-24, made parameters so the
 * reference's own predicate test (filter_test.go:304-338 uses (0, 0.8)) can be run on the GPU. */
typedef struct lig_thresholds {
  double  kv_cache_threshold;         /* kvCacheThreshold        = 0.8 */
  int64_t queue_threshold_critical;   /* queueThresholdCritical  = 5   */
  int64_t queueing_threshold_lora;    /* queueingThresholdLoRA   = 50  */
} lig_thresholds;

typedef struct lig_ctx lig_ctx; /* opaque; one per (process, device) */

/* Hard limits of the device records. */
#define LIG_MAX_PODS     32768   /* list entries and n_survivors are uint16 */
#define LIG_MAX_ADAPTERS 65534

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Replaces scheduling.NewScheduler (scheduler.go:93-99): binds the scheduler to one CUDA device
 * and reserves HBM + pinned staging for snapshots of up to max_pods x max_adapters and host
 * batches of up to max_batch requests (device-pointer batches are not limited by max_batch). */
int  lig_create(lig_ctx** out, int device, int max_pods, int max_adapters, int max_batch);
void lig_destroy(lig_ctx* ctx);

int  lig_set_thresholds(lig_ctx* ctx, const lig_thresholds* t);   /* scheduler.go:15-24 */
int  lig_get_thresholds(const lig_ctx* ctx, lig_thresholds* t);

/* ---- snapshot = the value of PodMetricsProvider.AllPodMetrics() (scheduler.go:108-115,
 * backend/provider.go:38-46) frozen at one refresh tick, in device layout -----------------------
 *
 * Packed blob, W = ceil(P/32), Ppad = 32*W, every section 16-byte aligned, in this order:
 *   double   kv[Ppad]          Metrics.KVCacheUsagePercent           backend/types.go:24
 *   int32_t  q[Ppad]           Metrics.WaitingQueueSize              backend/types.go:23
 *   uint16_t n_active[Ppad]    len(Metrics.ActiveModels)             backend/types.go:19
 *   uint16_t max_active[Ppad]  Metrics.MaxActiveModels (saturated)   backend/types.go:21
 *   uint32_t bitmap[A][W]      bit (a, p) = adapter a in pod p's ActiveModels (adapter-major)
 * Padding pods (index >= P) are ignored by every kernel.
 */
size_t lig_snapshot_bytes(int P, int A);

/* Pure host helper (no GPU): narrow Go-width pod metrics to the device record.
 *   q          must fit int32, else LIG_ERR_RANGE (never silently wrapped);
 *   n_active   must be in [0, LIG_MAX_ADAPTERS], else LIG_ERR_RANGE;
 *   max_active is saturated to [0, 65535]: canAcceptNewLoraPredicate (filter.go:175-177) is
 *              `n_active < max_active` and n_active <= 65534, so saturation cannot change it. */
int lig_pack_pods(int P, const int64_t* waiting_queue_size, const int64_t* n_active_models,
                  const int64_t* max_active_models, int32_t* q_out, uint16_t* n_active_out,
                  uint16_t* max_active_out);

/* Pure host helper (no GPU): lay the five arrays out as the packed blob described above. */
int lig_pack_snapshot(void* blob, int P, int A, const double* kv, const int32_t* q,
                      const uint16_t* n_active, const uint16_t* max_active,
                      const uint32_t* bitmap_adapter_major /* A x ceil(P/32) */);

/* Upload a snapshot from host arrays and build the per-class survivor tables on the device.
 * After it returns, `epoch` is resident; the previous epoch stays resident too (two slots), so
 * in-flight batches against it still complete.  Replaces the per-request
 * AllPodMetrics() materialisation of scheduler.go:114-115 by one pack per refresh tick. */
int lig_upload_snapshot(lig_ctx* ctx, uint64_t epoch, int P, int A, const double* kv,
                        const int32_t* q, const uint16_t* n_active, const uint16_t* max_active,
                        const uint32_t* bitmap_adapter_major);
/* The same without waiting for the device: the arrays are copied into the ctx's pinned staging
 * before the call returns (the caller may reuse them at once), the H2D copy and the table build
 * are only enqueued.  `epoch` is resident immediately; batches against it are ordered behind the
 * build on the device, so the refresher goroutine (backend/provider.go:81-88) never blocks on the
 * GPU and a refresh + the next batch cost one synchronisation, not three.  A device-side failure
 * of the build surfaces at the next call that synchronises. */
int lig_upload_snapshot_async(lig_ctx* ctx, uint64_t epoch, int P, int A, const double* kv,
                              const int32_t* q, const uint16_t* n_active, const uint16_t* max_active,
                              const uint32_t* bitmap_adapter_major);

/* Delta form of lig_upload_snapshot for the refresh tick (backend/provider.go:134-179 re-scrapes
 * every pod, but between two ticks most pods report the same metrics): `new_epoch` becomes
 * `base_epoch` (which must be resident and stays so, untouched) with n_dirty pods replaced.  P
 * and A are those of the base.  For dirty pod i: pod_idx[i], its four column values, and its
 * COMPLETE ActiveModels as adapter ids adapter_ids[adapter_offsets[i] .. adapter_offsets[i+1]).
 * Only the delta crosses PCIe; the base blob is copied device-to-device, the dirty rows and
 * bitmap bits are patched by a kernel, and the class tables are rebuilt.  LIG_ERR_INVALID when
 * the base epoch is the slot the new epoch would have to overwrite (it is the older of the two
 * resident epochs): fall back to lig_upload_snapshot. */
int lig_update_snapshot(lig_ctx* ctx, uint64_t new_epoch, uint64_t base_epoch, int n_dirty,
                        const int32_t* pod_idx, const double* kv, const int32_t* q,
                        const uint16_t* n_active, const uint16_t* max_active,
                        const int32_t* adapter_offsets, const int32_t* adapter_ids);

/* Same, but the packed blob is already in HBM (e.g. the receive buffer of an NCCL broadcast of
 * the snapshot).  The blob is copied into the ctx on `stream` (a cudaStream_t, may be NULL =
 * legacy default stream); table build is enqueued on the same stream, nothing is synchronised. */
int lig_upload_snapshot_device(lig_ctx* ctx, uint64_t epoch, int P, int A, const void* d_blob,
                               void* stream);

/* ---- the hot path ------------------------------------------------------------------------------
 * Replaces R calls of Scheduler.Schedule (scheduler.go:113-122) == R walks of the defaultFilter
 * tree (scheduler.go:26-91, filter.go:44-187) followed by rand.Intn(len(pods)).
 *
 * The random pick.  The reference draws from Go's auto-seeded global source, so its pick is not
 * reproducible; this ABI defines it as what `rand.New(src).Intn(n)` returns (math/rand, Go 1.22:
 * Intn -> Int31n; Int31() = Int63()>>32) when src is a SplitMix64 source private to the request:
 *     state = seed ^ rand_key;
 *     next(): state += 0x9E3779B97F4A7C15; z = state;
 *             z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9; z = (z ^ (z >> 27)) * 0x94D049BB133111EB;
 *             return z ^ (z >> 31);
 *     Int63() = next() >> 1            (so Int31() = next() >> 33)
 *     Int31n(n): if n is a power of two: Int31() & (n-1)
 *                else max = 2^31 - 1 - (2^31 % n); v = Int31(); while (v > max) v = Int31();
 *                     return v % n
 * and pod_idx is the k-th survivor in ascending pod index (= slice order, pods[i] scheduler.go:121).
 */

/* Page-locked, device-mapped host memory for request / pick buffers (what the Go shim exposes
 * to callers as unsafe.Slice).  Buffers from here (or any cudaHostAlloc /
 * cudaHostRegister memory) are read and written by the kernel in place; ordinary pageable
 * memory also works but bounces through the ctx's own pinned buffers. */
void* lig_host_alloc(size_t bytes);
void  lig_host_free(void* p);

/* Host buffers in, host buffers out.  Nothing is staged through HBM: the pick kernel reads the
 * descriptors from, and writes the picks to, page-locked host memory over PCIe directly.  Blocks
 * until `out` is complete. */
int lig_schedule_batch(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* reqs, int R,
                       lig_pick* out);

/* Asynchronous submit for the one-goroutine-per-stream caller model (handlers/server.go:51, up to
 * 40 000 concurrent streams, pkg/manifests/ext_proc.yaml:102-105): many threads may have batches
 * in flight on one ctx at once.  `reqs` and `out` MUST be page-locked (lig_host_alloc /
 * cudaHostAlloc / cudaHostRegister) and stay untouched until the wait returns.  On success
 * *ticket identifies the in-flight batch; lig_schedule_wait blocks until its picks are complete and
 * releases the ticket.  The ctx lock is held only while the work is enqueued, never while waiting.
 * LIG_ERR_BUSY when all LIG_MAX_TICKETS tickets are in flight. */
#define LIG_MAX_TICKETS 256
int lig_schedule_batch_async(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* reqs, int R,
                             lig_pick* out, int* ticket);
int lig_schedule_wait(lig_ctx* ctx, int ticket);

/* HBM-resident batch: d_reqs / d_out are device pointers (16-byte / 8-byte aligned); the work is
 * enqueued on `stream` and NOT synchronised.  This is the class-table fast path. */
int lig_schedule_batch_device(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* d_reqs,
                              int R, lig_pick* d_out, void* stream);

/* A queue of n_batches HBM-resident batches of R requests each against one epoch, enqueued
 * back to back on `stream` with one pass through the ABI (the micro-batcher's flush of several
 * pending batches).  Batch b uses seed `seed + b`, reads d_reqs[b] and writes d_out[b]; the two
 * pointer arrays are host arrays of device pointers.  Equivalent to n_batches calls of
 * lig_schedule_batch_device. */
int lig_schedule_batches_device(lig_ctx* ctx, uint64_t epoch, uint64_t seed,
                                const lig_req* const* d_reqs, int R, lig_pick* const* d_out,
                                int n_batches, void* stream);

/* ---- streaming micro-batches without a kernel launch ---------------------------------------------
 * lig_stream_open parks one persistent CTA on the device that polls a mailbox in page-locked host
 * memory; lig_stream_submit writes up to lig_stream_capacity() descriptors + a doorbell ticket
 * into it and spins until the picks and the completion ticket are back: two PCIe crossings, no
 * launch, no stream synchronisation (Scheduler.Schedule under a sustained request stream,
 * handlers/request.go:72).  Same results as lig_schedule_batch for the same (epoch, seed, reqs).
 * While a stream is open nothing in the process may synchronise the whole device
 * (cudaDeviceSynchronize, cudaFree): the resident kernel never finishes until lig_stream_close. */
int lig_stream_capacity(void);
int lig_stream_open(lig_ctx* ctx);
int lig_stream_submit(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* reqs, int n,
                      lig_pick* out);
int lig_stream_close(lig_ctx* ctx);

/* Direct scan: every request walks the whole tree over all P pods itself (one warp per request,
 * no class tables).  d_masks, when not NULL, receives the survivor set of every request as
 * R x ceil(P/32) words (bit p%32 of word p/32 = pod p survives) — the GPU analogue of
 * Filter.Filter's return value (filter.go:12-15, 44-73), used by the parity tests. */
int lig_schedule_scan_device(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* d_reqs,
                             int R, lig_pick* d_out, uint32_t* d_masks, void* stream);

/* Host-buffer wrapper of the direct scan (test hook; masks may be NULL). */
int lig_schedule_scan(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* reqs, int R,
                      lig_pick* out, uint32_t* masks);

/* Read back one class table entry of a resident snapshot (test hook): the survivor list the
 * fast path uses for requests (critical, adapter_id).  `list` must hold P entries. */
int lig_read_class(lig_ctx* ctx, uint64_t epoch, int critical, int adapter_id, int* status,
                   int* n_survivors, uint16_t* list);

/* ---- the step before Schedule, on the device ------------------------------------------------------
 * Replaces, per request, the resolve step of HandleRequestBody (handlers/request.go:42-56):
 * datastore.FetchModelData(model) (backend/datastore.go:70-76), backend.RandomWeightedDraw over
 * Spec.TargetModels (datastore.go:78-98) and backend.IsCritical (datastore.go:100-105).  The host
 * interns the InferenceModel names to dense model ids once per datastore change and the request
 * shrinks to ONE 32-bit model id; the device looks the model up, draws the target model, derives
 * Critical and schedules — one pass over the batch, 4 bytes in and 4 bytes out per decision.
 *
 * The model table belongs to a resident snapshot epoch: target models are given as adapter ids
 * interned against that snapshot (an id outside [0, A) = a model no pod lists in ActiveModels).
 *   target_offsets[n_models + 1]   CSR offsets into the two target arrays (TargetModels order)
 *   target_adapter_ids / weights   Spec.TargetModels[k].Name (interned) / .Weight
 *   critical[n_models]             IsCritical(model)
 *   self_adapter_ids[n_models]     the model's own name interned: used when TargetModels is empty
 *                                  (request.go:47: modelName stays the requested model)
 *   present[n_models] (nullable)   0 = FetchModelData returns nil for this id -> LIG_NO_MODEL
 * Weights must be >= 0 and sum to [1, 2^31-1] for every model with targets (Go's Int31n panics on
 * a non-positive sum, a negative weight makes the reference's loop ill-defined), at most 255
 * targets per model: LIG_ERR_RANGE otherwise.
 *
 * The draw.  The reference draws from rand.NewSource(rand.Int63()) (request.go:48 passes seed 0):
 * unseeded, not reproducible.  As for the pick, this ABI defines it on a SplitMix64 source private
 * to the request: state = seed ^ rand_key ^ LIG_DRAW_DOMAIN, randomVal = Int31n(sum of weights)
 * (Go 1.22 algorithm, see below), target = first k with randomVal < Weight[k], else
 * randomVal -= Weight[k] (datastore.go:91-97).  A model with exactly one target skips the draw
 * (its result cannot depend on it).  The pick then uses state = seed ^ rand_key as always, so
 * lig_schedule_models_batch(ids) == lig_schedule_batch(descriptors resolved on the host).
 * rand_key of request i of a call is first_index + i (a counter-based key: shards of one batch
 * pass their offset and get the single-device result). */
#define LIG_DRAW_DOMAIN 0xA0761D6478BD642Full
typedef struct lig_mpick {
  int16_t pod_idx;     /* as lig_pick.pod_idx (-1 unless LIG_OK)                                  */
  uint8_t status;      /* LIG_OK / LIG_DROP / LIG_EMPTY / LIG_NO_MODEL                            */
  uint8_t target_idx;  /* index of the drawn Spec.TargetModels entry (the body's "model" rewrite,
                          request.go:61-69); 255 = TargetModels empty, the model name passes through */
} lig_mpick;
int lig_upload_models(lig_ctx* ctx, uint64_t epoch, int n_models, const int32_t* target_offsets,
                      const int32_t* target_adapter_ids, const int32_t* target_weights,
                      const uint8_t* critical, const int32_t* self_adapter_ids, const uint8_t* present);
/* Same, enqueued only (see lig_upload_snapshot_async). */
int lig_upload_models_async(lig_ctx* ctx, uint64_t epoch, int n_models, const int32_t* target_offsets,
                            const int32_t* target_adapter_ids, const int32_t* target_weights,
                            const uint8_t* critical, const int32_t* self_adapter_ids, const uint8_t* present);
/* Host buffers (page-locked ones are read / written over PCIe in place), blocking. */
int lig_schedule_models_batch(lig_ctx* ctx, uint64_t epoch, uint64_t seed, uint64_t first_index,
                              const uint32_t* model_ids, int R, lig_mpick* out);
/* HBM-resident queue of n_batches batches of R model ids (4-byte aligned; 16-byte aligned buffers
 * take the TMA path), batch b with seed + b; enqueued on `stream`, not synchronised. */
int lig_schedule_models_batches_device(lig_ctx* ctx, uint64_t epoch, uint64_t seed, uint64_t first_index,
                                       const uint32_t* const* d_model_ids, int R,
                                       lig_mpick* const* d_out, int n_batches, void* stream);
/* Test hook: only the resolve step, returning the 16-byte descriptor the device built for every
 * request (adapter_id, Critical, rand_key = first_index + i) and its status (LIG_OK / LIG_NO_MODEL)
 * and target index in `out`. */
int lig_resolve_models(lig_ctx* ctx, uint64_t epoch, uint64_t seed, uint64_t first_index,
                       const uint32_t* model_ids, int R, lig_req* reqs_out, lig_mpick* out);

/* ---- in-batch load feedback (opt-in; the default path never mutates the snapshot) ----------------
 * The reference schedules every request against the scraped metrics as they are
 * (scheduler.go:113-122 never writes them), so all requests of one class that arrive inside a
 * 50 ms scrape window land on the same handful of survivors.  The simulator does account for
 * load per pick (simulations/llm_ig_simulation/src/loadbalancer.py:608-625: the chosen pod's
 * prefill queue grows immediately).  This entry point does the same at sub-batch granularity:
 * the batch is cut into windows of `sub_batch` requests; after every window the number of picks
 * each pod received — summed over all ranks with one ncclAllReduce(int32[P]) when the ctx has a
 * communicator (lig_comm_init_rank) — is added to that pod's WaitingQueueSize in a PRIVATE copy of
 * the snapshot, the class tables are rebuilt, and the next window is scheduled against them.
 *   n_windows <= 0: ceil(R / sub_batch) windows; ranks of one communicator must pass the same
 *   n_windows (trailing windows may be empty).  d_hist (nullable): int32[P], receives the total
 *   picks per pod of this call (all ranks).  Resident epochs are not modified.  Enqueued on
 *   `stream`, not synchronised.  Defined sequentially in oracle/feedback.py. */
int lig_schedule_batch_feedback_device(lig_ctx* ctx, uint64_t epoch, uint64_t seed, const lig_req* d_reqs,
                                       int R, lig_pick* d_out, int sub_batch, int n_windows,
                                       int32_t* d_hist, void* stream);

/* ---- several GPUs, one process (the reference runs ONE scheduler per process, main.go:137) ------
 * A lig_group owns one lig_ctx per listed CUDA device and an NCCL communicator over them
 * (ncclCommInitAll).  The request batch shards BY REQUEST (decisions are independent given a
 * frozen snapshot: Schedule never mutates pod metrics, scheduler.go:113-122); the snapshot is
 * replicated with one ncclBroadcast per refresh tick, received directly into every member's
 * resident snapshot slot and consumed in place by the class-table build.  Picks need no
 * collective: device g owns the contiguous shard [R*g/G, R*(g+1)/G) of the result. */
typedef struct lig_group lig_group;
int  lig_group_create(lig_group** out, const int* devices, int n_devices, int max_pods,
                      int max_adapters, int max_batch);
void lig_group_destroy(lig_group* g);
int  lig_group_size(const lig_group* g);
lig_ctx* lig_group_ctx(lig_group* g, int member);   /* member ctx, e.g. for the device-pointer API */
int  lig_group_set_thresholds(lig_group* g, const lig_thresholds* t);
/* lig_upload_snapshot for the whole group: one pack, one H2D to member 0, one ncclBroadcast over
 * NVLink, one class-table build per member.  Resident on every member when it returns. */
int  lig_group_upload_snapshot(lig_group* g, uint64_t epoch, int P, int A, const double* kv,
                               const int32_t* q, const uint16_t* n_active,
                               const uint16_t* max_active, const uint32_t* bitmap_adapter_major);
/* lig_schedule_batch over the group: contiguous request shards, results in request order.  The
 * pick of a request depends on (seed, rand_key) only, so the result equals the single-device
 * result bit for bit.  Page-locked buffers must be portable (lig_host_alloc is). */
int  lig_group_schedule_batch(lig_group* g, uint64_t epoch, uint64_t seed, const lig_req* reqs,
                              int R, lig_pick* out);

/* ---- several GPUs, one process per GPU (torchrun-style launch) ----------------------------------
 * Rank 0 calls lig_comm_unique_id and ships the LIG_COMM_ID_BYTES to the other ranks by any means
 * (a file, an env var, the launcher's store); every rank then calls lig_comm_init_rank on its own
 * ctx.  lig_comm_upload_snapshot_device replicates the root's packed blob (device memory on the
 * root; ignored elsewhere) into every rank's snapshot slot with one ncclBroadcast enqueued on
 * `stream`, followed by the class-table build on the same stream; nothing is synchronised.
 * lig_comm_upload_snapshot is the blocking host-array form (arrays read on the root only). */
#define LIG_COMM_ID_BYTES 128
int lig_comm_unique_id(void* id /* LIG_COMM_ID_BYTES */);
int lig_comm_init_rank(lig_ctx* ctx, int n_ranks, int rank, const void* id);
int lig_comm_upload_snapshot_device(lig_ctx* ctx, uint64_t epoch, int P, int A, const void* d_blob,
                                    int root, void* stream);
int lig_comm_upload_snapshot(lig_ctx* ctx, uint64_t epoch, int P, int A, const double* kv,
                             const int32_t* q, const uint16_t* n_active, const uint16_t* max_active,
                             const uint32_t* bitmap_adapter_major, int root);
/* In-place sum of an int32 vector over the ranks (ncclAllReduce on `stream`): the per-pod pick
 * histogram of the load-feedback mode. */
int lig_comm_allreduce_i32(lig_ctx* ctx, int32_t* d_values, int n, void* stream);

/* ---- introspection ------------------------------------------------------------------------- */
const char* lig_last_error(void);            /* thread-local, never NULL */
const char* lig_version(void);
int         lig_abi_version(void);
int         lig_device_count(void);          /* 0 when no CUDA device / driver */
uint64_t    lig_kernel_launches(const lig_ctx* ctx);  /* kernels this ctx launched so far */
int         lig_sm_count(const lig_ctx* ctx);
/* The pick kernel a device-resident queue launches against `epoch` right now: its name, resident
 * grid and block size (bench.py checks its committed ncu capture against this), and the size of
 * the compact tables (0 = not available / strided tables in use). */
int         lig_pick_kernel_info(lig_ctx* ctx, uint64_t epoch, char* name, int name_len, int* grid,
                                 int* threads, int* table_bytes, int* tables_in_smem);

#ifdef __cplusplus
}
#endif
#endif /* LIG_H_ */
