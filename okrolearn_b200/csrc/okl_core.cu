// Context, memory pool, host<->device transfer with dtype conversion, CUDA graphs, timing and the
// NCCL data-parallel exchange of libokl_b200.  See include/okl.h for the contract of each entry point.
#include "common.cuh"
#include <mutex>
#include <vector>
#include <functional>
#include <condition_variable>
#include <thread>

#include <dlfcn.h>
#include <nccl.h>

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void okl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* okl_last_error(void) { return g_err; }
extern "C" const char* okl_version(void) { return "okl_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int okl_ctx_create(int device, int precision, okl_ctx** out) {
  OKL_REQUIRE(out != nullptr, "okl_ctx_create: out is NULL");
  OKL_REQUIRE(precision >= OKL_FP32 && precision <= OKL_BF16, "okl_ctx_create: bad precision %d", precision);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    okl_set_error("okl_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    cudaGetLastError();
    return OKL_ERR_CUDA;
  }
  OKL_REQUIRE(device >= 0 && device < ndev, "okl_ctx_create: device %d out of range [0,%d)", device, ndev);
  OKL_CHECK_CUDA(cudaSetDevice(device));
  okl_ctx* c = new okl_ctx();
  c->device = device;
  c->precision = precision;
  cudaDeviceProp prop;
  OKL_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->cc = prop.major * 10 + prop.minor;
  c->total_mem = prop.totalGlobalMem;
  if (const char* e = getenv("OKL_PDL")) c->pdl = atoi(e);
  // the compute stream carries the critical path: highest priority, so its CTAs are placed before those of the gradient exchange
  // (comm stream, lowest priority) that runs next to it; graph kernel nodes inherit the priority of the stream they were captured on
  int prio_lo = 0, prio_hi = 0;
  OKL_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  const char* pe = getenv("OKL_STREAM_PRIORITY");
  const bool use_prio = pe && atoi(pe) != 0;          // opt-in: the 2-GPU A/B was within run-to-run noise
  OKL_CHECK_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, use_prio ? prio_hi : 0));
  OKL_CHECK_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, use_prio ? prio_lo : 0));
  c->lanes[0].stream = c->stream;
  for (int i = 1; i < 3; i++) OKL_CHECK_CUDA(cudaStreamCreateWithFlags(&c->lanes[i].stream, cudaStreamNonBlocking));
  for (int i = 0; i < 3; i++) OKL_CHECK_CUDA(cudaEventCreateWithFlags(&c->lanes[i].ev, cudaEventDisableTiming));
  OKL_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  OKL_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  OKL_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_mark, cudaEventDisableTiming));
  OKL_CHECK_CUDA(cudaEventCreate(&c->ev_t0));
  OKL_CHECK_CUDA(cudaEventCreate(&c->ev_t1));
  *out = c;
  return OKL_OK;
}

extern "C" int okl_comm_destroy(okl_ctx* ctx);

static void lane_save(okl_ctx* c) {
  okl_ctx::Lane& l = c->lanes[c->cur_lane];
  l.stream = c->stream; l.workspace = c->workspace; l.workspace_bytes = c->workspace_bytes; l.tile_counters = c->tile_counters;
  l.bf16_scratch = c->bf16_scratch; l.bf16_scratch_bytes = c->bf16_scratch_bytes;
}
static void lane_load(okl_ctx* c, int lane) {
  okl_ctx::Lane& l = c->lanes[lane];
  c->stream = l.stream; c->workspace = l.workspace; c->workspace_bytes = l.workspace_bytes; c->tile_counters = l.tile_counters;
  c->bf16_scratch = l.bf16_scratch; c->bf16_scratch_bytes = l.bf16_scratch_bytes;
  c->cur_lane = lane;
}

// Subsequent launches go to auxiliary lane `lane` (1 or 2), ordered after everything enqueued so far on the main lane.
extern "C" int okl_aux_begin(okl_ctx* ctx, int lane) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_REQUIRE(lane >= 1 && lane <= 2, "aux lane must be 1 or 2");
  OKL_REQUIRE(ctx->cur_lane == 0, "okl_aux_begin: already on an auxiliary lane");
  if (!ctx->fork_marked) OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_mark, ctx->stream));
  ctx->fork_marked = false;
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->lanes[lane].stream, ctx->ev_mark, 0));
  lane_save(ctx);
  lane_load(ctx, lane);
  ctx->lanes[lane].dirty = true;
  ctx->aux_dirty = true;
  return OKL_OK;
}

// Remember "now" on the main lane as the point the next okl_aux_begin forks from: lets the caller enqueue the critical-path kernel
// on the main lane FIRST (so it is scheduled first) and the leaf kernels afterwards, while both only depend on the marked point.
extern "C" int okl_aux_mark(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_REQUIRE(ctx->cur_lane == 0, "okl_aux_mark: not on the main lane");
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_mark, ctx->stream));
  ctx->fork_marked = true;
  return OKL_OK;
}

// Back to the main lane (the auxiliary work keeps running concurrently).
extern "C" int okl_aux_end(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (ctx->cur_lane == 0) return OKL_OK;
  lane_save(ctx);
  lane_load(ctx, 0);
  return OKL_OK;
}

// The main lane waits for everything issued on the auxiliary lanes.
extern "C" int okl_aux_join(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (ctx->cur_lane != 0) { lane_save(ctx); lane_load(ctx, 0); }
  if (!ctx->aux_dirty) return OKL_OK;
  for (int i = 1; i < 3; i++) {
    if (!ctx->lanes[i].dirty) continue;
    OKL_CHECK_CUDA(cudaEventRecord(ctx->lanes[i].ev, ctx->lanes[i].stream));
    OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->lanes[i].ev, 0));
    ctx->lanes[i].dirty = false;
  }
  ctx->aux_dirty = false;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (void* p : ctx->deferred_free) {
      auto it = ctx->live.find(p);
      if (it != ctx->live.end()) ctx->free_blocks.emplace(it->second, p);
    }
    ctx->deferred_free.clear();
  }
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------ batch pipeline
// Step s+1's batch is gathered (from HBM, or over PCIe from page-locked host memory) into the other input slot on a copy stream
// while step s computes.  The copy forks from the compute stream's current position - the previous user of that slot has been
// enqueued by then - and the step that consumes the slot waits for the slot's event.
static int ensure_pipeline(okl_ctx* ctx) {
  if (ctx->copy_stream) return OKL_OK;
  OKL_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  OKL_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy_fork, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    OKL_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_slot[i], cudaEventDisableTiming));
    OKL_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_loss[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < OKL_FLOW_DEPTH; i++) OKL_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_flow[i], cudaEventDisableTiming));
  OKL_CHECK_CUDA(cudaMallocHost(&ctx->loss_pinned, 4 * sizeof(float)));
  memset(ctx->loss_pinned, 0, 4 * sizeof(float));
  return OKL_OK;
}

// ---- stream mode: the batch crosses PCIe on the COPY ENGINES.  The gather kernel reads page-locked host memory zero-copy - and needs
// SM resources to do so: beside the persistent tensor-core kernels of a step (384 threads x 168 registers = the whole register file
// of an SM) none of its blocks becomes resident, so it ran BETWEEN them and delayed them by the PCIe time of a batch (cifar6, batch
// 1024: 1.45 ms per step end to end against 1.24 ms device-resident, independent of the step count and of the loss read-back;
// capping those kernels at 128 registers to make room cost more than it gave).  With the batch's row indices also known on the host,
// a few host threads gather the rows into a page-locked staging ring (one entry per batch the host may run ahead of the device) and ONE
// cudaMemcpyAsync per part moves them: DMA, no SM involved.  The gather of batch s+1 runs on the host while the device computes step s.
extern "C" int okl_batch_host_index(okl_ctx* ctx, const void* idx_host_i32) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  ctx->host_idx_next = (const int*)idx_host_i32;
  return OKL_OK;
}

static bool is_host_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

namespace {
// a handful of persistent worker threads for row gathers (std::thread creation per batch would cost as much as the gather)
class HostPool {
 public:
  explicit HostPool(int n) : stop_(false), gen_(0), pending_(0) {
    for (int i = 0; i < n; i++) workers_.emplace_back([this, i] { run(i); });
  }
  ~HostPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // fn(worker) on every worker, returns when all are done
  void run_all(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn; pending_ = (int)workers_.size(); gen_++;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }
 private:
  void run(int id) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* f;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        f = fn_;
      }
      if (f) (*f)(id);
      { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) done_.notify_all(); }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  bool stop_;
  unsigned long long gen_;
  int pending_;
};

// The gather of a 12 MB batch has to finish well inside a ~1 ms step: that takes ~6 memory-bound threads.  Every rank of a data-parallel
// job runs its own pool on the same host, so the pool exists only when the rank's share of the host cores covers it (16 cores: 1 or
// 2 ranks; measured at 8 ranks on 16 cores: 2.40 ms per step staged vs 2.02 ms with the gather kernel) - otherwise the gather kernel.
HostPool* host_pool(int nranks) {
  static HostPool* pool = nullptr;
  static std::once_flag once;
  std::call_once(once, [nranks] {
    int n = 6;
    if (const char* e = getenv("OKL_HOST_GATHER_THREADS")) n = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    const int share = hw > 0 ? hw / (nranks > 0 ? nranks : 1) - 1 : n;
    if (share < 5 && !getenv("OKL_HOST_GATHER_THREADS")) n = 0;
    else if (n > share && share >= 1) n = share;
    if (n >= 1) pool = new HostPool(n);
  });
  return pool;
}
}  // namespace

// rows idx[0 .. rows) of the host array src (row_bytes each) -> page-locked staging -> dst with one DMA copy; OKL_OK, or a negative
// value when staging is unavailable (the caller falls back to the gather kernel)
static int staged_gather_rows(okl_ctx* ctx, int ring, int part, void* dst, const void* src, size_t row_bytes, const int* idx, long long rows) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("OKL_DMA_GATHER"); enabled = e ? atoi(e) : 1; }
  const size_t bytes = row_bytes * (size_t)rows;
  if (!enabled || bytes > (size_t(64) << 20)) return -1;   // beyond that the host gather itself (memory-bound threads) would be the bottleneck
  void*& stg = ctx->stage[part][ring];
  size_t& cap = ctx->stage_bytes[part][ring];
  if (cap < bytes) {
    if (stg) cudaFreeHost(stg);
    stg = nullptr; cap = 0;
    if (cudaMallocHost(&stg, bytes) != cudaSuccess) { cudaGetLastError(); stg = nullptr; return -1; }
    cap = bytes;
  }
  char* out = (char*)stg;
  const char* in = (const char*)src;
  HostPool* pool = host_pool(ctx->nranks);
  if (!pool && bytes >= (size_t(1) << 20)) return -1;        // no cores to spare for the gather: the gather kernel takes this part
  if (pool && bytes >= (size_t(1) << 20)) {
    const int nw = pool->size();
    pool->run_all([&](int w) {
      const long long r0 = rows * w / nw, r1 = rows * (w + 1) / nw;
      for (long long i = r0; i < r1; i++) memcpy(out + (size_t)i * row_bytes, in + (size_t)idx[i] * row_bytes, row_bytes);
    });
  } else {
    for (long long i = 0; i < rows; i++) memcpy(out + (size_t)i * row_bytes, in + (size_t)idx[i] * row_bytes, row_bytes);
  }
  if (cudaMemcpyAsync(dst, stg, bytes, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess) { cudaGetLastError(); return -1; }
  return OKL_OK;
}

extern "C" int okl_batch_prefetch(okl_ctx* ctx, int slot, void* x_dst, const void* x_src, long long x_row_elems, void* t_dst,
                                  const void* t_src, long long t_row_elems, const void* idx, long long rows, void* x_nxc_dst, int x_channels,
                                  int x_nxc_bf16, void* idx_dst) {
  OKL_REQUIRE(ctx && (slot == 0 || slot == 1), "okl_batch_prefetch: slot must be 0 or 1");
  OKL_REQUIRE(!ctx->capturing && ctx->cur_lane == 0, "okl_batch_prefetch: not inside a capture / auxiliary lane");
  int s = ensure_pipeline(ctx);
  if (s != OKL_OK) return s;
  // flow control: the host may run at most OKL_FLOW_DEPTH batches ahead of the device (an unbounded launch backlog occasionally
  // made step times bimodal; a depth of 1-2 makes the device wait for host wake-ups)
  static int depth = 0;
  if (depth == 0) { const char* e = getenv("OKL_FLOW_DEPTH"); depth = e ? atoi(e) : OKL_FLOW_DEPTH; if (depth < 1 || depth > OKL_FLOW_DEPTH) depth = OKL_FLOW_DEPTH; }
  const int fi = ctx->flow_next % depth;
  if (ctx->flow_next >= (unsigned long long)depth) OKL_CHECK_CUDA(cudaEventSynchronize(ctx->ev_flow[fi]));
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_copy_fork, ctx->stream));
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_fork, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->copy_stream;
  // A channels-last first layer (tensor-core convolution): the (N, C, S) -> (N, S, C) permutation of the batch also happens here, on
  // the copy stream, instead of at the head of the step's critical path.  As bf16 it is produced straight from the dataset rows
  // (through the index) when no fp32 copy of the batch is wanted (x_dst NULL).
  const bool nxc = x_nxc_dst && x_channels > 1 && rows > 0 && x_row_elems % x_channels == 0;
  const int* hidx = ctx->host_idx_next;
  ctx->host_idx_next = nullptr;
  bool x_done = false;
  // stream mode with host-side indices, batches of at least 1 MB (small batches belong to short steps of small kernels, which the
  // gather kernel runs beside without trouble, and the host has only tens of microseconds per step there): staged DMA
  if (hidx && rows > 0 && (size_t)x_row_elems * 4 * (size_t)rows >= (size_t(1) << 20)) {
    // ring entry fi is free: the flow-control wait above covered the copy that last read it
    if (x_dst && x_src && x_row_elems > 0 && is_host_ptr(x_src) &&
        staged_gather_rows(ctx, fi, 0, x_dst, x_src, (size_t)x_row_elems * 4, hidx, rows) == OKL_OK)
      x_done = true;                                         // the kernel below skips this part
    if (x_done && t_dst && t_src && t_row_elems > 0 && is_host_ptr(t_src) &&
        staged_gather_rows(ctx, fi, 1, t_dst, t_src, (size_t)t_row_elems * 4, hidx, rows) == OKL_OK)
      t_dst = nullptr;
  }
  s = okl_gather_batch(ctx, x_done ? nullptr : x_dst, x_src, x_row_elems, t_dst, t_src, t_row_elems, idx, rows);
  if (s == OKL_OK && nxc) {
    if (x_nxc_bf16) s = okl_to_bf16_nxc(ctx, x_dst ? x_dst : x_src, x_dst ? nullptr : idx, x_nxc_dst, (int)rows, x_channels, x_row_elems / x_channels);
    else if (x_dst) s = okl_permute(ctx, x_nxc_dst, x_dst, (int)rows, x_channels, x_row_elems / x_channels, OKL_NCX);
    else { okl_set_error("okl_batch_prefetch: the fp32 channels-last copy needs the gathered batch (x_dst)"); s = OKL_ERR_INVALID; }
  }
  // consumers that read dataset rows through the index (okl_*_rows arguments) get the batch's index slice at a fixed per-slot address
  if (s == OKL_OK && idx_dst && rows > 0)
    s = cudaMemcpyAsync(idx_dst, idx, (size_t)rows * sizeof(int), cudaMemcpyDeviceToDevice, ctx->copy_stream) == cudaSuccess ? OKL_OK : OKL_ERR_CUDA;
  ctx->stream = main_stream;
  if (s != OKL_OK) return s;
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_slot[slot], ctx->copy_stream));
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_flow[fi], ctx->copy_stream));
  ctx->flow_next++;
  return OKL_OK;
}

extern "C" int okl_batch_wait(okl_ctx* ctx, int slot) {
  OKL_REQUIRE(ctx && (slot == 0 || slot == 1) && ctx->copy_stream, "okl_batch_wait: no prefetch issued for this slot");
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_slot[slot], 0));
  return OKL_OK;
}

// Loss read-back without draining the pipeline: the scalar is copied to page-locked host memory in stream order right after its
// step; the host collects it (okl_loss_fetch_wait) once the NEXT step has been enqueued.
extern "C" int okl_loss_fetch_async(okl_ctx* ctx, int slot, const void* dev_f32) {
  OKL_REQUIRE(ctx && dev_f32 && (slot == 0 || slot == 1), "okl_loss_fetch_async: bad argument");
  OKL_REQUIRE(!ctx->capturing, "okl_loss_fetch_async: not inside a capture");
  int s = ensure_pipeline(ctx);
  if (s != OKL_OK) return s;
  if (ctx->aux_dirty) {
    s = okl_aux_join(ctx);
    if (s != OKL_OK) return s;
  }
  // a one-thread kernel posts {value, sequence number} into page-locked host memory and the host polls the sequence number.
  // It runs on the COPY stream, behind an event recorded on the compute stream: anything enqueued between two step graphs on the
  // compute stream itself (DMA copy or kernel with a system-scope fence) costs ~10 us of step time, measured.
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_copy_fork, ctx->stream));
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_fork, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->copy_stream;
  s = okl_loss_post(ctx, ctx->loss_pinned + 2 * slot, dev_f32, ++ctx->loss_seq[slot]);
  ctx->stream = main_stream;
  return s;
}

extern "C" int okl_loss_fetch_wait(okl_ctx* ctx, int slot, float* out) {
  OKL_REQUIRE(ctx && out && (slot == 0 || slot == 1) && ctx->loss_pinned, "okl_loss_fetch_wait: nothing to wait for");
  volatile unsigned* seq = reinterpret_cast<volatile unsigned*>(ctx->loss_pinned + 2 * slot + 1);
  const unsigned want = ctx->loss_seq[slot];
  for (unsigned long long spins = 0; *seq != want; spins++) {
    if ((spins & 0xfffff) == 0xfffff) {                     // every ~1M polls: has the stream died?
      cudaError_t e = cudaStreamQuery(ctx->copy_stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) {
        okl_set_error("okl_loss_fetch_wait: stream failed: %s", cudaGetErrorString(e));
        return OKL_ERR_CUDA;
      }
      if (e == cudaSuccess && *seq != want) {                // stream drained but the post never arrived
        okl_set_error("okl_loss_fetch_wait: no loss was posted for slot %d", slot);
        return OKL_ERR_INVALID;
      }
    }
  }
  *out = *reinterpret_cast<volatile float*>(ctx->loss_pinned + 2 * slot);
  return OKL_OK;
}

extern "C" int okl_ctx_destroy(okl_ctx* ctx) {
  if (!ctx) return OKL_OK;
  cudaSetDevice(ctx->device);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_copy_fork);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_slot[i]); cudaEventDestroy(ctx->ev_loss[i]); }
    for (int i = 0; i < OKL_FLOW_DEPTH; i++) cudaEventDestroy(ctx->ev_flow[i]);
    cudaFreeHost(ctx->loss_pinned);
  }
  okl_aux_join(ctx);
  for (int i = 1; i < 3; i++) {
    cudaStreamSynchronize(ctx->lanes[i].stream);
    if (ctx->lanes[i].workspace) cudaFree(ctx->lanes[i].workspace);
    if (ctx->lanes[i].tile_counters) cudaFree(ctx->lanes[i].tile_counters);
    if (ctx->lanes[i].bf16_scratch) cudaFree(ctx->lanes[i].bf16_scratch);
    cudaStreamDestroy(ctx->lanes[i].stream);
  }
  for (int i = 0; i < 3; i++) cudaEventDestroy(ctx->lanes[i].ev);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->comm_stream);
  okl_comm_destroy(ctx);
  for (auto& kv : ctx->live) cudaFree(kv.first);
  if (ctx->leaf_scratch) cudaFree(ctx->leaf_scratch);
  for (int p = 0; p < 2; p++)
    for (int i = 0; i < OKL_FLOW_DEPTH; i++) if (ctx->stage[p][i]) cudaFreeHost(ctx->stage[p][i]);
  if (ctx->workspace) cudaFree(ctx->workspace);
  if (ctx->staging) cudaFree(ctx->staging);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->bf16_scratch) cudaFree(ctx->bf16_scratch);
  if (ctx->tile_counters) cudaFree(ctx->tile_counters);
  if (ctx->head_scratch) cudaFree(ctx->head_scratch);
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_join);
  cudaEventDestroy(ctx->ev_t0);
  cudaEventDestroy(ctx->ev_t1);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
  return OKL_OK;
}

extern "C" int okl_ctx_sync(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (!ctx->capturing) {
    int s = okl_aux_join(ctx);
    if (s != OKL_OK) return s;
  }
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  return OKL_OK;
}

extern "C" int okl_ctx_set_precision(okl_ctx* ctx, int precision) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_REQUIRE(precision >= OKL_FP32 && precision <= OKL_BF16, "bad precision %d", precision);
  ctx->precision = precision;
  return OKL_OK;
}

extern "C" int okl_ctx_get_precision(okl_ctx* ctx, int* precision) {
  OKL_REQUIRE(ctx && precision, "NULL argument");
  *precision = ctx->precision;
  return OKL_OK;
}

extern "C" int okl_device_info(okl_ctx* ctx, int* sm_count, size_t* total_mem, int* cc) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (sm_count) *sm_count = ctx->sm_count;
  if (total_mem) *total_mem = ctx->total_mem;
  if (cc) *cc = ctx->cc;
  return OKL_OK;
}

extern "C" int okl_launch_count(okl_ctx* ctx, unsigned long long* out) {
  OKL_REQUIRE(ctx && out, "NULL argument");
  *out = ctx->launches;
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------
// caching device allocator.  Blocks are rounded to 512 B (small) / 2 MiB (>= 1 MiB) and recycled by
// exact rounded size, so a steady-state train step never calls cudaMalloc (and a captured graph's
// buffers stay where they are).
// ------------------------------------------------------------------------------------------------
static size_t round_size(size_t b) {
  if (b == 0) b = 1;
  if (b < (1u << 20)) return (b + 511) & ~size_t(511);
  return (b + (2u << 20) - 1) & ~size_t((2u << 20) - 1);
}

extern "C" int okl_malloc(okl_ctx* ctx, size_t bytes, void** out) {
  OKL_REQUIRE(ctx && out, "NULL argument");
  size_t r = round_size(bytes);
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->free_blocks.find(r);
  if (it != ctx->free_blocks.end()) {
    *out = it->second;
    ctx->free_blocks.erase(it);
    return OKL_OK;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, r);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    // release the cache and retry once
    for (auto& kv : ctx->free_blocks) {
      cudaFree(kv.second);
      ctx->live.erase(kv.second);
    }
    ctx->free_blocks.clear();
    e = cudaMalloc(&p, r);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    okl_set_error("okl_malloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? OKL_ERR_OOM : OKL_ERR_CUDA;
  }
  ctx->live[p] = r;
  *out = p;
  return OKL_OK;
}

extern "C" int okl_free(okl_ctx* ctx, void* p) {
  if (!p) return OKL_OK;
  OKL_REQUIRE(ctx, "ctx is NULL");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->live.find(p);
  OKL_REQUIRE(it != ctx->live.end(), "okl_free: %p was not returned by okl_malloc", p);
  // stream-ordered reuse: every consumer runs on ctx->stream (the comm stream is joined before reuse).  A block released while
  // a CUDA graph is being captured stays out of the pool: the graph's kernels keep writing to it on every replay.
  if (ctx->capturing) return OKL_OK;
  if (ctx->aux_dirty) {                   // an auxiliary lane may still be reading this block: recycle it at the next join
    ctx->deferred_free.push_back(p);
    return OKL_OK;
  }
  ctx->free_blocks.emplace(it->second, p);
  return OKL_OK;
}

extern "C" int okl_pool_reserve(okl_ctx* ctx, size_t bytes) {
  void* p = nullptr;
  int s = okl_malloc(ctx, bytes, &p);
  if (s != OKL_OK) return s;
  return okl_free(ctx, p);
}

int okl_ensure_workspace(okl_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->workspace_bytes) return OKL_OK;
  if (ctx->capturing) {
    okl_set_error("workspace of %zu bytes needed during graph capture but only %zu reserved (run one eager step first)",
                  bytes, ctx->workspace_bytes);
    return OKL_ERR_INVALID;
  }
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->workspace) cudaFree(ctx->workspace);
  size_t nb = round_size(bytes + bytes / 4);
  OKL_CHECK_CUDA(cudaMalloc(&ctx->workspace, nb));
  ctx->workspace_bytes = nb;
  return OKL_OK;
}

int okl_ensure_staging(okl_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->staging_bytes) return OKL_OK;
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->staging) cudaFree(ctx->staging);
  size_t nb = round_size(bytes);
  OKL_CHECK_CUDA(cudaMalloc(&ctx->staging, nb));
  ctx->staging_bytes = nb;
  return OKL_OK;
}

extern "C" int okl_memset(okl_ctx* ctx, void* p, int byte, size_t bytes) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (bytes == 0) return OKL_OK;
  OKL_CHECK_CUDA(cudaMemsetAsync(p, byte, bytes, ctx->stream));
  return OKL_OK;
}

extern "C" int okl_pinned_alloc(size_t bytes, void** out) {
  OKL_REQUIRE(out, "out is NULL");
  OKL_CHECK_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return OKL_OK;
}

extern "C" int okl_pinned_free(void* p) {
  if (p) OKL_CHECK_CUDA(cudaFreeHost(p));
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------
// transfers with on-device dtype conversion
// ------------------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = (D)src[i];
}

static size_t dtype_size(int dt) {
  switch (dt) {
    case OKL_F32: return 4;
    case OKL_F64: return 8;
    case OKL_I64: return 8;
    case OKL_I32: return 4;
    case OKL_U8: return 1;
  }
  return 0;
}

template <typename D>
static int h2d_convert(okl_ctx* ctx, D* dst, const void* host, size_t n, int src_dtype, bool dst_is_same_raw) {
  size_t es = dtype_size(src_dtype);
  OKL_REQUIRE(es != 0, "bad src dtype %d", src_dtype);
  if (n == 0) return OKL_OK;
  if (dst_is_same_raw) {
    OKL_CHECK_CUDA(cudaMemcpyAsync(dst, host, n * es, cudaMemcpyHostToDevice, ctx->stream));
    OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));   // host buffer is borrowed for this call only
    return OKL_OK;
  }
  int s = okl_ensure_staging(ctx, n * es);
  if (s != OKL_OK) return s;
  OKL_CHECK_CUDA(cudaMemcpyAsync(ctx->staging, host, n * es, cudaMemcpyHostToDevice, ctx->stream));
  int grid = okl_grid_1d(ctx, n, 256);
  switch (src_dtype) {
    case OKL_F32: convert_kernel<float, D><<<grid, 256, 0, ctx->stream>>>((const float*)ctx->staging, dst, n); break;
    case OKL_F64: convert_kernel<double, D><<<grid, 256, 0, ctx->stream>>>((const double*)ctx->staging, dst, n); break;
    case OKL_I64: convert_kernel<long long, D><<<grid, 256, 0, ctx->stream>>>((const long long*)ctx->staging, dst, n); break;
    case OKL_I32: convert_kernel<int, D><<<grid, 256, 0, ctx->stream>>>((const int*)ctx->staging, dst, n); break;
    case OKL_U8: convert_kernel<unsigned char, D><<<grid, 256, 0, ctx->stream>>>((const unsigned char*)ctx->staging, dst, n); break;
  }
  OKL_LAUNCHED(ctx, "convert_kernel");
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return OKL_OK;
}

extern "C" int okl_h2d(okl_ctx* ctx, void* dst_f32, const void* host, size_t n, int src_dtype) {
  OKL_REQUIRE(ctx && (n == 0 || (dst_f32 && host)), "NULL argument");
  return h2d_convert<float>(ctx, (float*)dst_f32, host, n, src_dtype, src_dtype == OKL_F32);
}

extern "C" int okl_h2d_i32(okl_ctx* ctx, void* dst_i32, const void* host, size_t n, int src_dtype) {
  OKL_REQUIRE(ctx && (n == 0 || (dst_i32 && host)), "NULL argument");
  return h2d_convert<int>(ctx, (int*)dst_i32, host, n, src_dtype, src_dtype == OKL_I32);
}

extern "C" int okl_h2d_async_raw(okl_ctx* ctx, void* dst, const void* pinned_host, size_t bytes) {
  OKL_REQUIRE(ctx && (bytes == 0 || (dst && pinned_host)), "NULL argument");
  if (bytes) OKL_CHECK_CUDA(cudaMemcpyAsync(dst, pinned_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return OKL_OK;
}

extern "C" int okl_d2h_async_raw(okl_ctx* ctx, void* pinned_host, const void* src, size_t bytes) {
  OKL_REQUIRE(ctx && (bytes == 0 || (src && pinned_host)), "NULL argument");
  if (bytes) OKL_CHECK_CUDA(cudaMemcpyAsync(pinned_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return OKL_OK;
}

extern "C" int okl_d2h(okl_ctx* ctx, void* host, const void* src_f32, size_t n, int dst_dtype) {
  OKL_REQUIRE(ctx && (n == 0 || (src_f32 && host)), "NULL argument");
  OKL_REQUIRE(dst_dtype == OKL_F32 || dst_dtype == OKL_F64, "okl_d2h: dst dtype must be F32 or F64");
  if (n == 0) return OKL_OK;
  if (ctx->aux_dirty && !ctx->capturing) {
    int sj = okl_aux_join(ctx);
    if (sj != OKL_OK) return sj;
  }
  if (ctx->comm_pending) {
    int s = okl_comm_wait(ctx);
    if (s != OKL_OK) return s;
  }
  if (dst_dtype == OKL_F32) {
    OKL_CHECK_CUDA(cudaMemcpyAsync(host, src_f32, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    int s = okl_ensure_staging(ctx, n * 8);
    if (s != OKL_OK) return s;
    convert_kernel<float, double><<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)src_f32,
                                                                                   (double*)ctx->staging, n);
    OKL_LAUNCHED(ctx, "convert_kernel");
    OKL_CHECK_CUDA(cudaMemcpyAsync(host, ctx->staging, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return OKL_OK;
}

extern "C" int okl_d2d(okl_ctx* ctx, void* dst, const void* src, size_t bytes) {
  OKL_REQUIRE(ctx && (bytes == 0 || (dst && src)), "NULL argument");
  if (bytes) OKL_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return OKL_OK;
}

extern "C" int okl_read_scalar(okl_ctx* ctx, const void* dev_f32, float* out) {
  OKL_REQUIRE(ctx && dev_f32 && out, "NULL argument");
  if (ctx->aux_dirty && !ctx->capturing) {
    int sj = okl_aux_join(ctx);
    if (sj != OKL_OK) return sj;
  }
  OKL_CHECK_CUDA(cudaMemcpyAsync(out, dev_f32, 4, cudaMemcpyDeviceToHost, ctx->stream));
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------
// CUDA graphs
// ------------------------------------------------------------------------------------------------
extern "C" int okl_graph_begin(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_REQUIRE(!ctx->capturing, "okl_graph_begin: already capturing");
  OKL_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  ctx->capturing = true;
  ctx->capture_launches = 0;
  return OKL_OK;
}

extern "C" int okl_graph_end(okl_ctx* ctx, okl_graph** out) {
  OKL_REQUIRE(ctx && out, "NULL argument");
  OKL_REQUIRE(ctx->capturing, "okl_graph_end: not capturing");
  {
    int sj = okl_aux_join(ctx);          // auxiliary lanes must rejoin the origin stream before the capture ends
    if (sj != OKL_OK) return sj;
  }
  if (ctx->comm_pending) {                         // join the comm stream back before ending the capture
    int s = okl_comm_wait(ctx);
    if (s != OKL_OK) return s;
  }
  ctx->capturing = false;
  okl_graph* g = new okl_graph();
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
  if (e != cudaSuccess) {
    okl_set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    delete g;
    return OKL_ERR_CUDA;
  }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    okl_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    cudaGraphDestroy(g->graph);
    delete g;
    return OKL_ERR_CUDA;
  }
  g->launches = ctx->capture_launches;
  *out = g;
  return OKL_OK;
}

extern "C" int okl_graph_launch(okl_ctx* ctx, okl_graph* g) {
  OKL_REQUIRE(ctx && g && g->exec, "NULL argument");
  OKL_CHECK_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += g->launches;
  return OKL_OK;
}

extern "C" int okl_graph_destroy(okl_graph* g) {
  if (!g) return OKL_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------
// timing
// ------------------------------------------------------------------------------------------------
extern "C" int okl_timer_start(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
  return OKL_OK;
}

extern "C" int okl_timer_stop(okl_ctx* ctx, float* ms) {
  OKL_REQUIRE(ctx && ms, "NULL argument");
  {
    int sj = okl_aux_join(ctx);
    if (sj != OKL_OK) return sj;
  }
  if (ctx->comm_pending) {
    int s = okl_comm_wait(ctx);
    if (s != OKL_OK) return s;
  }
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_t1, ctx->stream));
  OKL_CHECK_CUDA(cudaEventSynchronize(ctx->ev_t1));
  OKL_CHECK_CUDA(cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
  return OKL_OK;
}

__global__ void flush_kernel(float4* __restrict__ p, size_t n4, float v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) p[i] = make_float4(v, v, v, v);
}

extern "C" int okl_flush_l2(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (!ctx->flush_buf) {
    ctx->flush_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
    OKL_CHECK_CUDA(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((float4*)ctx->flush_buf, ctx->flush_bytes / 16, 1.0f);
  cudaError_t e = cudaPeekAtLastError();          // not counted: not part of the product path
  if (e != cudaSuccess) {
    okl_set_error("flush kernel: %s", cudaGetErrorString(e));
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------
// NCCL (resolved with dlopen so the library loads on machines without NCCL; one rank per process)
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.GetErrorString;
}

int need_nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) {
    okl_set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    return OKL_ERR_NCCL;
  }
  return OKL_OK;
}
}  // namespace

#define OKL_CHECK_NCCL(expr)                                                              \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != ncclSuccess) {                                                              \
      okl_set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                   \
      return OKL_ERR_NCCL;                                                                \
    }                                                                                     \
  } while (0)

extern "C" int okl_comm_unique_id(void* out128) {
  OKL_REQUIRE(out128, "out128 is NULL");
  int s = need_nccl();
  if (s != OKL_OK) return s;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  OKL_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return OKL_OK;
}

extern "C" int okl_comm_init(okl_ctx* ctx, const void* id128, int rank, int nranks) {
  OKL_REQUIRE(ctx && id128, "NULL argument");
  OKL_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / nranks %d", rank, nranks);
  int s = need_nccl();
  if (s != OKL_OK) return s;
  OKL_CHECK_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  OKL_CHECK_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return OKL_OK;
}

extern "C" int okl_comm_destroy(okl_ctx* ctx) {
  if (ctx && ctx->nccl_comm && g_nccl.ok) {
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
  }
  return OKL_OK;
}

extern "C" int okl_comm_wait(okl_ctx* ctx);
extern "C" int okl_allreduce_async(okl_ctx* ctx, void* p, size_t count) {
  OKL_REQUIRE(ctx && p, "NULL argument");
  if (ctx->nranks == 1 || count == 0) return OKL_OK;
  if (okl_p2p_applicable(ctx, p, count)) {          // small bucket inside the symmetric buffer: one peer-memory kernel, no NCCL
    if (ctx->comm_pending && ctx->comm_p2p_kind == (okl_p2p_is_ll(count) ? 2 : 1)) {   // same flags / staging in flight: serialise
      int s = okl_comm_wait(ctx);
      if (s != OKL_OK) return s;
    }
    return okl_p2p_allreduce(ctx, p, count);
  }
  OKL_REQUIRE(ctx->nccl_comm, "okl_allreduce_async: communicator not initialised");
  // fork: comm stream waits for everything enqueued so far on the compute stream (works under capture too)
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
  OKL_CHECK_NCCL(g_nccl.AllReduce(p, p, count, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->comm_stream));
  ctx->comm_pending = true;
  return OKL_OK;
}

// Early exchange: the all-reduce runs on the comm stream as soon as everything enqueued so far on the main lane AND on the
// auxiliary lanes (where the weight gradients are produced) has finished, concurrently with whatever the caller enqueues next
// (the remaining backward); okl_comm_wait joins it.
extern "C" int okl_allreduce_fork(okl_ctx* ctx, void* p, size_t count) {
  OKL_REQUIRE(ctx && p, "NULL argument");
  if (ctx->nranks == 1 || count == 0) return OKL_OK;
  OKL_REQUIRE(ctx->cur_lane == 0, "okl_allreduce_fork: call from the main lane");
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
  for (int i = 1; i < 3; i++) {
    if (!ctx->lanes[i].dirty) continue;
    OKL_CHECK_CUDA(cudaEventRecord(ctx->lanes[i].ev, ctx->lanes[i].stream));
    OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->lanes[i].ev, 0));
  }
  ctx->comm_pending = true;
  if (okl_p2p_applicable(ctx, p, count)) {
    ctx->comm_p2p_kind = okl_p2p_is_ll(count) ? 2 : 1;
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->comm_stream;
    int s = okl_p2p_allreduce(ctx, p, count);
    ctx->stream = main_stream;
    return s;
  }
  OKL_REQUIRE(ctx->nccl_comm, "okl_allreduce_fork: communicator not initialised");
  OKL_CHECK_NCCL(g_nccl.AllReduce(p, p, count, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->comm_stream));
  return OKL_OK;
}

extern "C" int okl_comm_wait(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (!ctx->comm_pending) return OKL_OK;
  OKL_CHECK_CUDA(cudaEventRecord(ctx->ev_join, ctx->comm_stream));
  OKL_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  ctx->comm_pending = false;
  ctx->comm_p2p_kind = 0;
  return OKL_OK;
}

extern "C" int okl_allreduce_max_host(okl_ctx* ctx, double* value) {
  OKL_REQUIRE(ctx && value, "NULL argument");
  if (ctx->nranks == 1) return OKL_OK;
  OKL_REQUIRE(ctx->nccl_comm, "communicator not initialised");
  double* d = nullptr;
  int s = okl_malloc(ctx, sizeof(double), (void**)&d);
  if (s != OKL_OK) return s;
  OKL_CHECK_CUDA(cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  OKL_CHECK_NCCL(g_nccl.AllReduce(d, d, 1, ncclFloat64, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  OKL_CHECK_CUDA(cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  OKL_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return okl_free(ctx, d);
}

extern "C" int okl_barrier(okl_ctx* ctx) {
  double v = 0.0;
  int s = okl_ctx_sync(ctx);
  if (s != OKL_OK) return s;
  return okl_allreduce_max_host(ctx, &v);
}

// ------------------------------------------------------------------------------------------------ debug timeline (not in okl.h)
// A one-thread kernel after every launch records %globaltimer, in stream order and inside captured graphs alike, so the
// effective cost of every link of a step's kernel chain (launch gap + execution) can be read back.  Perturbs timing by the
// stamp kernels' own latency (calibrated by the two back-to-back stamps at the start).
namespace {
__global__ void timeline_stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
}  // namespace

void okl_timeline_stamp(okl_ctx* ctx, const char* what) {
  if (ctx->tl_n >= 4096) return;
  char nm[96];
  snprintf(nm, sizeof(nm), "%s@%d%s", what, ctx->cur_lane, ctx->capturing ? "g" : "");
  ctx->tl_names.push_back(nm);
  timeline_stamp_kernel<<<1, 1, 0, ctx->stream>>>(ctx->tl_buf + ctx->tl_n++);
}

extern "C" int okl_debug_timeline_begin(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  if (!ctx->tl_buf) {
    OKL_CHECK_CUDA(cudaMalloc(&ctx->tl_buf, 4096 * sizeof(unsigned long long)));
    OKL_CHECK_CUDA(cudaMemset(ctx->tl_buf, 0, 4096 * sizeof(unsigned long long)));
  }
  ctx->tl_n = 0;
  ctx->tl_names.clear();
  okl_timeline_stamp(ctx, "start0");
  okl_timeline_stamp(ctx, "start1");
  return OKL_OK;
}

// copies the stamps (ns) to out[0..n) and the '\n'-separated labels to names; stops recording
extern "C" int okl_debug_timeline_read(okl_ctx* ctx, unsigned long long* out, int cap, char* names, int names_cap, int* n) {
  OKL_REQUIRE(ctx && out && names && n && ctx->tl_buf, "okl_debug_timeline_read: no timeline");
  OKL_CHECK_CUDA(cudaDeviceSynchronize());
  const int cnt = ctx->tl_n < cap ? ctx->tl_n : cap;
  OKL_CHECK_CUDA(cudaMemcpy(out, ctx->tl_buf, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::string all;
  for (int i = 0; i < cnt; i++) { all += ctx->tl_names[i]; all += '\n'; }
  snprintf(names, names_cap, "%s", all.c_str());
  *n = cnt;
  cudaFree(ctx->tl_buf);
  ctx->tl_buf = nullptr;
  return OKL_OK;
}
