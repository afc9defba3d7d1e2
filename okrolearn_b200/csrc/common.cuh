// Internal shared declarations for libokl_b200 (context, error plumbing, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: a no-op unless a tool (nsys, ncu --nvtx) injects its library

#include "../../include/okl.h"

struct okl_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  unsigned long long launches = 0;
};

#define OKL_FLOW_DEPTH 8
struct okl_ctx {
  int device = 0;
  int precision = OKL_FP32;
  int sm_count = 148;
  int cc = 100;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;        // compute
  cudaStream_t comm_stream = nullptr;   // NCCL
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  unsigned long long launches = 0;      // kernels launched eagerly + graph replays
  bool capturing = false;
  unsigned long long capture_launches = 0;
  // caching device allocator
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  std::map<void*, size_t> live;         // every block handed out or cached -> rounded size
  // scratch
  void* workspace = nullptr;            // split-K partials etc.
  size_t workspace_bytes = 0;
  void* staging = nullptr;              // dtype conversion staging for h2d / d2h
  size_t staging_bytes = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  void* tc_dbg = nullptr;               // debug: timestamp buffer for the tcgen05 GEMM
  int* tile_counters = nullptr;         // split-K arrival counters of the tcgen05 GEMM (all zero between launches)
  void* bf16_scratch = nullptr;         // bf16 copies of GEMM operands (OKL_BF16 mode)
  void* head_scratch = nullptr;         // arrival counter + per-block loss partials of the fused classifier head
  void* leaf_scratch = nullptr;         // ring of partial sums handed from a main-lane kernel to a second stage on an auxiliary lane
  size_t leaf_off = 0;
  size_t bf16_scratch_bytes = 0;
  // auxiliary lanes: independent branches of a step (weight / bias gradients are leaves of the backward chain) run on their own
  // streams with their own scratch; the fields above (stream, workspace, tile_counters, bf16_scratch) always describe the CURRENT lane
  struct Lane {
    cudaStream_t stream = nullptr;
    void* workspace = nullptr; size_t workspace_bytes = 0;
    int* tile_counters = nullptr;
    void* bf16_scratch = nullptr; size_t bf16_scratch_bytes = 0;
    cudaEvent_t ev = nullptr;
    bool dirty = false;
  };
  Lane lanes[3];
  int cur_lane = 0;
  cudaEvent_t ev_mark = nullptr;
  bool fork_marked = false;
  bool aux_dirty = false;
  std::vector<void*> deferred_free;     // pool blocks released while an aux lane may still read them
  // comm
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  bool comm_pending = false;
  int comm_p2p_kind = 0;                // peer-memory exchange in flight on the comm stream: 0 none, 1 two-shot, 2 packet (LL)
  unsigned long long* tl_buf = nullptr; // debug timeline: one globaltimer stamp after every kernel (okl_debug_timeline_*)
  int tl_n = 0;
  std::vector<std::string> tl_names;
  // double-buffered batch prefetch + lagged loss read-back (okl_batch_prefetch / okl_batch_wait / okl_loss_fetch_*)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy_fork = nullptr, ev_slot[2] = {nullptr, nullptr}, ev_loss[2] = {nullptr, nullptr};
  float* loss_pinned = nullptr;         // 2 slots x {value, sequence number}
  unsigned loss_seq[2] = {0, 0};
  cudaEvent_t ev_flow[8] = {};
  unsigned long long flow_next = 0;
  // stream mode: host copy of the next batch's row indices (set by okl_batch_host_index, consumed by the next okl_batch_prefetch) and
  // the page-locked staging ring the host gathers the rows into ([inputs / targets][batch the host may run ahead])
  const int* host_idx_next = nullptr;
  void* stage[2][OKL_FLOW_DEPTH] = {}; size_t stage_bytes[2][OKL_FLOW_DEPTH] = {};
  int pdl = 0;                          // programmatic dependent launch: 0 off (default: measured no gain), 1 on except while an aux lane is forked, 2 always
};

void okl_set_error(const char* fmt, ...);

#define OKL_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      okl_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return OKL_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define OKL_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      okl_set_error(__VA_ARGS__);       \
      return OKL_ERR_INVALID;           \
    }                                   \
  } while (0)

// account one kernel launch and surface launch-configuration errors immediately
void okl_timeline_stamp(okl_ctx* ctx, const char* what);
// NVTX: every kernel enqueue leaves a named marker on the host timeline (OKL_NVTX=0 turns it off); the C-ABI entry points of a
// train step additionally open ranges (okl_graph_launch: "okl:step").  SURVEY.md section 5 (tracing).
static inline bool okl_nvtx_on() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("OKL_NVTX"); on = (e && *e == '0') ? 0 : 1; }
  return on != 0;
}
static inline int okl_post_launch(okl_ctx* ctx, const char* what) {
  if (okl_nvtx_on()) nvtxMarkA(what);
  if (ctx->capturing) ctx->capture_launches++; else ctx->launches++;
  if (ctx->tl_buf) okl_timeline_stamp(ctx, what);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    okl_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return OKL_ERR_CUDA;
  }
  return OKL_OK;
}
#define OKL_LAUNCHED(ctx, what)                          \
  do {                                                   \
    int _s = okl_post_launch(ctx, what);                 \
    if (_s != OKL_OK) return _s;                         \
  } while (0)

// ---- programmatic dependent launch (PDL).  A step is a chain of ~10 short kernels; with a plain stream / graph edge each
// boundary costs the full drain -> launch -> ramp latency.  Kernels on the chain call okl_pdl_sync() before touching global
// memory: it waits until the predecessor grid has completed and flushed (a no-op when the launch did not carry the attribute),
// then lets the successor grid start launching, so the successor's CTAs are already resident when this grid drains.  The wait
// comes first on purpose: a successor can then only start once its predecessor has passed its own wait, so completion is
// transitive along the chain and everything before the wait needs no data produced by earlier kernels.
#ifdef __CUDACC__
__device__ __forceinline__ void okl_pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline void okl_launch(okl_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  // while weight-gradient kernels run on an auxiliary lane, early-resident CTAs of the next main-lane kernel would squat on the
  // SMs those kernels need - keep plain edges there
  const bool pdl = ctx->pdl == 2 || (ctx->pdl == 1 && (ctx->cur_lane != 0 || !ctx->aux_dirty));
  cfg.attrs = &at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);      // errors surface through OKL_LAUNCHED
}
#endif

int okl_ensure_workspace(okl_ctx* ctx, size_t bytes);   // grows ctx->workspace (not during capture)
int okl_ensure_staging(okl_ctx* ctx, size_t bytes);
int okl_loss_post(okl_ctx* ctx, float* host_slot, const void* dev_loss, unsigned seq);

static inline int okl_grid_1d(const okl_ctx* ctx, size_t work_items, int block, int max_waves = 8) {
  size_t blocks = (work_items + block - 1) / block;
  size_t cap = (size_t)ctx->sm_count * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- entry points implemented in other translation units -------------------------------------
// FFMA (exact fp32) GEMM-shaped kernels: okl_ffma.cu
int okl_ffma_dense_fwd(okl_ctx*, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act);
int okl_ffma_dense_dw(okl_ctx*, const float* X, const float* G, float* dW, int M, int K, int N);
int okl_ffma_dense_dx(okl_ctx*, const float* G, const float* W, float* dX, int M, int K, int N, const float* relu_mask);
int okl_ffma_conv_fprop(okl_ctx*, const okl_conv_desc*, const float* x, const float* W, const float* b, float* y);
int okl_ffma_conv_dgrad(okl_ctx*, const okl_conv_desc*, const float* g, const float* W, float* dx, const float* relu_mask);
int okl_ffma_conv_wgrad(okl_ctx*, const okl_conv_desc*, const float* x, const float* g, float* dW);
// reductions: okl_elementwise.cu
int okl_colsum(okl_ctx*, const float* G, float* out, long long rows, int cols);                       // out[c] = sum_r G[r,c]
int okl_channel_sum(okl_ctx*, const float* g, float* out, int N, int C, long long S, int layout);    // out[c] = sum_{n,s}
// tcgen05 tensor-core kernels (TF32 / BF16 operands, fp32 accumulate in TMEM): okl_tc_gemm.cu / okl_tc_conv.cu
bool okl_tc_dense_supported(okl_ctx*, int M, int K, int N);
int okl_tc_dense_fwd(okl_ctx*, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act);
int okl_tc_dense_dw(okl_ctx*, const float* X, const float* G, float* dW, int M, int K, int N);
int okl_tc_dense_dx(okl_ctx*, const float* G, const float* W, float* dX, int M, int K, int N, const float* relu_mask);
bool okl_tc_conv_supported(okl_ctx*, const okl_conv_desc*);
bool okl_tc_conv_smallc_supported(okl_ctx*, const okl_conv_desc*);
int okl_tc_conv_smallc_fprop(okl_ctx*, const okl_conv_desc*, const float* x, const float* W, const float* b, float* y);
int okl_tc_conv_smallc_wgrad(okl_ctx*, const okl_conv_desc*, const float* x, const float* g, float* dW, float* db);
int okl_tc_conv_fprop(okl_ctx*, const okl_conv_desc*, const float* x, const float* W, const float* b, float* y);
int okl_tc_conv_dgrad(okl_ctx*, const okl_conv_desc*, const float* g, const float* W, float* dx, const float* relu_mask);
int okl_tc_conv_wgrad(okl_ctx*, const okl_conv_desc*, const float* x, const float* g, float* dW);
// direct fp32 kernels for tiny reduction dims (first layers): okl_conv_small.cu
bool okl_conv_small_fprop_ok(const okl_conv_desc*);
bool okl_conv_small_wgrad_ok(const okl_conv_desc*);
int okl_conv_small_fprop(okl_ctx*, const okl_conv_desc*, const float* x, const float* W, const float* b, float* y);
int okl_conv_small_wgrad(okl_ctx*, const okl_conv_desc*, const float* x, const float* g, float* dW, float* db);
// one-kernel forward (+softmax) and one-kernel backward for tiny output widths (classifier heads): okl_dense_small.cu
bool okl_dense_small_ok(int M, int K, int N);
int okl_dense_small_fwd(okl_ctx*, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act);
int okl_dense_small_bwd(okl_ctx*, const float* X, const float* W, const float* G, float* dX, float* dW, float* db, const float* relu_mask,
                        int M, int K, int N);
// peer-memory (NVLink) all-reduce for small buckets: okl_p2p.cu
bool okl_p2p_applicable(okl_ctx*, const void* p, size_t count);
int okl_p2p_allreduce(okl_ctx*, void* p, size_t count);
bool okl_p2p_is_ll(size_t count);
