// Latency-optimised gradient all-reduce over NVLink peer memory (one process per GPU, one node).
//
// For the small gradient buckets of a CNN (a few MB) an NCCL all-reduce is dominated by launch / protocol latency, not by the
// 900 GB/s links.  Here every rank owns a *symmetric* buffer (plain cudaMalloc, exported with cudaIpcGetMemHandle and mapped by all
// peers through NVSwitch), the network's gradient arena lives inside it, and ONE kernel per step does a two-shot all-reduce
// directly on peer memory:
//     ready barrier  : each rank raises its flag in every peer's buffer (release.sys) and polls its OWN copy of the flags
//     reduce-scatter : rank r sums slice r over all ranks' buffers in fixed rank order (bit-identical result on every rank) ...
//     all-gather     : ... and stores the sum straight into slice r of every rank's buffer (128-bit remote stores)
//     done barrier   : the last CTA of each rank raises the done flag everywhere; the kernel ends once all peers have done so
// Sequence numbers live on the device (the kernel bumps them), so the kernel is replayable from a CUDA graph.  All spins are
// bounded (trap after ~10 s) so a missing peer cannot wedge a GPU.
#include "common.cuh"

namespace {

constexpr int P2P_MAX_RANKS = 8;
constexpr size_t P2P_FLAGS_BYTES = 4096;    // flags + counters at the start of the symmetric buffer
// one-shot "LL" exchange for tiny buckets (NCCL's low-latency protocol): a 16-byte packet carries two floats and two copies of the
// sequence number and is written with ONE store, so the receiver needs no separate flag, fence or barrier - it polls the packet
// itself.  Latency = one one-way NVLink trip instead of the four of the two-shot kernel (ready, pull, push-ack, done).
constexpr int LL_MAX_FLOATS = 4096;
constexpr int LL_PKTS = LL_MAX_FLOATS / 2;
constexpr size_t LL_BYTES = (size_t)2 * P2P_MAX_RANKS * LL_PKTS * 16;   // [parity][source rank][packet]
constexpr size_t P2P_HDR_BYTES = P2P_FLAGS_BYTES + LL_BYTES;

struct P2PHeader {                          // lives at the start of every rank's symmetric buffer
  unsigned ready[P2P_MAX_RANKS];            // ready[p] = last sequence number for which rank p's gradients are complete
  unsigned done[P2P_MAX_RANKS];             // done[p]  = last sequence number for which rank p has written its slice everywhere
  unsigned seq;                             // sequence number of the last completed all-reduce on this rank
  unsigned cta_count;                       // CTAs of the running kernel that have finished their writes
  unsigned ll_seq;                          // sequence number of the last completed LL all-reduce on this rank
};

struct P2PPeers {
  float* data[P2P_MAX_RANKS];               // data region of every rank (own rank included), same layout everywhere
  P2PHeader* hdr[P2P_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned target) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - target) < 0) {
    if (clock64() - t0 > 20000000000LL) {
      printf("okl p2p all-reduce: timed out waiting for a peer (flag %p target %u)\n", flag, target);
      __trap();
    }
  }
}

// THREADS = 512: the exchange has the GPU to itself.  THREADS = 128 with <= 32 registers per thread: the overlapped variant
// (okl_allreduce_fork) - one such CTA fits into the register-file space a resident wave of the weight-gradient kernels leaves
// free on every SM, so the exchange really runs next to them instead of queueing behind them.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) p2p_allreduce_kernel(const P2PPeers peers, const int rank, const int nranks, const size_t off4,
                                                                     const size_t n4) {
  P2PHeader* me = peers.hdr[rank];
  const unsigned s = me->seq + 1;           // every CTA reads it before the last CTA bumps it at the very end
  // ---- ready barrier: my gradients are complete (stream order); tell everyone, then wait for everyone
  if (blockIdx.x == 0 && threadIdx.x < nranks) st_release_sys(&peers.hdr[threadIdx.x]->ready[rank], s);
  if (threadIdx.x < nranks) spin_until(&me->ready[threadIdx.x], s);
  __syncthreads();
  // ---- reduce-scatter + all-gather on my slice, 128-bit accesses, fixed summation order (rank 0, 1, 2, ...), at most four
  // peer loads in flight per thread (16 registers)
  const size_t chunk = (n4 + nranks - 1) / nranks;
  const size_t beg = (size_t)rank * chunk, end = min(n4, beg + chunk);
  for (size_t i = beg + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p0 = 0; p0 < P2P_MAX_RANKS; p0 += 4) {
      if (p0 < nranks) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (p0 + q < nranks) v[q] = reinterpret_cast<const float4*>(peers.data[p0 + q])[off4 + i];
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (p0 + q < nranks) {
            if (p0 + q == 0) a = v[q];
            else { a.x += v[q].x; a.y += v[q].y; a.z += v[q].z; a.w += v[q].w; }
          }
      }
    }
#pragma unroll
    for (int p = 0; p < P2P_MAX_RANKS; p++)
      if (p < nranks) reinterpret_cast<float4*>(peers.data[p])[off4 + i] = a;
  }
  // ---- done barrier
  __threadfence_system();
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(&me->cta_count, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < nranks) st_release_sys(&peers.hdr[threadIdx.x]->done[rank], s);
    if (threadIdx.x < nranks) spin_until(&me->done[threadIdx.x], s);
    __syncthreads();
    if (threadIdx.x == 0) {
      me->cta_count = 0;
      me->seq = s;
      __threadfence();
    }
  }
}

__device__ __forceinline__ uint4* ll_slot(P2PHeader* h, unsigned par, int src, int i) {
  return reinterpret_cast<uint4*>(reinterpret_cast<char*>(h) + P2P_FLAGS_BYTES) + ((size_t)(par * P2P_MAX_RANKS + src) * LL_PKTS + i);
}

// One CTA.  Every rank pushes its whole vector, as packets, into slot [parity][rank] of every peer, then sums the R contributions
// in rank order (its own from the gradient buffer, the others polled out of its own staging area) - the same order on every
// rank, so all ranks end with identical bits.  Staging alternates between two halves by sequence parity: a rank can only start
// all-reduce s+2 after s+1 completed, which needed every peer's s+1 packets, which they sent after finishing s - so nobody still
// reads the half that s+2 overwrites.
__global__ void __launch_bounds__(1024) p2p_ll_allreduce_kernel(const P2PPeers peers, const int rank, const int nranks, const size_t off,
                                                                const int n) {
  P2PHeader* me = peers.hdr[rank];
  const unsigned s = me->ll_seq + 1, par = s & 1u;
  float* mine = peers.data[rank] + off;
  const int npk = n >> 1;
  for (int i = threadIdx.x; i < npk; i += blockDim.x) {
    const float2 v = *reinterpret_cast<const float2*>(mine + 2 * i);
    const uint4 pk = make_uint4(__float_as_uint(v.x), s, __float_as_uint(v.y), s);
#pragma unroll
    for (int p = 0; p < P2P_MAX_RANKS; p++)
      if (p < nranks && p != rank) {
        uint4* dst = ll_slot(peers.hdr[p], par, rank, i);
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
      }
  }
  for (int i = threadIdx.x; i < npk; i += blockDim.x) {
    const float2 v = *reinterpret_cast<const float2*>(mine + 2 * i);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; r++)
      if (r < nranks) {
        float2 c = v;
        if (r != rank) {
          const uint4* src = ll_slot(me, par, r, i);
          uint4 q;
          const long long t0 = clock64();
          do {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(src) : "memory");
            if (clock64() - t0 > 20000000000LL) {
              printf("okl p2p LL all-reduce: timed out waiting for rank %d (seq %u)\n", r, s);
              __trap();
            }
          } while (q.y != s || q.w != s);
          c = make_float2(__uint_as_float(q.x), __uint_as_float(q.z));
        }
        if (r == 0) acc = c; else { acc.x += c.x; acc.y += c.y; }
      }
    *reinterpret_cast<float2*>(mine + 2 * i) = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) me->ll_seq = s;
}

struct P2PState {
  void* base = nullptr;                     // local symmetric buffer (header + data)
  size_t bytes = 0, bump = 0;
  void* peer_base[P2P_MAX_RANKS] = {};
  bool imported[P2P_MAX_RANKS] = {};
  int nimported = 0;
};

std::map<okl_ctx*, P2PState> g_p2p;

}  // namespace

// Allocate this rank's symmetric buffer (`bytes` of data) - call once, before okl_p2p_export.
extern "C" int okl_p2p_alloc(okl_ctx* ctx, size_t bytes) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  P2PState& st = g_p2p[ctx];
  OKL_REQUIRE(st.base == nullptr, "okl_p2p_alloc: already allocated");
  OKL_REQUIRE(ctx->nranks <= P2P_MAX_RANKS, "okl_p2p: at most %d ranks", P2P_MAX_RANKS);
  OKL_CHECK_CUDA(cudaSetDevice(ctx->device));
  st.bytes = P2P_HDR_BYTES + ((bytes + 255) & ~size_t(255));
  OKL_CHECK_CUDA(cudaMalloc(&st.base, st.bytes));
  OKL_CHECK_CUDA(cudaMemset(st.base, 0, st.bytes));
  OKL_CHECK_CUDA(cudaDeviceSynchronize());
  st.peer_base[ctx->rank] = st.base;
  st.imported[ctx->rank] = true;
  st.nimported = 1;
  return OKL_OK;
}

extern "C" int okl_p2p_export(okl_ctx* ctx, void* out64) {
  OKL_REQUIRE(ctx && out64, "NULL argument");
  P2PState& st = g_p2p[ctx];
  OKL_REQUIRE(st.base, "okl_p2p_export: call okl_p2p_alloc first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  OKL_CHECK_CUDA(cudaIpcGetMemHandle(&h, st.base));
  memcpy(out64, &h, 64);
  return OKL_OK;
}

extern "C" int okl_p2p_import(okl_ctx* ctx, int rank, const void* handle64) {
  OKL_REQUIRE(ctx && handle64, "NULL argument");
  P2PState& st = g_p2p[ctx];
  OKL_REQUIRE(st.base, "okl_p2p_import: call okl_p2p_alloc first");
  OKL_REQUIRE(rank >= 0 && rank < ctx->nranks && rank != ctx->rank, "okl_p2p_import: bad rank %d", rank);
  if (st.imported[rank]) return OKL_OK;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  OKL_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  st.peer_base[rank] = p;
  st.imported[rank] = true;
  st.nimported++;
  return OKL_OK;
}

// Sub-allocate from the local symmetric data region (same call sequence on every rank => same offsets everywhere).
extern "C" int okl_p2p_suballoc(okl_ctx* ctx, size_t bytes, void** out) {
  OKL_REQUIRE(ctx && out, "NULL argument");
  P2PState& st = g_p2p[ctx];
  OKL_REQUIRE(st.base, "okl_p2p_suballoc: no symmetric buffer");
  size_t b = (bytes + 255) & ~size_t(255);
  if (P2P_HDR_BYTES + st.bump + b > st.bytes) {
    okl_set_error("okl_p2p_suballoc: symmetric buffer exhausted (%zu + %zu > %zu)", st.bump, b, st.bytes - P2P_HDR_BYTES);
    return OKL_ERR_OOM;
  }
  *out = (char*)st.base + P2P_HDR_BYTES + st.bump;
  st.bump += b;
  return OKL_OK;
}

// 1 if [p, p + count floats) lies in the symmetric region, every peer is mapped and the peer-memory kernel applies.
bool okl_p2p_applicable(okl_ctx* ctx, const void* p, size_t count) {
  auto it = g_p2p.find(ctx);
  if (it == g_p2p.end() || !it->second.base || it->second.nimported != ctx->nranks) return false;
  const P2PState& st = it->second;
  const char* lo = (const char*)st.base + P2P_HDR_BYTES;
  const char* q = (const char*)p;
  return q >= lo && q + count * 4 <= (const char*)st.base + st.bytes && ((q - lo) & 15) == 0 && (count & 3) == 0;
}

// tiny buckets use the packet kernel, which has its own sequence counter and may run concurrently with a two-shot exchange
bool okl_p2p_is_ll(size_t count) { return count <= (size_t)LL_MAX_FLOATS; }

int okl_p2p_allreduce(okl_ctx* ctx, void* p, size_t count) {
  P2PState& st = g_p2p[ctx];
  P2PPeers peers;
  memset(&peers, 0, sizeof(peers));
  for (int r = 0; r < ctx->nranks; r++) {
    peers.hdr[r] = (P2PHeader*)st.peer_base[r];
    peers.data[r] = (float*)((char*)st.peer_base[r] + P2P_HDR_BYTES);
  }
  if (count <= (size_t)LL_MAX_FLOATS) {                  // tiny bucket: one-shot packet exchange, one CTA
    const size_t off = ((char*)p - ((char*)st.base + P2P_HDR_BYTES)) / 4;
    const int npk = (int)(count / 2);
    int threads = (npk + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 32) threads = 32;
    p2p_ll_allreduce_kernel<<<1, threads, 0, ctx->stream>>>(peers, ctx->rank, ctx->nranks, off, (int)count);
    OKL_LAUNCHED(ctx, "p2p_ll_allreduce");
    return OKL_OK;
  }
  const size_t off4 = ((char*)p - ((char*)st.base + P2P_HDR_BYTES)) / 16, n4 = count / 4;
  const size_t chunk = (n4 + ctx->nranks - 1) / ctx->nranks;
  static int small_ok = -1;
  if (small_ok < 0) { const char* e = getenv("OKL_P2P_SMALL"); small_ok = e ? atoi(e) : 1; }
  // overlapped with the remaining backward and at most ~4 float4 per thread: the small-footprint variant
  if (small_ok && ctx->comm_p2p_kind == 1 && ctx->stream == ctx->comm_stream && chunk <= (size_t)4 * 128 * ctx->sm_count) {
    int grid = (int)((chunk + 127) / 128);
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    if (grid < 1) grid = 1;
    p2p_allreduce_kernel<128, 16><<<grid, 128, 0, ctx->stream>>>(peers, ctx->rank, ctx->nranks, off4, n4);
    OKL_LAUNCHED(ctx, "p2p_allreduce");
    return OKL_OK;
  }
  int grid = (int)((chunk + 511) / 512);                 // one float4 per thread: the exchange is latency-, not bandwidth-bound
  if (grid > 2 * ctx->sm_count) grid = 2 * ctx->sm_count;
  if (grid < 1) grid = 1;
  p2p_allreduce_kernel<512, 1><<<grid, 512, 0, ctx->stream>>>(peers, ctx->rank, ctx->nranks, off4, n4);
  OKL_LAUNCHED(ctx, "p2p_allreduce");
  return OKL_OK;
}
