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
constexpr int LL_MAX_FLOATS = 16384;          // one-shot packet exchange up to 64 KB (a single small conv layer: 49 KB)
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
__global__ void __launch_bounds__(1024) p2p_ll_allreduce_kernel(const P2PPeers peers, const int rank, const int nranks, float* mine,
                                                                const int n) {
  P2PHeader* me = peers.hdr[rank];
  const unsigned s = me->ll_seq + 1, par = s & 1u;
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


// ------------------------------------------------------------------------------------------------ NVLS: in-switch reduction
// The gradient arena can live in a region that every rank binds to ONE multicast object (cuMulticastCreate / cuMulticastBindMem,
// driver API).  Through the multicast address, `multimem.ld_reduce` returns the sum of all ranks' copies of a 16-byte word - the
// NVSwitch adds them on the way - and `multimem.st` writes a word into every rank's copy.  The all-reduce is then ONE pass over
// 1/R of the bucket per rank (R x less NVLink traffic than the two-shot kernel's R peer loads + R peer stores per word, and a single
// round trip); the ready / done flags are the same device-side sequence-number barriers as above (in the cudaIpc header).
__device__ __forceinline__ float4 mm_ld_reduce_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f32x4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) mc_allreduce_kernel(const P2PPeers peers, const int rank, const int nranks, float* __restrict__ mc,
                                                                    const size_t n4) {
  P2PHeader* me = peers.hdr[rank];
  const unsigned s = me->seq + 1;
  if (blockIdx.x == 0 && threadIdx.x < nranks) st_release_sys(&peers.hdr[threadIdx.x]->ready[rank], s);
  if (threadIdx.x < nranks) spin_until(&me->ready[threadIdx.x], s);
  __syncthreads();
  const size_t chunk = (n4 + nranks - 1) / nranks;
  const size_t beg = (size_t)rank * chunk, end = min(n4, beg + chunk);
  for (size_t i = beg + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = mm_ld_reduce_f32x4(mc + 4 * i);       // sum over all ranks, added inside the switch
    mm_st_f32x4(mc + 4 * i, v);                            // ... and broadcast to every rank's copy
  }
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


// The driver API is resolved at run time (cudaGetDriverEntryPoint): the library must load on a machine without libcuda.so.1
// (CPU-only build / test container), exactly like cuTensorMapEncodeTiled in the tcgen05 kernels.
struct CuApi {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*CtxGetDevice)(CUdevice*) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  bool ok = false;
};

CuApi& cu_api() {
  static CuApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  bool all = true;
  auto get = [&](const char* name, void** out) {
    cudaDriverEntryPointQueryResult qr;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) { all = false; cudaGetLastError(); return; }
    *out = p;
  };
  get("cuGetErrorString", (void**)&api.GetErrorString);
  get("cuCtxGetDevice", (void**)&api.CtxGetDevice);
  get("cuDeviceGetAttribute", (void**)&api.DeviceGetAttribute);
  get("cuMulticastGetGranularity", (void**)&api.MulticastGetGranularity);
  get("cuMulticastCreate", (void**)&api.MulticastCreate);
  get("cuMulticastAddDevice", (void**)&api.MulticastAddDevice);
  get("cuMulticastBindMem", (void**)&api.MulticastBindMem);
  get("cuMemExportToShareableHandle", (void**)&api.MemExportToShareableHandle);
  get("cuMemImportFromShareableHandle", (void**)&api.MemImportFromShareableHandle);
  get("cuMemCreate", (void**)&api.MemCreate);
  get("cuMemGetAllocationGranularity", (void**)&api.MemGetAllocationGranularity);
  get("cuMemAddressReserve", (void**)&api.MemAddressReserve);
  get("cuMemMap", (void**)&api.MemMap);
  get("cuMemSetAccess", (void**)&api.MemSetAccess);
  api.ok = all;
  return api;
}

struct McState {
  CUmemGenericAllocationHandle mc = 0, mem = 0;
  size_t size = 0, bump = 0;
  CUdeviceptr uc = 0, mcva = 0;
  bool added = false, ready = false;
};
std::map<okl_ctx*, McState> g_mc;

#define OKL_CHECK_CU(expr)                                                                            \
  do {                                                                                                \
    CUresult _r = (expr);                                                                             \
    if (_r != CUDA_SUCCESS) {                                                                         \
      const char* _m = nullptr;                                                                       \
      if (cu_api().GetErrorString) cu_api().GetErrorString(_r, &_m);                                                                    \
      okl_set_error("%s failed: %s (CUresult %d)", #expr, _m ? _m : "?", (int)_r);                    \
      return OKL_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

int mc_props(okl_ctx* ctx, size_t bytes, CUmulticastObjectProp& prop, size_t& size) {
  memset(&prop, 0, sizeof(prop));
  prop.numDevices = (unsigned)ctx->nranks;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  prop.flags = 0;
  prop.size = bytes;
  size_t gran = 0;
  OKL_CHECK_CU(cu_api().MulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  if (gran == 0) gran = 2u << 20;
  size = (bytes + gran - 1) / gran * gran;
  prop.size = size;
  return OKL_OK;
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
static bool mc_contains(okl_ctx* ctx, const void* p, size_t count) {
  auto it = g_mc.find(ctx);
  if (it == g_mc.end() || !it->second.ready) return false;
  const char* lo = (const char*)it->second.uc;
  const char* q = (const char*)p;
  return q >= lo && q + count * 4 <= lo + it->second.size && ((q - lo) & 15) == 0 && (count & 3) == 0;
}

bool okl_p2p_applicable(okl_ctx* ctx, const void* p, size_t count) {
  auto it = g_p2p.find(ctx);
  if (it == g_p2p.end() || !it->second.base || it->second.nimported != ctx->nranks) return false;
  if (mc_contains(ctx, p, count)) return true;            // multicast-bound arena: NVLS kernel (flags still live in the cudaIpc header)
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
    const int npk = (int)(count / 2);
    int threads = (npk + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 32) threads = 32;
    p2p_ll_allreduce_kernel<<<1, threads, 0, ctx->stream>>>(peers, ctx->rank, ctx->nranks, (float*)p, (int)count);
    OKL_LAUNCHED(ctx, "p2p_ll_allreduce");
    return OKL_OK;
  }
  if (mc_contains(ctx, p, count)) {                      // NVLS: multimem.ld_reduce + multimem.st on the multicast mapping
    const McState& mc = g_mc[ctx];
    float* mcp = (float*)(mc.mcva + ((char*)p - (char*)mc.uc));
    const size_t n4 = count / 4, chunk = (n4 + ctx->nranks - 1) / ctx->nranks;
    if (ctx->comm_p2p_kind == 1 && ctx->stream == ctx->comm_stream && chunk <= (size_t)8 * 128 * ctx->sm_count) {
      int grid = (int)((chunk + 127) / 128);
      if (grid > ctx->sm_count) grid = ctx->sm_count;
      if (grid < 1) grid = 1;
      mc_allreduce_kernel<128, 16><<<grid, 128, 0, ctx->stream>>>(peers, ctx->rank, ctx->nranks, mcp, n4);
    } else {
      int grid = (int)((chunk + 511) / 512);
      if (grid > 2 * ctx->sm_count) grid = 2 * ctx->sm_count;
      if (grid < 1) grid = 1;
      mc_allreduce_kernel<512, 1><<<grid, 512, 0, ctx->stream>>>(peers, ctx->rank, ctx->nranks, mcp, n4);
    }
    OKL_LAUNCHED(ctx, "mc_allreduce");
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


// ------------------------------------------------------------------------------------------------ NVLS multicast region (C ABI)
// Bring-up (okrolearn_b200/dist.py): rank 0 okl_mc_create -> POSIX fd -> sent to the other ranks over a Unix socket (SCM_RIGHTS) ->
// okl_mc_import; every rank okl_mc_add_device; barrier; every rank okl_mc_bind (physical memory, bind, unicast + multicast mappings).
extern "C" int okl_mc_supported(okl_ctx* ctx) {
  if (!ctx || !cu_api().ok) return 0;
  CUdevice dev;
  if (cu_api().CtxGetDevice(&dev) != CUDA_SUCCESS) {
    cudaSetDevice(ctx->device);
    cudaFree(0);
    if (cu_api().CtxGetDevice(&dev) != CUDA_SUCCESS) return 0;
  }
  int v = 0;
  if (cu_api().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
  return v ? 1 : 0;
}

extern "C" int okl_mc_create(okl_ctx* ctx, size_t bytes, int* out_fd) {
  OKL_REQUIRE(ctx && out_fd && bytes > 0, "bad argument");
  OKL_REQUIRE(cu_api().ok, "okl_mc_create: the CUDA driver lacks the multicast API");
  McState& st = g_mc[ctx];
  OKL_REQUIRE(st.mc == 0, "okl_mc_create: already created");
  OKL_CHECK_CUDA(cudaSetDevice(ctx->device));
  CUmulticastObjectProp prop;
  int s = mc_props(ctx, bytes, prop, st.size);
  if (s != OKL_OK) return s;
  OKL_CHECK_CU(cu_api().MulticastCreate(&st.mc, &prop));
  int fd = -1;
  OKL_CHECK_CU(cu_api().MemExportToShareableHandle(&fd, st.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *out_fd = fd;
  return OKL_OK;
}

extern "C" int okl_mc_import(okl_ctx* ctx, int fd, size_t bytes) {
  OKL_REQUIRE(ctx && fd >= 0 && bytes > 0, "bad argument");
  OKL_REQUIRE(cu_api().ok, "okl_mc_import: the CUDA driver lacks the multicast API");
  McState& st = g_mc[ctx];
  OKL_REQUIRE(st.mc == 0, "okl_mc_import: already created");
  OKL_CHECK_CUDA(cudaSetDevice(ctx->device));
  CUmulticastObjectProp prop;
  int s = mc_props(ctx, bytes, prop, st.size);
  if (s != OKL_OK) return s;
  OKL_CHECK_CU(cu_api().MemImportFromShareableHandle(&st.mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  return OKL_OK;
}

extern "C" int okl_mc_add_device(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  McState& st = g_mc[ctx];
  OKL_REQUIRE(st.mc != 0, "okl_mc_add_device: no multicast object");
  CUdevice dev;
  OKL_CHECK_CU(cu_api().CtxGetDevice(&dev));
  OKL_CHECK_CU(cu_api().MulticastAddDevice(st.mc, dev));
  st.added = true;
  return OKL_OK;
}

extern "C" int okl_mc_bind(okl_ctx* ctx) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  McState& st = g_mc[ctx];
  OKL_REQUIRE(st.mc != 0 && st.added, "okl_mc_bind: call okl_mc_add_device on every rank first");
  CUdevice dev;
  OKL_CHECK_CU(cu_api().CtxGetDevice(&dev));
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  OKL_CHECK_CU(cu_api().MemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  OKL_REQUIRE(gran == 0 || st.size % gran == 0, "okl_mc_bind: multicast size %zu is not a multiple of the allocation granularity %zu", st.size, gran);
  OKL_CHECK_CU(cu_api().MemCreate(&st.mem, st.size, &ap, 0));
  OKL_CHECK_CU(cu_api().MulticastBindMem(st.mc, 0, st.mem, 0, st.size, 0));
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = dev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  OKL_CHECK_CU(cu_api().MemAddressReserve(&st.uc, st.size, gran, 0, 0));
  OKL_CHECK_CU(cu_api().MemMap(st.uc, st.size, 0, st.mem, 0));
  OKL_CHECK_CU(cu_api().MemSetAccess(st.uc, st.size, &ad, 1));
  OKL_CHECK_CU(cu_api().MemAddressReserve(&st.mcva, st.size, gran, 0, 0));
  OKL_CHECK_CU(cu_api().MemMap(st.mcva, st.size, 0, st.mc, 0));
  OKL_CHECK_CU(cu_api().MemSetAccess(st.mcva, st.size, &ad, 1));
  OKL_CHECK_CUDA(cudaMemset((void*)st.uc, 0, st.size));
  OKL_CHECK_CUDA(cudaDeviceSynchronize());
  st.ready = true;
  return OKL_OK;
}

extern "C" int okl_mc_suballoc(okl_ctx* ctx, size_t bytes, void** out) {
  OKL_REQUIRE(ctx && out, "NULL argument");
  McState& st = g_mc[ctx];
  OKL_REQUIRE(st.ready, "okl_mc_suballoc: no multicast region");
  size_t b = (bytes + 255) & ~size_t(255);
  if (st.bump + b > st.size) {
    okl_set_error("okl_mc_suballoc: multicast region exhausted (%zu + %zu > %zu)", st.bump, b, st.size);
    return OKL_ERR_OOM;
  }
  *out = (void*)(st.uc + st.bump);
  st.bump += b;
  return OKL_OK;
}
