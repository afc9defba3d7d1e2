// DenseLayer kernels for a tiny output width (N <= 16: the classifier head, e.g. 128 -> 10).  A 64x64 GEMM tile would be >= 75 %
// padding and split-K + reduce costs two launches per product, while the whole problem is a few hundred kFLOP: here the forward
// (+ bias, optionally + the SoftmaxActivationLayer that follows, okl.py:364-373) is ONE kernel and the whole backward
// (dW = X^T G, db = sum_rows G, dX = G W^T with the preceding ReLU's mask, okl.py:35-37) is ONE kernel.  Exact fp32.
#include "common.cuh"

namespace {

constexpr int SN_MAXN = 16;
constexpr int SN_MC = 128;                 // rows of X / G staged per pass of the dW blocks

__device__ __forceinline__ float sn_act(float v, int act) {
  switch (act) {
    case OKL_ACT_RELU: return v > 0.f ? v : 0.f;
    case OKL_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case OKL_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// one warp per row: lanes stride over k, N running sums per lane, butterfly reduce, then bias + activation / softmax
__global__ void __launch_bounds__(256) dense_smalln_fwd(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                                                       float* __restrict__ Y, int M, int K, int N, int act) {
  extern __shared__ float Ws[];            // [K][N]
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) Ws[i] = __ldg(W + i);
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int m = warp; m < M; m += nwarps) {
    float acc[SN_MAXN];
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) acc[n] = 0.f;
    const float* xr = X + (size_t)m * K;
    for (int k = lane; k < K; k += 32) {
      const float xv = __ldg(xr + k);
      const float* wr = Ws + k * N;
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++)
        if (n < N) acc[n] = fmaf(xv, wr[n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) {
      if (n < N) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], s);
        if (b) acc[n] += __ldg(b + n);
      }
    }
    if (act == 4) {                          // fused stable softmax over the N outputs (+ nan_to_num, okl.py:364-370)
      float mx = -INFINITY, sum = 0.f;
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++)
        if (n < N) mx = fmaxf(mx, acc[n]);
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++)
        if (n < N) { acc[n] = expf(acc[n] - mx); sum += acc[n]; }
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++)
        if (n < N) { float v = acc[n] / sum; acc[n] = isnan(v) ? 0.f : v; }
    } else {
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++)
        if (n < N) acc[n] = sn_act(acc[n], act);
    }
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++)
      if (n < N && lane == n) Y[(size_t)m * N + n] = acc[n];
  }
}

// blocks [0, nbx): dX[m][k] = sum_n G[m][n] W[k][n] (* relu mask);  blocks [nbx, nbx + nbw): dW[k][n] = sum_m X[m][k] G[m][n]
// (thread = (k, n), k fastest so X reads are coalesced and G reads broadcast);  last block: db[n] = sum_m G[m][n]
__global__ void __launch_bounds__(256) dense_smalln_bwd(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ G,
                                                       float* __restrict__ dX, float* __restrict__ dW, float* __restrict__ db,
                                                       const float* __restrict__ mask, int M, int K, int N, int nbx, int nbw) {
  extern __shared__ float Ws[];            // [K][N] (dX blocks only)
  const int bid = blockIdx.x;
  if (bid < nbx) {
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) Ws[i] = __ldg(W + i);
    __syncthreads();
    const size_t total = (size_t)M * K;
    for (size_t i = (size_t)bid * blockDim.x + threadIdx.x; i < total; i += (size_t)nbx * blockDim.x) {
      const int m = (int)(i / K), k = (int)(i - (size_t)m * K);
      const float* gr = G + (size_t)m * N;
      const float* wr = Ws + k * N;
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++)
        if (n < N) s = fmaf(__ldg(gr + n), wr[n], s);
      if (mask && !(__ldg(mask + i) > 0.f)) s = 0.f;
      dX[i] = s;
    }
  } else if (bid < nbx + nbw) {
    // dW for 32 consecutive k: the (M-chunk x 32) slab of X and the matching rows of G are first pulled into shared memory with
    // every load of a thread in flight at once, then thread (lane = k, warp = n, n + 8) accumulates from shared memory.
    float* xs = Ws;                          // [SN_MC][32]
    float* gs = Ws + SN_MC * 32;             // [SN_MC][N]
    const int k0 = (bid - nbx) * 32, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a0 = 0.f, a1 = 0.f;
    for (int m0 = 0; m0 < M; m0 += SN_MC) {
      const int mc = min(SN_MC, M - m0);
      __syncthreads();
      for (int r = warp; r < mc; r += 8) xs[r * 32 + lane] = (k0 + lane < K) ? __ldg(X + (size_t)(m0 + r) * K + k0 + lane) : 0.f;
      for (int i = threadIdx.x; i < mc * N; i += blockDim.x) gs[i] = __ldg(G + (size_t)m0 * N + i);
      __syncthreads();
      const int n0 = warp, n1 = warp + 8;
      for (int r = 0; r < mc; r++) {
        const float xv = xs[r * 32 + lane];
        if (n0 < N) a0 = fmaf(xv, gs[r * N + n0], a0);
        if (n1 < N) a1 = fmaf(xv, gs[r * N + n1], a1);
      }
    }
    if (k0 + lane < K) {
      if (warp < N) dW[(size_t)(k0 + lane) * N + warp] = a0;
      if (warp + 8 < N) dW[(size_t)(k0 + lane) * N + warp + 8] = a1;
    }
  } else if (db) {
    const int n = threadIdx.x & 15, part = threadIdx.x >> 4;      // 16 partial sums per column
    float s = 0.f;
    if (n < N)
      for (int m = part; m < M; m += 16) s += __ldg(G + (size_t)m * N + n);
    // combine the 16 parts of each column in fixed order
    __shared__ float parts[16][SN_MAXN];
    parts[part][n] = s;
    __syncthreads();
    if (part == 0 && n < N) {
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 16; q++) a += parts[q][n];
      db[n] = a;
    }
  }
}

}  // namespace

bool okl_dense_small_ok(int M, int K, int N) { return N >= 1 && N <= SN_MAXN && (long long)K * N <= 12000 && M >= 1; }

int okl_dense_small_fwd(okl_ctx* ctx, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act) {
  int warps = M;
  int grid = (warps * 32 + 255) / 256;
  const int cap = ctx->sm_count * 4;
  if (grid > cap) grid = cap;
  dense_smalln_fwd<<<grid, 256, (size_t)K * N * sizeof(float), ctx->stream>>>(X, W, b, Y, M, K, N, act);
  OKL_LAUNCHED(ctx, "dense_smalln_fwd");
  return OKL_OK;
}

int okl_dense_small_bwd(okl_ctx* ctx, const float* X, const float* W, const float* G, float* dX, float* dW, float* db, const float* mask,
                        int M, int K, int N) {
  int nbx = 0;
  if (dX) {
    nbx = (int)(((size_t)M * K + 255) / 256);
    const int cap = ctx->sm_count * 4;
    if (nbx > cap) nbx = cap;
  }
  const int nbw = dW ? (K + 31) / 32 : 0;
  const int total = nbx + nbw + (db ? 1 : 0);
  if (total == 0) return OKL_OK;
  size_t smem = (size_t)K * N * sizeof(float), smem_w = (size_t)SN_MC * (32 + N) * sizeof(float);
  if (smem_w > smem) smem = smem_w;
  dense_smalln_bwd<<<total, 256, smem, ctx->stream>>>(X, W, G, dX, dW, db, mask, M, K, N, nbx, nbw);
  OKL_LAUNCHED(ctx, "dense_smalln_bwd");
  return OKL_OK;
}
