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
  okl_pdl_sync();
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
  okl_pdl_sync();
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


// Classifier head in ONE kernel (rows are independent): logits = X W + b, stable softmax, cross-entropy value, the loss gradient
// (clip(p) - onehot) / denom, its pass through the softmax Jacobian (okl.py:376-393) and dX = dZ W^T (optionally with the ReLU
// mask of the layer that produced X).  One warp per row; lane l owns k = l, l + 32, ... for BOTH products, so its KJ x N slice of
// W is loaded once into registers (all loads of a row - x, W, target - are in flight together: one memory round trip before the
// arithmetic).  Summation orders are those of dense_smalln_fwd / ce_fused_rowthread / dense_smalln_bwd, so the results match the
// unfused chain bit for bit except the loss sum, which is reduced per block and then over blocks in fixed order by the last
// block to arrive (deterministic for a given grid).
template <int KJ>
__global__ void __launch_bounds__(256) dense_softmax_ce_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                                                              const int* __restrict__ t, float* __restrict__ P, float* __restrict__ grad,
                                                              float* __restrict__ dZ, float* __restrict__ dX, const float* __restrict__ mask,
                                                              float* __restrict__ loss, float* __restrict__ loss_acc, float* __restrict__ partial,
                                                              unsigned* __restrict__ counter, int M, int K, int N, float inv_denom) {
  __shared__ float wl[8];
  __shared__ bool is_last;
  okl_pdl_sync();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + wib, nwarps = gridDim.x * 8;
  float wreg[KJ][SN_MAXN], bias[SN_MAXN];
#pragma unroll
  for (int j = 0; j < KJ; j++) {
    const int k = lane + 32 * j;
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) wreg[j][n] = (k < K && n < N) ? __ldg(W + (size_t)k * N + n) : 0.f;
  }
#pragma unroll
  for (int n = 0; n < SN_MAXN; n++) bias[n] = (b && n < N) ? __ldg(b + n) : 0.f;
  float lsum = 0.f;
  for (int m = warp; m < M; m += nwarps) {
    const float* xr = X + (size_t)m * K;
    float xv[KJ];
#pragma unroll
    for (int j = 0; j < KJ; j++) xv[j] = (lane + 32 * j < K) ? __ldg(xr + lane + 32 * j) : 0.f;
    const int tc = __ldg(t + m);
    float acc[SN_MAXN];
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) acc[n] = 0.f;
#pragma unroll
    for (int j = 0; j < KJ; j++)
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++) acc[n] = fmaf(xv[j], wreg[j][n], acc[n]);
    float mx = -INFINITY, sum = 0.f;
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) {
      if (n < N) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], s);
        if (b) acc[n] += bias[n];
        mx = fmaxf(mx, acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++)
      if (n < N) { acc[n] = expf(acc[n] - mx); sum += acc[n]; }
    float dot = 0.f, pt = 1.f, gv[SN_MAXN];
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) {
      gv[n] = 0.f;
      if (n < N) {
        float v = acc[n] / sum;
        v = isnan(v) ? 0.f : v;            // nan_to_num, okl.py:370
        acc[n] = v;
        float g = fminf(fmaxf(v, 1e-12f), 1.0f);      // np.clip(p, 1e-12, 1 - 1e-12), okl.py:1391 (upper bound rounds to 1.0f)
        if (n == tc) { pt = g; g -= 1.f; }
        g *= inv_denom;
        gv[n] = g;
        dot += g * v;
      }
    }
    lsum += -logf(pt);
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++)
      if (n < N) {
        const float dz = acc[n] * (gv[n] - dot);
        if (lane == n) {
          P[(size_t)m * N + n] = acc[n];
          grad[(size_t)m * N + n] = gv[n];
          dZ[(size_t)m * N + n] = dz;
        }
        gv[n] = dz;
      }
    if (dX) {
#pragma unroll
      for (int j = 0; j < KJ; j++) {
        const int k = lane + 32 * j;
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < SN_MAXN; n++)
          if (n < N) s = fmaf(gv[n], wreg[j][n], s);
        // the mask is the post-ReLU input itself when given (X == mask in the train loop), but honour a distinct pointer
        if (mask && !((mask == X ? xv[j] : __ldg(mask + (size_t)m * K + k)) > 0.f)) s = 0.f;
        if (k < K) dX[(size_t)m * K + k] = s;
      }
    }
  }
  if (lane == 0) wl[wib] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += wl[w];
    partial[blockIdx.x] = s;
    is_last = false;
    if (counter) {                               // no counter: head_loss_finalize_kernel sums the partials (off the critical path)
      __threadfence();
      is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (is_last && wib == 0) {                  // last block to arrive: fixed-order sum of the per-block partials
    __threadfence();
    float tot = 0.f;
    for (unsigned i = lane; i < gridDim.x; i += 32) tot += __ldcg(partial + i);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
    if (lane == 0) {
      const float l = tot / (float)M;
      loss[0] = l;
      if (loss_acc) loss_acc[0] += l;
      *counter = 0u;
    }
  }
}

// loss = sum(partials) / M in the same fixed order as the in-kernel tail; one warp
__global__ void head_loss_finalize_kernel(const float* __restrict__ partial, int nblk, int M, float* __restrict__ loss, float* __restrict__ loss_acc) {
  const int lane = threadIdx.x;
  float tot = 0.f;
  for (int i = lane; i < nblk; i += 32) tot += partial[i];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
  if (lane == 0) {
    const float l = tot / (float)M;
    loss[0] = l;
    if (loss_acc) loss_acc[0] += l;
  }
}

}  // namespace

bool okl_dense_small_ok(int M, int K, int N) { return N >= 1 && N <= SN_MAXN && (long long)K * N <= 12000 && M >= 1; }
bool okl_dense_head_ok(int M, int K, int N) { return N >= 1 && N <= SN_MAXN && K >= 1 && K <= 256 && M >= 1; }

int okl_dense_small_fwd(okl_ctx* ctx, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act) {
  int warps = M;
  int grid = (warps * 32 + 255) / 256;
  const int cap = ctx->sm_count * 4;
  if (grid > cap) grid = cap;
  okl_launch(ctx, dense_smalln_fwd, dim3(grid), dim3(256), (size_t)K * N * sizeof(float), X, W, b, Y, M, K, N, act);
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
  okl_launch(ctx, dense_smalln_bwd, dim3(total), dim3(256), smem, X, W, G, dX, dW, db, mask, M, K, N, nbx, nbw);
  OKL_LAUNCHED(ctx, "dense_smalln_bwd");
  return OKL_OK;
}

// DenseLayer(K, N <= 16) -> SoftmaxActivationLayer -> CrossEntropyLoss, forward and backward down to dX (see the kernel).
extern "C" int okl_dense_softmax_ce(okl_ctx* ctx, const void* X, const void* W, const void* b, const void* targets_i32, void* P, void* loss,
                                    void* grad, void* dZ, void* dX, const void* dx_relu_mask, int M, int K, int N, float denom,
                                    void* loss_accum) {
  OKL_REQUIRE(ctx && X && W && targets_i32 && P && loss && grad && dZ, "NULL argument");
  OKL_REQUIRE(okl_dense_head_ok(M, K, N) && denom > 0.f, "okl_dense_softmax_ce: needs 1 <= N <= %d and 1 <= K <= 256", SN_MAXN);
  if (!ctx->head_scratch) {
    OKL_CHECK_CUDA(cudaMalloc(&ctx->head_scratch, 8192));
    OKL_CHECK_CUDA(cudaMemset(ctx->head_scratch, 0, 8192));
  }
  int grid = (M + 7) / 8;
  const int cap = ctx->sm_count * 4 < 2000 ? ctx->sm_count * 4 : 2000;
  if (grid > cap) grid = cap;
  // The loss VALUE is not needed by anything on the step's critical path: its final reduction (fence + ticket + last-block sum
  // inside the kernel, ~2 us) is done by a one-warp kernel on auxiliary lane 1 instead, joined with the other leaves of the step.
  static int defer_env = -1;
  if (defer_env < 0) { const char* e = getenv("OKL_HEAD_DEFER"); defer_env = e ? atoi(e) : 1; }
  const bool defer = defer_env && ctx->cur_lane == 0 && ctx->nranks == 1;     // data-parallel runs keep the verified in-kernel tail
  unsigned* counter = defer ? nullptr : (unsigned*)ctx->head_scratch;
  float* partial = (float*)ctx->head_scratch + 16;
  if (K <= 128)
    okl_launch(ctx, dense_softmax_ce_kernel<4>, dim3(grid), dim3(256), 0, (const float*)X, (const float*)W, (const float*)b,
               (const int*)targets_i32, (float*)P, (float*)grad, (float*)dZ, (float*)dX, (const float*)dx_relu_mask, (float*)loss,
               (float*)loss_accum, partial, counter, M, K, N, 1.f / denom);
  else
    okl_launch(ctx, dense_softmax_ce_kernel<8>, dim3(grid), dim3(256), 0, (const float*)X, (const float*)W, (const float*)b,
               (const int*)targets_i32, (float*)P, (float*)grad, (float*)dZ, (float*)dX, (const float*)dx_relu_mask, (float*)loss,
               (float*)loss_accum, partial, counter, M, K, N, 1.f / denom);
  OKL_LAUNCHED(ctx, "dense_softmax_ce");
  if (defer) {
    int s = okl_aux_mark(ctx);
    if (s == OKL_OK) s = okl_aux_begin(ctx, 1);
    if (s != OKL_OK) return s;
    head_loss_finalize_kernel<<<1, 32, 0, ctx->stream>>>(partial, grid, M, (float*)loss, (float*)loss_accum);
    s = okl_post_launch(ctx, "head_loss_finalize");
    const int e = okl_aux_end(ctx);
    return s != OKL_OK ? s : e;
  }
  return OKL_OK;
}
