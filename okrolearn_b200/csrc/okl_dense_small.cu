// DenseLayer kernels for a tiny output width (N <= 16: the classifier head, e.g. 128 -> 10).  A 64x64 GEMM tile would be >= 75 %
// padding and split-K + reduce costs two launches per product, while the whole problem is a few hundred kFLOP: here the forward
// (+ bias, optionally + the SoftmaxActivationLayer that follows, okl.py:364-373) is ONE kernel and the whole backward
// (dW = X^T G, db = sum_rows G, dX = G W^T with the preceding ReLU's mask, okl.py:35-37) is ONE kernel.  Exact fp32.
#include "common.cuh"
#include <cuda_bf16.h>

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

// ================================================================================================ classifier head after a conv stack
// FlattenLayer -> DenseLayer(K = C * S up to a few thousand, N <= 16) -> SoftmaxActivationLayer -> CrossEntropyLoss (okl.py:643-651,
// 29-47, 364-393, 1389-1409) fed by a channels-last bf16 activation (N, S, C).  Round 1 ran this as twelve launches (bf16 -> fp32,
// channels-last -> NCHW permute, FFMA GEMM + split-K reduce, softmax, CE, two FFMA GEMMs + reduce, column sum, permute back,
// fp32 -> bf16; ~170 us at batch 1024, ncu).  Here: the weights are re-indexed ONCE per step into the activation's own order
// (k' = s * C + c instead of k = c * S + s) as a bf16 [N][K] bank, so nothing is ever permuted or widened:
//   head_big_kernel : one warp per sample - logits from shared-memory weights, softmax, CE value and gradient, softmax backward, and
//                     dX written straight back as channels-last bf16 (the pooling backward's input)
//   head_dw_kernel  : dW partials over row chunks (fixed order), head_dw_reduce: dW in the reference's row order + db
namespace {

constexpr int HB_THREADS = 256;

// Wt[n][k'] = bf16(W[k][n]),  k = c * S + s,  k' = s * C + c
__global__ void head_bank_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wt, int C, int S, int N) {
  const int K = C * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
    const int k = i / N, n = i - k * N;
    const int c = k / S, s = k - c * S;
    Wt[(size_t)n * K + s * C + c] = __float2bfloat16(W[i]);
  }
}

__device__ __forceinline__ void hb_unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// One block per sample at a time, the reduction split over its eight warps (a warp-per-sample version with the bank staged in shared
// memory left most of the chip idle at batch 1024: 128 blocks, 80 KB of weights staged per block, 37 us).  The bank (N * K bf16, 80 KB
// for Dense(4096, 10)) is read through L1 instead.  Every sum has a fixed order: lanes by shuffle tree, warps in index order.
__global__ void __launch_bounds__(HB_THREADS, 4) head_big_kernel(const uint4* __restrict__ x16, const uint4* __restrict__ Wt, const float* __restrict__ b,
                                                                 const int* __restrict__ t, float* __restrict__ P, float* __restrict__ grad,
                                                                 float* __restrict__ dZ, uint4* __restrict__ dx16, float* __restrict__ partial, int M,
                                                                 int K, int N, float inv_denom) {
  __shared__ float red[2][8][SN_MAXN];
  okl_pdl_sync();
  const int K8 = K >> 3;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (K8 + 7) >> 3, i0 = wib * per, i1 = min(K8, i0 + per);          // this warp's slice of the reduction
  float lsum = 0.f;
  int it = 0;
  for (int m = blockIdx.x; m < M; m += gridDim.x, it ^= 1) {
    const uint4* xr = x16 + (size_t)m * K8;
    float acc[SN_MAXN];
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) acc[n] = 0.f;
    for (int i = i0 + lane; i < i1; i += 32) {
      float xv[8];
      hb_unpack8(__ldg(xr + i), xv);
#pragma unroll
      for (int n = 0; n < SN_MAXN; n++) {
        if (n < N) {
          float wv[8];
          hb_unpack8(__ldg(Wt + (size_t)n * K8 + i), wv);
#pragma unroll
          for (int j = 0; j < 8; j++) acc[n] = fmaf(xv[j], wv[j], acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) {
      if (n < N) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], s);
        if (lane == 0) red[it][wib][n] = acc[n];
      }
    }
    __syncthreads();                                   // red[it] is rewritten two samples later: one barrier per sample is enough
    const int tc = __ldg(t + m);
    float mx = -INFINITY, sum = 0.f;
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++) {
      if (n < N) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) a += red[it][w][n];
        if (b) a += __ldg(b + n);
        acc[n] = a;
        mx = fmaxf(mx, a);
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
        v = isnan(v) ? 0.f : v;                      // nan_to_num, okl.py:370
        acc[n] = v;
        float g = fminf(fmaxf(v, 1e-12f), 1.0f);     // np.clip(p, 1e-12, 1 - 1e-12), okl.py:1391
        if (n == tc) { pt = g; g -= 1.f; }
        g *= inv_denom;
        gv[n] = g;
        dot += g * v;
      }
    }
    if (wib == 0) lsum += -logf(pt);
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++)
      if (n < N) {
        const float dz = acc[n] * (gv[n] - dot);     // softmax backward of (clip(p) - onehot) / samples, okl.py:382-391
        if (wib == 0 && lane == n) {
          P[(size_t)m * N + n] = acc[n];
          grad[(size_t)m * N + n] = gv[n];
          dZ[(size_t)m * N + n] = dz;
        }
        gv[n] = dz;
      }
    if (dx16) {
      uint4* dr = dx16 + (size_t)m * K8;
      for (int i = i0 + lane; i < i1; i += 32) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = 0.f;
#pragma unroll
        for (int n = 0; n < SN_MAXN; n++) {
          if (n < N) {
            float wv[8];
            hb_unpack8(__ldg(Wt + (size_t)n * K8 + i), wv);
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = fmaf(gv[n], wv[j], o[j]);
          }
        }
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          __nv_bfloat162 q = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
          pk[e] = *reinterpret_cast<uint32_t*>(&q);
        }
        dr[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = lsum;
}

// part[rs][k'][n] = sum over the rows of chunk rs of x[row][k'] * dz[row][n]; thread = two adjacent k' (one 4-byte load per row)
__global__ void __launch_bounds__(HB_THREADS) head_dw_kernel(const uint32_t* __restrict__ x16, const float* __restrict__ dZ, float* __restrict__ part,
                                                             int M, int K, int N, int rows_per_chunk) {
  extern __shared__ float sdz[];                   // [rows_per_chunk][N]
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  for (int i = threadIdx.x; i < (r1 - r0) * N; i += HB_THREADS) sdz[i] = dZ[(size_t)r0 * N + i];
  __syncthreads();
  const int kp = (blockIdx.x * HB_THREADS + threadIdx.x) * 2;
  if (kp >= K) return;
  float a0[SN_MAXN], a1[SN_MAXN];
#pragma unroll
  for (int n = 0; n < SN_MAXN; n++) a0[n] = a1[n] = 0.f;
  const uint32_t* xp = x16 + (kp >> 1);
  const int K2 = K >> 1;
  for (int r = r0; r < r1; r++) {
    const uint32_t v = __ldg(xp + (size_t)r * K2);
    const float x0 = __uint_as_float(v << 16), x1 = __uint_as_float(v & 0xFFFF0000u);
    const float* dz = sdz + (r - r0) * N;
#pragma unroll
    for (int n = 0; n < SN_MAXN; n++)
      if (n < N) { a0[n] = fmaf(x0, dz[n], a0[n]); a1[n] = fmaf(x1, dz[n], a1[n]); }
  }
  float* dst = part + ((size_t)blockIdx.y * K + kp) * N;
#pragma unroll
  for (int n = 0; n < SN_MAXN; n++)
    if (n < N) { dst[n] = a0[n]; dst[N + n] = a1[n]; }
}

// dW[k][n] = sum_rs part[rs][k'(k)][n] (k = c * S + s <- k' = s * C + c);  db[n] = sum_rows dZ[row][n] (one warp per class)
__global__ void head_dw_reduce(const float* __restrict__ part, const float* __restrict__ dZ, float* __restrict__ dW, float* __restrict__ db, int M,
                               int C, int S, int N, int chunks) {
  const int K = C * S;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * N) {
    const int k = i / N, n = i - k * N;
    const int c = k / S, s = k - c * S;
    const float* src = part + (size_t)(s * C + c) * N + n;
    float a = 0.f;
    for (int r = 0; r < chunks; r++) a += src[(size_t)r * K * N];
    dW[i] = a;
  }
  if (db && blockIdx.x == 0) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = w; n < N; n += blockDim.x >> 5) {
      float a = 0.f;
      for (int r = lane; r < M; r += 32) a += dZ[(size_t)r * N + n];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      if (lane == 0) db[n] = a;
    }
  }
}

}  // namespace

int head_big_leaves(okl_ctx* ctx, const void* x16, void* loss, void* dZ, void* dW, void* db, float* partial, int grid, int M, int C, int S, int N,
                    void* loss_accum);

extern "C" int okl_head_big_supported(int M, int C, int S, int N) {
  const long long K = (long long)C * S;
  return (M >= 1 && N >= 1 && N <= SN_MAXN && C % 8 == 0 && S >= 1 && K % 256 == 0 && K * N * 2 <= 160 * 1024) ? 1 : 0;
}

// x16: (M, S, C) channels-last bf16 view of the (M, C, S) activation the FlattenLayer flattens; W (C * S, N) and b (N) fp32 masters;
// outputs: P (probabilities), loss (mean CE), grad (dL/dP), dZ (dL/dlogits) fp32; dx16 (M, S, C) bf16 (optional); dW, db fp32.
// wt_scratch: N * K bf16 of device scratch for the re-indexed weight bank.
extern "C" int okl_head_big(okl_ctx* ctx, const void* x16, const void* W, const void* b, const void* targets_i32, void* P, void* loss, void* grad,
                            void* dZ, void* dx16, void* dW, void* db, void* wt_scratch, int M, int C, int S, int N, float denom, void* loss_accum,
                            int aux_lane) {
  OKL_REQUIRE(ctx && x16 && W && targets_i32 && P && loss && grad && dZ && dW && wt_scratch, "NULL argument");
  OKL_REQUIRE(okl_head_big_supported(M, C, S, N) && denom > 0.f, "okl_head_big: unsupported shape M=%d C=%d S=%d N=%d", M, C, S, N);
  const int K = C * S;
  if (!ctx->head_scratch) {
    OKL_CHECK_CUDA(cudaMalloc(&ctx->head_scratch, 8192));
    OKL_CHECK_CUDA(cudaMemset(ctx->head_scratch, 0, 8192));
  }
  float* partial = (float*)ctx->head_scratch + 16;
  head_bank_kernel<<<okl_grid_1d(ctx, (size_t)K * N, 256), 256, 0, ctx->stream>>>((const float*)W, (__nv_bfloat16*)wt_scratch, C, S, N);
  OKL_LAUNCHED(ctx, "head_bank");
  int grid = M;
  if (grid > 4 * ctx->sm_count) grid = 4 * ctx->sm_count;
  okl_launch(ctx, head_big_kernel, dim3(grid), dim3(HB_THREADS), (size_t)0, (const uint4*)x16, (const uint4*)wt_scratch, (const float*)b,
             (const int*)targets_i32, (float*)P, (float*)grad, (float*)dZ, (uint4*)dx16, partial, M, K, N, 1.f / denom);
  OKL_LAUNCHED(ctx, "head_big");
  // only dX is on the step's critical path: the loss value, dW and db are leaves and go to an auxiliary lane when the caller has one
  // (joined with the other leaves before the update / gradient exchange)
  const bool aux = aux_lane >= 1 && aux_lane <= 2 && ctx->cur_lane == 0;
  if (aux) {
    int s0 = okl_aux_mark(ctx);
    if (s0 == OKL_OK) s0 = okl_aux_begin(ctx, aux_lane);
    if (s0 != OKL_OK) return s0;
  }
  const int st = head_big_leaves(ctx, x16, loss, dZ, dW, db, partial, grid, M, C, S, N, loss_accum);
  if (aux) {
    const int e = okl_aux_end(ctx);
    return st != OKL_OK ? st : e;
  }
  return st;
}

int head_big_leaves(okl_ctx* ctx, const void* x16, void* loss, void* dZ, void* dW, void* db, float* partial, int grid, int M, int C, int S, int N,
                    void* loss_accum) {
  const int K = C * S;
  head_loss_finalize_kernel<<<1, 32, 0, ctx->stream>>>(partial, grid, M, (float*)loss, (float*)loss_accum);
  OKL_LAUNCHED(ctx, "head_loss_finalize");
  // weight gradient: row chunks so that the grid fills the chip; fixed-order second stage
  int chunks = (2 * ctx->sm_count) / ((K / 2 + HB_THREADS - 1) / HB_THREADS);
  if (chunks < 1) chunks = 1;
  if (chunks > 32) chunks = 32;
  int rpc = (M + chunks - 1) / chunks;
  if (rpc > 512) rpc = 512;                         // shared-memory stage of dZ: rows_per_chunk * N floats
  chunks = (M + rpc - 1) / rpc;
  int s = okl_ensure_workspace(ctx, (size_t)chunks * K * N * sizeof(float));
  if (s != OKL_OK) return s;
  head_dw_kernel<<<dim3((K / 2 + HB_THREADS - 1) / HB_THREADS, chunks), HB_THREADS, (size_t)rpc * N * sizeof(float), ctx->stream>>>(
      (const uint32_t*)x16, (const float*)dZ, (float*)ctx->workspace, M, K, N, rpc);
  OKL_LAUNCHED(ctx, "head_dw");
  head_dw_reduce<<<(K * N + 255) / 256, 256, 0, ctx->stream>>>((const float*)ctx->workspace, (const float*)dZ, (float*)dW, (float*)db, M, C, S, N,
                                                              chunks);
  OKL_LAUNCHED(ctx, "head_dw_reduce");
  return OKL_OK;
}
