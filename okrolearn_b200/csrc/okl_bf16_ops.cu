// Memory-bound companions of the BF16 tensor-core path: activations and gradients that flow between tcgen05 layers are stored as
// channels-last bf16 ONLY (the fp32 copy of the first generation doubled-and-a-half every tensor's HBM traffic and needed separate
// conversion passes).  All kernels: 16-byte accesses (8 channels per thread), grids sized to whole waves, fp32 arithmetic inside.
//   okl_from_bf16            : bf16 -> fp32 (a host mirror / an fp32 consumer asks for the tensor)
//   okl_pool2_bf16_fwd / bwd : MaxPoolingLayer(2, 2) on (N, H, W, C) bf16 (okl.py:1731-1772, R4 tie rule: every maximum of a window
//                              receives the gradient); bwd optionally applies the preceding ReLU's backward and ALSO emits the
//                              per-channel sum of the gradient it writes - the bias gradient of the convolution in front of it
//                              (okl.py:870) - as fixed-order partials, so no separate reduction pass reads the tensor again
//   okl_channel_sum_bf16     : out[c] = sum_rows g[row, c]  (bias gradients elsewhere), two fixed-order stages
#include "common.cuh"

#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void from_bf16_kernel(const uint4* __restrict__ s, float4* __restrict__ d, size_t n8, const __nv_bfloat16* __restrict__ s_tail,
                                 float* __restrict__ d_tail, int tail) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(s[i], f);
    d[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    d[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) d_tail[threadIdx.x] = __bfloat162float(s_tail[threadIdx.x]);
}

// one thread per (window, 8 channels)
__global__ void pool2_bf16_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, size_t total8, int C8, int OW, int OH, int W, int H) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t r = i / C8;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const size_t n = r / OH;
    const uint4* p = x + ((n * H + 2 * oh) * W + 2 * ow) * C8 + c;
    float a[8], b[8], d[8], e[8], m[8];
    unpack8(p[0], a); unpack8(p[C8], b); unpack8(p[(size_t)W * C8], d); unpack8(p[(size_t)W * C8 + C8], e);
#pragma unroll
    for (int j = 0; j < 8; j++) m[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(d[j], e[j]));
    y[i] = pack8(m);
  }
}

constexpr int PB_THREADS = 256;

__global__ void __launch_bounds__(PB_THREADS) pool2_bf16_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g, uint4* __restrict__ dx,
                                                                    size_t total8, int C8, int OW, int OH, int W, int H, int relu_mask,
                                                                    float* __restrict__ db_partial) {
  __shared__ float red[PB_THREADS][9];
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = 0.f;
  // gridDim.x * blockDim.x is a multiple of C8 (C8 divides 256), so a thread keeps its 8 channels over the whole grid-stride loop
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t r = i / C8;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const size_t n = r / OH;
    const size_t off = ((n * H + 2 * oh) * W + 2 * ow) * C8 + c;
    float a[8], b[8], d[8], e[8], gv[8], o0[8], o1[8], o2[8], o3[8];
    unpack8(x[off], a); unpack8(x[off + C8], b); unpack8(x[off + (size_t)W * C8], d); unpack8(x[off + (size_t)W * C8 + C8], e);
    unpack8(g[i], gv);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float m = fmaxf(fmaxf(a[j], b[j]), fmaxf(d[j], e[j]));
      float v = gv[j];
      if (relu_mask && !(m > 0.f)) v = 0.f;
      o0[j] = a[j] == m ? v : 0.f; o1[j] = b[j] == m ? v : 0.f; o2[j] = d[j] == m ? v : 0.f; o3[j] = e[j] == m ? v : 0.f;
      acc[j] += (o0[j] + o1[j]) + (o2[j] + o3[j]);
    }
    dx[off] = pack8(o0); dx[off + C8] = pack8(o1); dx[off + (size_t)W * C8] = pack8(o2); dx[off + (size_t)W * C8 + C8] = pack8(o3);
  }
  if (db_partial) {
    // fixed-order block reduction over the threads that share a channel group; partial[block][C]
#pragma unroll
    for (int j = 0; j < 8; j++) red[threadIdx.x][j] = acc[j];
    __syncthreads();
    if ((int)threadIdx.x < C8) {
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; j++) s[j] = 0.f;
      for (int t = threadIdx.x; t < PB_THREADS; t += C8)
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] += red[t][j];
      float* dst = db_partial + ((size_t)blockIdx.x * C8 + threadIdx.x) * 8;
#pragma unroll
      for (int j = 0; j < 8; j++) dst[j] = s[j];
    }
  }
}

// partial[block][C]: every block sums a contiguous chunk of rows; thread = (row lane, 8-channel group)
__global__ void __launch_bounds__(PB_THREADS) chansum_bf16_stage1(const uint4* __restrict__ g, float* __restrict__ partial, long long rows, int C8,
                                                                  long long rows_per_block) {
  __shared__ float red[PB_THREADS][9];
  const int cg = threadIdx.x % C8, rl = threadIdx.x / C8, lanes = PB_THREADS / C8;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = 0.f;
  long long r = r0 + rl;
  for (; r + 3LL * lanes < r1; r += 4LL * lanes) {           // four independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = g[(size_t)(r + (long long)u * lanes) * C8 + cg];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] += f[j];
    }
  }
  for (; r < r1; r += lanes) {
    float f[8];
    unpack8(g[(size_t)r * C8 + cg], f);
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if ((int)threadIdx.x < C8) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = 0.f;
    for (int t = threadIdx.x; t < PB_THREADS; t += C8)
#pragma unroll
      for (int j = 0; j < 8; j++) s[j] += red[t][j];
    float* dst = partial + ((size_t)blockIdx.x * C8 + threadIdx.x) * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) dst[j] = s[j];
  }
}

// out[c] = sum_blocks partial[block][c]: one warp per channel, lanes stride over the blocks, fixed-order shuffle tree (the serial
// one-thread-per-channel version took 50-90 us for ~1200 partials: pure load latency)
__global__ void chansum_stage2(const float* __restrict__ partial, float* __restrict__ out, int blocks, int C) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f;
  int k = lane;
  for (; k + 32 < blocks; k += 64) {
    s0 += partial[(size_t)k * C + c];
    s1 += partial[(size_t)(k + 32) * C + c];
  }
  if (k < blocks) s0 += partial[(size_t)k * C + c];
  float s = s0 + s1;
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
  if (lane == 0) out[c] = s;
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// ------------------------------------------------------------------------------------------------ layout-crossing kernels
// The caller's tensors are fp32 (N, C, S) (reference layout); the tensor-core convolutions read and write bf16 (N, S, C).  These two
// kernels cross both the layout and the precision in ONE pass through a shared-memory tile of 32 channels x 128 positions (512-byte
// global rows on the fp32 side, 64-byte channel groups on the bf16 side; tile[c][s] with a 129-word pitch: the transposed reads of a
// warp - 4 channel groups x 8 positions - hit 32 different banks).
constexpr int XT_C = 32, XT_S = 128, XT_PITCH = XT_S + 1;

// fills tile[c][s] with src[n][c0 + c][s0 + s] (zero outside the tensor); 256 threads
__device__ __forceinline__ void xt_load_tile(float* tile, const float* __restrict__ src, size_t img, int Cc, long long S, int c0, long long s0, bool vec) {
  if (vec) {
#pragma unroll
    for (int k = 0; k < XT_C * XT_S / 4 / 256; k++) {
      const int idx = threadIdx.x + k * 256, c = idx >> 5, q = idx & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + c < Cc && s0 + q * 4 < S) v = __ldcs(reinterpret_cast<const float4*>(src + img + (size_t)(c0 + c) * S + s0 + q * 4));
      float* d = tile + c * XT_PITCH + q * 4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int idx = threadIdx.x; idx < XT_C * XT_S; idx += 256) {
      const int c = idx >> 7, sp = idx & 127;
      tile[c * XT_PITCH + sp] = (c0 + c < Cc && s0 + sp < S) ? __ldcs(src + img + (size_t)(c0 + c) * S + s0 + sp) : 0.f;
    }
  }
}

// x (N, C, S) fp32 -> x16 (N, S, C) bf16; C % 8 == 0
// rows (optional): sample n of the output is row rows[n] of x - the batch of a shuffled epoch is read straight from the dataset
__global__ void __launch_bounds__(256) to_bf16_nxc_kernel(const float* __restrict__ x, const int* __restrict__ rows, __nv_bfloat16* __restrict__ y, int Cc,
                                                          long long S, int vec) {
  __shared__ float tile[XT_C * XT_PITCH];
  const size_t img = (size_t)blockIdx.z * Cc * S;
  const int c0 = blockIdx.x * XT_C;
  const long long s0 = (long long)blockIdx.y * XT_S;
  xt_load_tile(tile, x, rows ? (size_t)rows[blockIdx.z] * Cc * S : img, Cc, S, c0, s0, vec != 0);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < XT_S * (XT_C / 8) / 256; k++) {
    const int item = threadIdx.x + k * 256, sp = item >> 2, ch = item & 3;
    if (s0 + sp < S && c0 + ch * 8 < Cc) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] = tile[(ch * 8 + j) * XT_PITCH + sp];
      *reinterpret_cast<uint4*>(y + img + (size_t)(s0 + sp) * Cc + c0 + ch * 8) = pack8(f);
    }
  }
}

// MSELoss on the bf16 channels-last outputs of a tensor-core convolution against fp32 (N, C, S) targets (okl.py:1833-1839):
// partial sums of (o - t)^2 and the gradient 2 (o - t) / denom as bf16 (N, S, C) - with RELU the backward of the ReLU that produced
// the outputs (g * (o > 0), okl.py:107) is applied in the same pass.  Neither tensor is permuted or converted in memory.
template <bool RELU>
__global__ void __launch_bounds__(256) mse_nxc_bf16_kernel(const __nv_bfloat16* __restrict__ o, const float* __restrict__ t, const int* __restrict__ t_rows,
                                                           float* __restrict__ partial, __nv_bfloat16* __restrict__ grad, int Cc, long long S, float scale,
                                                           int vec) {
  __shared__ float tile[XT_C * XT_PITCH];
  __shared__ float sh[8];
  const size_t img = (size_t)blockIdx.z * Cc * S;
  const int c0 = blockIdx.x * XT_C;
  const long long s0 = (long long)blockIdx.y * XT_S;
  xt_load_tile(tile, t, t_rows ? (size_t)t_rows[blockIdx.z] * Cc * S : img, Cc, S, c0, s0, vec != 0);
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < XT_S * (XT_C / 8) / 256; k++) {
    const int item = threadIdx.x + k * 256, sp = item >> 2, ch = item & 3;
    if (s0 + sp < S && c0 + ch * 8 < Cc) {
      const size_t i = img + (size_t)(s0 + sp) * Cc + c0 + ch * 8;
      float ov[8], g[8];
      unpack8(__ldcs(reinterpret_cast<const uint4*>(o + i)), ov);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float d = ov[j] - tile[(ch * 8 + j) * XT_PITCH + sp];
        acc += d * d;
        g[j] = (RELU && !(ov[j] > 0.f)) ? 0.f : d * scale;
      }
      *reinterpret_cast<uint4*>(grad + i) = pack8(g);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) tot += sh[w];                // fixed order
    partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
  }
}

// loss = scale * sum(partial) in a fixed order (one block)
__global__ void mse_final_sum_kernel(const float* __restrict__ partial, int count, float* __restrict__ out, float scale, float* __restrict__ acc) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += sh[w];
    out[0] = tot * scale;
    if (acc) acc[0] += tot * scale;
  }
}

}  // namespace

extern "C" int okl_to_bf16_nxc(okl_ctx* ctx, const void* src_ncx_f32, const void* src_rows_i32, void* dst_nxc_bf16, int N, int Cc, long long S) {
  OKL_REQUIRE(ctx && src_ncx_f32 && dst_nxc_bf16, "NULL argument");
  OKL_REQUIRE(N >= 1 && Cc >= 8 && Cc % 8 == 0 && S >= 1, "okl_to_bf16_nxc: C must be a multiple of 8 (N=%d C=%d)", N, Cc);
  dim3 grid((unsigned)((Cc + XT_C - 1) / XT_C), (unsigned)((S + XT_S - 1) / XT_S), (unsigned)N);
  OKL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "okl_to_bf16_nxc: extent too large");
  const int vec = (S % 4 == 0 && (reinterpret_cast<uintptr_t>(src_ncx_f32) & 15) == 0) ? 1 : 0;
  to_bf16_nxc_kernel<<<grid, 256, 0, ctx->stream>>>((const float*)src_ncx_f32, (const int*)src_rows_i32, (__nv_bfloat16*)dst_nxc_bf16, Cc, S, vec);
  OKL_LAUNCHED(ctx, "to_bf16_nxc");
  return OKL_OK;
}

extern "C" int okl_mse_fwd_bwd_nxc_bf16(okl_ctx* ctx, const void* o16_nxc, const void* t_ncx, const void* t_rows_i32, void* loss, void* grad16_nxc, int N,
                                        int Cc, long long S, float denom, void* loss_accum, int relu_outputs) {
  OKL_REQUIRE(ctx && o16_nxc && t_ncx && loss && grad16_nxc, "NULL argument");
  OKL_REQUIRE(N >= 1 && Cc >= 8 && Cc % 8 == 0 && S >= 1 && denom > 0.f, "okl_mse_fwd_bwd_nxc_bf16: C must be a multiple of 8 (N=%d C=%d)", N, Cc);
  dim3 grid((unsigned)((Cc + XT_C - 1) / XT_C), (unsigned)((S + XT_S - 1) / XT_S), (unsigned)N);
  OKL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "okl_mse_fwd_bwd_nxc_bf16: extent too large");
  const size_t nblk = (size_t)grid.x * grid.y * grid.z;
  OKL_REQUIRE(nblk < (1u << 30), "okl_mse_fwd_bwd_nxc_bf16: too many tiles");
  int s = okl_ensure_workspace(ctx, nblk * sizeof(float));
  if (s != OKL_OK) return s;
  float* partial = (float*)ctx->workspace;
  const int vec = (S % 4 == 0 && (reinterpret_cast<uintptr_t>(t_ncx) & 15) == 0) ? 1 : 0;
  if (relu_outputs)
    mse_nxc_bf16_kernel<true><<<grid, 256, 0, ctx->stream>>>((const __nv_bfloat16*)o16_nxc, (const float*)t_ncx, (const int*)t_rows_i32, partial,
                                                             (__nv_bfloat16*)grad16_nxc, Cc, S, 2.f / denom, vec);
  else
    mse_nxc_bf16_kernel<false><<<grid, 256, 0, ctx->stream>>>((const __nv_bfloat16*)o16_nxc, (const float*)t_ncx, (const int*)t_rows_i32, partial,
                                                              (__nv_bfloat16*)grad16_nxc, Cc, S, 2.f / denom, vec);
  OKL_LAUNCHED(ctx, "mse_fwd_bwd_nxc_bf16");
  mse_final_sum_kernel<<<1, 1024, 0, ctx->stream>>>((const float*)partial, (int)nblk, (float*)loss, 1.f / ((float)N * (float)Cc * (float)S), (float*)loss_accum);
  OKL_LAUNCHED(ctx, "final_sum");
  return OKL_OK;
}

extern "C" int okl_from_bf16(okl_ctx* ctx, const void* src_bf16, void* dst_f32, size_t n) {
  OKL_REQUIRE(ctx && (n == 0 || (src_bf16 && dst_f32)), "NULL argument");
  if (n == 0) return OKL_OK;
  OKL_REQUIRE((reinterpret_cast<uintptr_t>(src_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_f32) & 15) == 0, "okl_from_bf16: buffers must be 16-byte aligned");
  const size_t n8 = n / 8;
  const int tail = (int)(n - n8 * 8);
  from_bf16_kernel<<<okl_grid_1d(ctx, n8 ? n8 : 1, 256), 256, 0, ctx->stream>>>((const uint4*)src_bf16, (float4*)dst_f32, n8,
                                                                               (const __nv_bfloat16*)src_bf16 + n8 * 8, (float*)dst_f32 + n8 * 8, tail);
  OKL_LAUNCHED(ctx, "from_bf16");
  return OKL_OK;
}

extern "C" int okl_pool2_bf16_supported(int C, int H, int W) { return (C % 8 == 0 && pow2(C / 8) && C / 8 <= 256 && H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2) ? 1 : 0; }

extern "C" int okl_pool2_bf16_fwd(okl_ctx* ctx, const void* x16, void* y16, int N, int C, int H, int W) {
  OKL_REQUIRE(ctx && x16 && y16, "NULL argument");
  OKL_REQUIRE(okl_pool2_bf16_supported(C, H, W), "okl_pool2_bf16_fwd: unsupported shape C=%d H=%d W=%d", C, H, W);
  const size_t total8 = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total8 == 0) return OKL_OK;
  pool2_bf16_fwd_kernel<<<okl_grid_1d(ctx, total8, 256, 16), 256, 0, ctx->stream>>>((const uint4*)x16, (uint4*)y16, total8, C / 8, W / 2, H / 2, W, H);
  OKL_LAUNCHED(ctx, "pool2_bf16_fwd");
  return OKL_OK;
}

// dx16 = gradient routed to the window maxima (x16 = the pooling layer's input; relu_mask: that input is a ReLU output whose backward is
// applied here).  db (optional, fp32 [C]) = per-channel sum of dx16's values before bf16 rounding.
constexpr size_t LEAF_RING = 8u << 20;

extern "C" int okl_pool2_bf16_bwd(okl_ctx* ctx, const void* x16, const void* g16, void* dx16, int N, int C, int H, int W, int relu_mask, void* db,
                                  int aux_lane) {
  OKL_REQUIRE(ctx && x16 && g16 && dx16, "NULL argument");
  OKL_REQUIRE(okl_pool2_bf16_supported(C, H, W), "okl_pool2_bf16_bwd: unsupported shape C=%d H=%d W=%d", C, H, W);
  const size_t total8 = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total8 == 0) return OKL_OK;
  int grid = okl_grid_1d(ctx, total8, PB_THREADS, 4);
  float* partial = nullptr;
  // db is a leaf of the backward chain: with an auxiliary lane its second stage leaves the critical path (~7 us per pooling layer).  The
  // partials then live in a ring of their own - the lane's second stage may still be pending when later main-lane kernels reuse the
  // workspace; a step's pooling layers take different slices, and the step's okl_aux_join precedes the next step's writes.
  const size_t pbytes = ((size_t)grid * C * sizeof(float) + 255) & ~size_t(255);
  const bool aux = db && aux_lane >= 1 && aux_lane <= 2 && ctx->cur_lane == 0 && pbytes <= LEAF_RING / 4;
  if (db && aux) {
    if (!ctx->leaf_scratch) {
      OKL_REQUIRE(!ctx->capturing, "okl_pool2_bf16_bwd: scratch ring needed during graph capture (run one eager step first)");
      OKL_CHECK_CUDA(cudaMalloc(&ctx->leaf_scratch, LEAF_RING));
    }
    if (ctx->leaf_off + pbytes > LEAF_RING) ctx->leaf_off = 0;
    partial = (float*)((char*)ctx->leaf_scratch + ctx->leaf_off);
    ctx->leaf_off += pbytes;
  } else if (db) {
    int s = okl_ensure_workspace(ctx, (size_t)grid * C * sizeof(float));
    if (s != OKL_OK) return s;
    partial = (float*)ctx->workspace;
  }
  pool2_bf16_bwd_kernel<<<grid, PB_THREADS, 0, ctx->stream>>>((const uint4*)x16, (const uint4*)g16, (uint4*)dx16, total8, C / 8, W / 2, H / 2, W, H,
                                                             relu_mask, partial);
  OKL_LAUNCHED(ctx, "pool2_bf16_bwd");
  if (db) {
    if (aux) {
      int s = okl_aux_mark(ctx);
      if (s == OKL_OK) s = okl_aux_begin(ctx, aux_lane);
      if (s != OKL_OK) return s;
    }
    chansum_stage2<<<(C * 32 + 255) / 256, 256, 0, ctx->stream>>>(partial, (float*)db, grid, C);
    const int st = okl_post_launch(ctx, "chansum_stage2");
    if (aux) {
      const int e = okl_aux_end(ctx);
      return st != OKL_OK ? st : e;
    }
    return st;
  }
  return OKL_OK;
}

extern "C" int okl_channel_sum_bf16(okl_ctx* ctx, const void* g16, void* out, long long rows, int C) {
  OKL_REQUIRE(ctx && g16 && out, "NULL argument");
  OKL_REQUIRE(C % 8 == 0 && pow2(C / 8) && C / 8 <= 256 && rows >= 0, "okl_channel_sum_bf16: C must be 8 x a power of two <= 2048 (C=%d)", C);
  long long blocks = 4LL * ctx->sm_count;
  const int lanes = PB_THREADS / (C / 8);
  long long min_rows = 4LL * lanes;
  if (blocks * min_rows > rows) blocks = (rows + min_rows - 1) / min_rows;
  if (blocks < 1) blocks = 1;
  long long rpb = (rows + blocks - 1) / blocks;
  if (rpb < 1) rpb = 1;
  blocks = rows > 0 ? (rows + rpb - 1) / rpb : 1;
  int s = okl_ensure_workspace(ctx, (size_t)blocks * C * sizeof(float));
  if (s != OKL_OK) return s;
  chansum_bf16_stage1<<<(unsigned)blocks, PB_THREADS, 0, ctx->stream>>>((const uint4*)g16, (float*)ctx->workspace, rows, C / 8, rpb);
  OKL_LAUNCHED(ctx, "chansum_bf16_stage1");
  chansum_stage2<<<(C * 32 + 255) / 256, 256, 0, ctx->stream>>>((const float*)ctx->workspace, (float*)out, (int)blocks, C);
  OKL_LAUNCHED(ctx, "chansum_stage2");
  return OKL_OK;
}
