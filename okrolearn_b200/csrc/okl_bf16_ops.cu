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

}  // namespace

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
extern "C" int okl_pool2_bf16_bwd(okl_ctx* ctx, const void* x16, const void* g16, void* dx16, int N, int C, int H, int W, int relu_mask, void* db) {
  OKL_REQUIRE(ctx && x16 && g16 && dx16, "NULL argument");
  OKL_REQUIRE(okl_pool2_bf16_supported(C, H, W), "okl_pool2_bf16_bwd: unsupported shape C=%d H=%d W=%d", C, H, W);
  const size_t total8 = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total8 == 0) return OKL_OK;
  int grid = okl_grid_1d(ctx, total8, PB_THREADS, 4);
  float* partial = nullptr;
  if (db) {
    int s = okl_ensure_workspace(ctx, (size_t)grid * C * sizeof(float));
    if (s != OKL_OK) return s;
    partial = (float*)ctx->workspace;
  }
  pool2_bf16_bwd_kernel<<<grid, PB_THREADS, 0, ctx->stream>>>((const uint4*)x16, (const uint4*)g16, (uint4*)dx16, total8, C / 8, W / 2, H / 2, W, H,
                                                             relu_mask, partial);
  OKL_LAUNCHED(ctx, "pool2_bf16_bwd");
  if (db) {
    chansum_stage2<<<(C * 32 + 255) / 256, 256, 0, ctx->stream>>>(partial, (float*)db, grid, C);
    OKL_LAUNCHED(ctx, "chansum_stage2");
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
