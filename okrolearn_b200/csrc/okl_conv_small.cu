// Direct fp32 kernels for convolutions with a tiny reduction dimension K = C * taps (first layers: MNIST 1x3x3 -> K = 9,
// CIFAR 3x3x3 -> K = 27, video 3x3x3x3 -> K = 81).  With K that small a tensor-core tile would be >= 85% padding and the
// layer is bound by writing / reading the (N, O, *out) activation tensor, so these kernels are built around coalesced
// 128-byte accesses to that tensor and keep the filter bank in shared memory (SURVEY.md 7.3 "tiny-K convs").
//
//   fprop : one thread per output position, OT output channels at a time; filters in smem as [k][o] so one
//           broadcast LDS.128 feeds four FMAs; stores are coalesced along the output row for every channel.
//   wgrad : dW[o,k] = sum_pix g[pix,o] * patch[pix,k] plus db[o] = sum_pix g[pix,o] as the extra column k = K (a
//           virtual all-ones tap).  Tiles of 128 positions are staged in smem; warp w owns 4 output channels,
//           lane l owns 4 positions of the tile (float4), accumulators stay in registers across all tiles of the
//           block; lanes are combined with shuffles and blocks by a second, fixed-order kernel (deterministic).
// Both follow okl.py:790-813 / 848-883 (and the 1-D / 3-D twins) through the same im2col index math as okl_ffma.cu.
#include "common.cuh"
#include "conv_geom.cuh"

using namespace okl_conv;

namespace {

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case OKL_ACT_RELU: return v > 0.f ? v : 0.f;
    case OKL_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case OKL_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------------ fprop
constexpr int FP_OT = 16;

template <bool PAD>
__global__ void __launch_bounds__(256) conv_smallk_fprop(const float* __restrict__ x, const float* __restrict__ W,
                                                        const float* __restrict__ b, float* __restrict__ y, const ConvGeom g, const int Mpix,
                                                        const int K, const int Opad, const int act) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                 // [K][Opad]
  float* bs = sm + K * Opad;      // [Opad]
  for (int i = threadIdx.x; i < K * Opad; i += blockDim.x) {
    int k = i / Opad, o = i - k * Opad;
    Ws[i] = o < g.O ? __ldg(W + (size_t)o * K + k) : 0.f;
  }
  for (int i = threadIdx.x; i < Opad; i += blockDim.x) bs[i] = (b && i < g.O) ? __ldg(b + i) : 0.f;
  __syncthreads();

  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Mpix; m += gridDim.x * blockDim.x) {
    unsigned t, ow, oh, od;
    g.dOW.divmod((unsigned)m, t, ow); g.dOH.divmod(t, t, oh); g.dOD.divmod(t, t, od);
    const int n = (int)t;
    const int id0 = (int)od * g.S[0] - g.P[0], ih0 = (int)oh * g.S[1] - g.P[1], iw0 = (int)ow * g.S[2] - g.P[2];
    const float* xn = x + n * g.xs[0];
    float* yp = y + n * g.ys[0] + od * g.ys[2] + oh * g.ys[3] + ow * g.ys[4];   // any output layout
    for (int oc = 0; oc < Opad; oc += FP_OT) {
      float acc[FP_OT];
#pragma unroll
      for (int j = 0; j < FP_OT; j++) acc[j] = bs[oc + j];
      int k = 0;
      for (int c = 0; c < g.C; c++) {
        const float* xc = xn + c * g.xs[1];
        for (int td = 0; td < g.Kk[0]; td++) {
          const int id = id0 + td;
          const bool okd = !PAD || (unsigned)id < (unsigned)g.I[0];
          for (int th = 0; th < g.Kk[1]; th++) {
            const int ih = ih0 + th;
            const bool okh = !PAD || (okd && (unsigned)ih < (unsigned)g.I[1]);
            const float* xr = xc + id * g.xs[2] + ih * g.xs[3];
            for (int tw = 0; tw < g.Kk[2]; tw++, k++) {
              const int iw = iw0 + tw;
              const float xv = (!PAD || (okh && (unsigned)iw < (unsigned)g.I[2])) ? __ldg(xr + iw * g.xs[4]) : 0.f;
              const float4* wr = reinterpret_cast<const float4*>(Ws + k * Opad + oc);
#pragma unroll
              for (int q = 0; q < FP_OT / 4; q++) {
                const float4 w4 = wr[q];
                acc[4 * q + 0] = fmaf(xv, w4.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(xv, w4.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(xv, w4.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(xv, w4.w, acc[4 * q + 3]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < FP_OT; j++)
        if (oc + j < g.O) yp[(long long)(oc + j) * g.ys[1]] = act_apply(acc[j], act);
    }
  }
}

// Fast path: unit stride along the innermost dimension, KH x KW window known at compile time.  Each thread produces TWO
// adjacent outputs of a row for OT = 32 channels at a time: one (KW+1)-wide input window per filter row feeds both outputs,
// one broadcast LDS.128 of four filter taps feeds eight FMAs, and every channel's pair is written with one 64-bit store.
template <int KH, int KW, bool PAD>
__global__ void __launch_bounds__(256, 2) conv_smallk_fprop2(const float* __restrict__ x, const float* __restrict__ W,
                                                            const float* __restrict__ b, float* __restrict__ y, const ConvGeom g,
                                                            const int K, const int Opad, const int act, const FastDiv dPairs,
                                                            const int total) {
  constexpr int OT = 32;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                 // [K][Opad]
  float* bs = sm + K * Opad;      // [Opad]
  for (int i = threadIdx.x; i < K * Opad; i += blockDim.x) {
    int k = i / Opad, o = i - k * Opad;
    Ws[i] = o < g.O ? __ldg(W + (size_t)o * K + k) : 0.f;
  }
  for (int i = threadIdx.x; i < Opad; i += blockDim.x) bs[i] = (b && i < g.O) ? __ldg(b + i) : 0.f;
  __syncthreads();
  const bool nxc_out = g.ys[1] == 1;                       // channels-last output: 32 channels of a pixel are contiguous
  const bool pair_store = !nxc_out && (g.Oe[2] & 1) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0;
  const bool vec_out = nxc_out && (g.O & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;

  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
    unsigned t, wp, oh, od;
    dPairs.divmod((unsigned)it, t, wp); g.dOH.divmod(t, t, oh); g.dOD.divmod(t, t, od);
    const int n = (int)t;
    const int ow0 = 2 * (int)wp;
    const bool has2 = ow0 + 1 < g.Oe[2];
    const int id0 = (int)od * g.S[0] - g.P[0], ih0 = (int)oh * g.S[1] - g.P[1], iw0 = ow0 - g.P[2];
    const float* xn = x + n * g.xs[0];
    float* yp = y + n * g.ys[0] + od * g.ys[2] + oh * g.ys[3] + ow0 * g.ys[4];
    for (int oc = 0; oc < Opad; oc += OT) {
      float a0[OT], a1[OT];
#pragma unroll
      for (int j = 0; j < OT; j++) { a0[j] = bs[oc + j]; a1[j] = a0[j]; }
      int kbase = 0;
      for (int c = 0; c < g.C; c++) {
        for (int td = 0; td < g.Kk[0]; td++, kbase += KH * KW) {
          const int id = id0 + td;
          if (PAD && (unsigned)id >= (unsigned)g.I[0]) continue;
          const float* xp = xn + c * g.xs[1] + id * g.xs[2];
#pragma unroll
          for (int th = 0; th < KH; th++) {
            const int ih = ih0 + th;
            const bool okh = !PAD || (unsigned)ih < (unsigned)g.I[1];
            const float* xr = xp + ih * g.xs[3] + iw0;
            float xw[KW + 1];
#pragma unroll
            for (int u = 0; u < KW + 1; u++) {
              const bool ok = okh && (!PAD || (unsigned)(iw0 + u) < (unsigned)g.I[2]) && (u < KW || has2);
              xw[u] = ok ? __ldg(xr + u) : 0.f;
            }
#pragma unroll
            for (int tw = 0; tw < KW; tw++) {
              const float4* wr = reinterpret_cast<const float4*>(Ws + (kbase + th * KW + tw) * Opad + oc);
              const float x0 = xw[tw], x1 = xw[tw + 1];
#pragma unroll
              for (int q = 0; q < OT / 4; q++) {
                const float4 w4 = wr[q];
                a0[4 * q + 0] = fmaf(x0, w4.x, a0[4 * q + 0]); a1[4 * q + 0] = fmaf(x1, w4.x, a1[4 * q + 0]);
                a0[4 * q + 1] = fmaf(x0, w4.y, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(x1, w4.y, a1[4 * q + 1]);
                a0[4 * q + 2] = fmaf(x0, w4.z, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(x1, w4.z, a1[4 * q + 2]);
                a0[4 * q + 3] = fmaf(x0, w4.w, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(x1, w4.w, a1[4 * q + 3]);
              }
            }
          }
        }
      }
      if (vec_out) {
#pragma unroll
        for (int j = 0; j < OT; j += 4) {
          if (oc + j < g.O) {
            *reinterpret_cast<float4*>(yp + oc + j) =
                make_float4(act_apply(a0[j], act), act_apply(a0[j + 1], act), act_apply(a0[j + 2], act), act_apply(a0[j + 3], act));
            if (has2)
              *reinterpret_cast<float4*>(yp + g.ys[4] + oc + j) =
                  make_float4(act_apply(a1[j], act), act_apply(a1[j + 1], act), act_apply(a1[j + 2], act), act_apply(a1[j + 3], act));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < OT; j++) {
          if (oc + j < g.O) {
            float* dst = yp + (long long)(oc + j) * g.ys[1];
            const float v0 = act_apply(a0[j], act), v1 = act_apply(a1[j], act);
            if (pair_store) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
            else { dst[0] = v0; if (has2) dst[g.ys[4]] = v1; }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ wgrad (+ bias grad)
constexpr int WG_PT = 128;      // positions per tile
constexpr int WG_OB = 32;       // output channels per block (8 warps x 4)

template <int KT, bool PAD>
__global__ void __launch_bounds__(256, KT <= 12 ? 2 : 1)
conv_smallk_wgrad(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partial, const ConvGeom g, const int Mpix,
                  const int K, const int ntiles) {
  __shared__ __align__(16) float gs[WG_OB][WG_PT];
  __shared__ __align__(16) float xs[KT][WG_PT];
  __shared__ int koff[KT];          // offset of tap k from a position's window origin
  __shared__ int ktap[KT];          // packed (td, th, tw) of tap k for the padding bounds checks
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o_base = blockIdx.y * WG_OB;
  const int p = tid & (WG_PT - 1), half = tid >> 7;          // this thread stages position p, rows half, half+2, ...
  static_assert(WG_OB % 2 == 0, "two staging threads per position");
  if (tid < KT) {
    int off = 0, pk = 0;
    if (tid < K) {
      unsigned u, tw, th, td;
      g.dKW.divmod((unsigned)tid, u, tw); g.dKH.divmod(u, u, th); g.dKD.divmod(u, u, td);
      off = (int)(u * g.xs[1] + td * g.xs[2] + th * g.xs[3] + tw * g.xs[4]);
      pk = (int)((td << 20) | (th << 10) | tw);
    }
    koff[tid] = off;
    ktap[tid] = pk;
  }
  __syncthreads();

  float acc[4][KT];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < KT; k++) acc[i][k] = 0.f;

  constexpr int NG = WG_OB / 2, NX = (KT + 1) / 2;            // values this thread stages per tile
  float gnext[NG], xnext[NX];
  auto fetch = [&](int tile) {
    const int m = tile * WG_PT + p;
    const bool ok = m < Mpix;
    unsigned t, ow, oh, od;
    g.dOW.divmod((unsigned)m, t, ow); g.dOH.divmod(t, t, oh); g.dOD.divmod(t, t, od);
    const float* gp = gy + t * g.ys[0] + od * g.ys[2] + oh * g.ys[3] + ow * g.ys[4];
#pragma unroll
    for (int i = 0; i < NG; i++) {
      const int o = half + 2 * i;
      gnext[i] = (ok && o_base + o < g.O) ? __ldg(gp + (long long)(o_base + o) * g.ys[1]) : 0.f;
    }
    const int id0 = (int)od * g.S[0] - g.P[0], ih0 = (int)oh * g.S[1] - g.P[1], iw0 = (int)ow * g.S[2] - g.P[2];
    const float* xb = x + t * g.xs[0] + id0 * g.xs[2] + ih0 * g.xs[3] + iw0 * g.xs[4];
#pragma unroll
    for (int i = 0; i < NX; i++) {
      const int k = half + 2 * i;
      float v = 0.f;
      if (k < K && ok) {
        bool in = true;
        if (PAD) {
          const int pk = ktap[k];
          in = (unsigned)(id0 + (pk >> 20)) < (unsigned)g.I[0] && (unsigned)(ih0 + ((pk >> 10) & 1023)) < (unsigned)g.I[1] &&
               (unsigned)(iw0 + (pk & 1023)) < (unsigned)g.I[2];
        }
        if (in) v = __ldg(xb + koff[k]);
      } else if (k == K && ok) {
        v = 1.f;                                              // the all-ones tap: column K accumulates the bias gradient
      }
      xnext[i] = v;
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) fetch(tile);
  for (; tile < ntiles; tile += gridDim.x) {
    __syncthreads();                                          // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < NG; i++) gs[half + 2 * i][p] = gnext[i];
#pragma unroll
    for (int i = 0; i < NX; i++)
      if (half + 2 * i < KT) xs[half + 2 * i][p] = xnext[i];
    __syncthreads();
    if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);   // global loads of the next tile fly during the FMAs below
    float4 gv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) gv[i] = *reinterpret_cast<const float4*>(&gs[warp * 4 + i][lane * 4]);
#pragma unroll
    for (int k = 0; k < KT; k++) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[k][lane * 4]);
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i][k] += gv[i].x * xv.x + gv[i].y * xv.y + gv[i].z * xv.z + gv[i].w * xv.w;
    }
  }
  // combine the 32 lanes (fixed butterfly order), lane 0 writes partial[block][o][k]
  float* out = partial + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * WG_OB + warp * 4) * KT;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < KT; k++) {
      float v = acc[i][k];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) out[i * KT + k] = v;
    }
}

// dW[o, k] = sum_b partial[y][b][o_local][k] (k < K) ; db[o] = sum_b partial[...][K].  One warp per output element: the
// lanes stride over the blocks' partials (independent loads), then a fixed-order butterfly.
__global__ void conv_smallk_wgrad_reduce(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ db, int O, int K,
                                         int KT, int gx) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (idx >= O * (K + 1)) return;
  const int o = idx / (K + 1), k = idx - o * (K + 1);
  const int oy = o / WG_OB, ol = o - oy * WG_OB;
  const float* p = partial + ((size_t)oy * gx * WG_OB + ol) * KT + k;
  float s = 0.f;
  for (int b = lane; b < gx; b += 32) s += p[(size_t)b * WG_OB * KT];
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
  if (lane == 0) {
    if (k < K) { if (dW) dW[(size_t)o * K + k] = s; }
    else if (db) db[o] = s;
  }
}

}  // namespace

bool okl_conv_small_fprop_ok(const okl_conv_desc* d) {
  if (d->x_layout != OKL_NCX && d->C != 1) return false;
  long long K = d->C;
  for (int i = 0; i < d->nd; i++) K *= d->k[i];
  int Opad = (d->O + FP_OT - 1) / FP_OT * FP_OT;
  return K <= 96 && K * Opad + Opad <= 11000;
}

bool okl_conv_small_wgrad_ok(const okl_conv_desc* d) {
  if (d->x_layout != OKL_NCX && d->C != 1) return false;
  long long K = d->C;
  for (int i = 0; i < d->nd; i++) K *= d->k[i];
  return K + 1 <= 28;
}

template <int KH, int KW>
static int launch_fprop2(okl_ctx* ctx, const ConvGeom& g, const float* x, const float* W, const float* b, float* y, int K, int act) {
  const int Opad = (g.O + 31) / 32 * 32;
  const size_t smem = (size_t)(K * Opad + Opad) * sizeof(float);
  const int npairs = (g.Oe[2] + 1) / 2;
  const long long total_ll = (long long)g.N * g.Oe[0] * g.Oe[1] * npairs;
  const int total = (int)total_ll;
  int grid = (total + 255) / 256;
  const int cap = ctx->sm_count * 16;
  if (grid > cap) grid = cap;
  const bool pad = g.P[0] || g.P[1] || g.P[2];
  if (pad) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(conv_smallk_fprop2<KH, KW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_smallk_fprop2<KH, KW, true><<<grid, 256, smem, ctx->stream>>>(x, W, b, y, g, K, Opad, act, FastDiv((unsigned)npairs), total);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(conv_smallk_fprop2<KH, KW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_smallk_fprop2<KH, KW, false><<<grid, 256, smem, ctx->stream>>>(x, W, b, y, g, K, Opad, act, FastDiv((unsigned)npairs), total);
  }
  OKL_LAUNCHED(ctx, "conv_smallk_fprop2");
  return OKL_OK;
}

int okl_conv_small_fprop(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* W, const float* b, float* y) {
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  const int Mpix = g.N * g.Oe[0] * g.Oe[1] * g.Oe[2], K = g.C * g.T;
  if (Mpix == 0) return OKL_OK;
  if (g.S[2] == 1 && (long long)K * ((g.O + 31) / 32 * 32) <= 11000) {     // two-outputs-per-thread fast path
    if (g.Kk[1] == 3 && g.Kk[2] == 3) return launch_fprop2<3, 3>(ctx, g, x, W, b, y, K, d->act);
    if (g.Kk[1] == 1 && g.Kk[2] == 3) return launch_fprop2<1, 3>(ctx, g, x, W, b, y, K, d->act);
    if (g.Kk[1] == 5 && g.Kk[2] == 5) return launch_fprop2<5, 5>(ctx, g, x, W, b, y, K, d->act);
  }
  const int Opad = (g.O + FP_OT - 1) / FP_OT * FP_OT;
  const size_t smem = (size_t)(K * Opad + Opad) * sizeof(float);
  int grid = (Mpix + 255) / 256;
  const int cap = ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  const bool pad = g.P[0] || g.P[1] || g.P[2];
  if (pad) conv_smallk_fprop<true><<<grid, 256, smem, ctx->stream>>>(x, W, b, y, g, Mpix, K, Opad, d->act);
  else conv_smallk_fprop<false><<<grid, 256, smem, ctx->stream>>>(x, W, b, y, g, Mpix, K, Opad, d->act);
  OKL_LAUNCHED(ctx, "conv_smallk_fprop");
  return OKL_OK;
}

int okl_conv_small_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* gy, float* dW, float* db) {
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  const int Mpix = g.N * g.Oe[0] * g.Oe[1] * g.Oe[2], K = g.C * g.T;
  const int KT = K + 1 <= 12 ? 12 : 28;
  const int ntiles = (Mpix + WG_PT - 1) / WG_PT;
  const int gy_ = (g.O + WG_OB - 1) / WG_OB;
  int gx = ((KT <= 12 ? 4 : 2) * ctx->sm_count + gy_ - 1) / gy_;
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  s = okl_ensure_workspace(ctx, (size_t)gx * gy_ * WG_OB * KT * sizeof(float));
  if (s != OKL_OK) return s;
  float* partial = (float*)ctx->workspace;
  dim3 grid(gx, gy_);
  const bool pad = g.P[0] || g.P[1] || g.P[2];
  OKL_REQUIRE(g.Kk[0] < 1024 && g.Kk[1] < 1024 && g.Kk[2] < 1024, "filter extent too large");
  if (KT == 12) {
    if (pad) conv_smallk_wgrad<12, true><<<grid, 256, 0, ctx->stream>>>(x, gy, partial, g, Mpix, K, ntiles);
    else conv_smallk_wgrad<12, false><<<grid, 256, 0, ctx->stream>>>(x, gy, partial, g, Mpix, K, ntiles);
  } else {
    if (pad) conv_smallk_wgrad<28, true><<<grid, 256, 0, ctx->stream>>>(x, gy, partial, g, Mpix, K, ntiles);
    else conv_smallk_wgrad<28, false><<<grid, 256, 0, ctx->stream>>>(x, gy, partial, g, Mpix, K, ntiles);
  }
  OKL_LAUNCHED(ctx, "conv_smallk_wgrad");
  const int total = g.O * (K + 1);
  conv_smallk_wgrad_reduce<<<(total * 32 + 255) / 256, 256, 0, ctx->stream>>>(partial, dW, db, g.O, K, KT, gx);
  OKL_LAUNCHED(ctx, "conv_smallk_wgrad_reduce");
  return OKL_OK;
}
