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
  okl_pdl_sync();
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
template <int KH, int KW, bool PAD, int THREADS = 256, int MINB = 2>
__global__ void __launch_bounds__(THREADS, MINB) conv_smallk_fprop2(const float* __restrict__ x, const float* __restrict__ W,
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
  const int k0 = blockIdx.z * KT;                             // this block's slice of the K + 1 taps (large K: several slices)
  const int p = tid & (WG_PT - 1), half = tid >> 7;          // this thread stages position p, rows half, half+2, ...
  static_assert(WG_OB % 2 == 0, "two staging threads per position");
  if (tid < KT) {
    int off = 0, pk = 0;
    if (k0 + tid < K) {
      unsigned u, tw, th, td;
      g.dKW.divmod((unsigned)(k0 + tid), u, tw); g.dKH.divmod(u, u, th); g.dKD.divmod(u, u, td);
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
      if (k0 + k < K && ok) {
        bool in = true;
        if (PAD) {
          const int pk = ktap[k];
          in = (unsigned)(id0 + (pk >> 20)) < (unsigned)g.I[0] && (unsigned)(ih0 + ((pk >> 10) & 1023)) < (unsigned)g.I[1] &&
               (unsigned)(iw0 + (pk & 1023)) < (unsigned)g.I[2];
        }
        if (in) v = __ldg(xb + koff[k]);
      } else if (k0 + k == K && ok) {
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
      for (int i = 0; i < 4; i++)
        acc[i][k] = fmaf(gv[i].w, xv.w, fmaf(gv[i].z, xv.z, fmaf(gv[i].y, xv.y, fmaf(gv[i].x, xv.x, acc[i][k]))));
    }
  }
  // combine the 32 lanes (fixed butterfly order), lane 0 writes partial[block][o][k]
  float* out = partial + ((size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * WG_OB + warp * 4) * KT;
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
                                         int KT, int gx, int OB, int gy) {
  okl_pdl_sync();
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (idx >= O * (K + 1)) return;
  const int o = idx / (K + 1), k = idx - o * (K + 1);
  const int oy = o / OB, ol = o - oy * OB;
  const int kz = k / KT, kl = k - kz * KT;                    // tap slice (blockIdx.z of the producer) and column inside it
  const float* p = partial + ((size_t)(kz * gy + oy) * gx * OB + ol) * KT + kl;
  float s = 0.f;
  for (int b = lane; b < gx; b += 32) s += p[(size_t)b * OB * KT];
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
  return K <= 96 && K + 1 <= 28 * 4;                                     // up to four 28-tap slices (video 3x3x3x3: K + 1 = 82 -> 3 slices)
}

static bool env_flag_fprop3() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("OKL_FPROP2_SMALLBLOCK"); v = e ? atoi(e) : 1; }
  return v != 0;
}

template <int KH, int KW>
static int launch_fprop2(okl_ctx* ctx, const ConvGeom& g, const float* x, const float* W, const float* b, float* y, int K, int act) {
  const int Opad = (g.O + 31) / 32 * 32;
  const size_t smem = (size_t)(K * Opad + Opad) * sizeof(float);
  const int npairs = (g.Oe[2] + 1) / 2;
  const long long total_ll = (long long)g.N * g.Oe[0] * g.Oe[1] * npairs;
  const int total = (int)total_ll;
  const bool pad = g.P[0] || g.P[1] || g.P[2];
  if (env_flag_fprop3()) {
    // the fully unrolled KH x KW body with 64 accumulators needs ~160 registers; at the 128 the two-blocks-of-256
    // variant allows, the accumulators spill inside the FMA loop (1 local load per 3 FMAs, measured 15x off the FMA peak).
    // Three blocks of 128 threads with up to 170 registers keep everything in registers.
    int grid3 = (total + 127) / 128;
    if (grid3 > ctx->sm_count * 24) grid3 = ctx->sm_count * 24;
    if (pad) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(conv_smallk_fprop2<KH, KW, true, 128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      conv_smallk_fprop2<KH, KW, true, 128, 3><<<grid3, 128, smem, ctx->stream>>>(x, W, b, y, g, K, Opad, act, FastDiv((unsigned)npairs), total);
    } else {
      if (smem > 48 * 1024) cudaFuncSetAttribute(conv_smallk_fprop2<KH, KW, false, 128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      conv_smallk_fprop2<KH, KW, false, 128, 3><<<grid3, 128, smem, ctx->stream>>>(x, W, b, y, g, K, Opad, act, FastDiv((unsigned)npairs), total);
    }
    OKL_LAUNCHED(ctx, "conv_smallk_fprop2_3d");
    return OKL_OK;
  }
  int grid = (total + 255) / 256;
  const int cap = ctx->sm_count * 16;
  if (grid > cap) grid = cap;
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
  const int gz = (K + 1 + KT - 1) / KT;                       // tap slices
  const int ntiles = (Mpix + WG_PT - 1) / WG_PT;
  const int gy_ = (g.O + WG_OB - 1) / WG_OB;
  int gx = ((KT <= 12 ? 4 : 2) * ctx->sm_count) / (gy_ * gz);  // floor: whole waves of resident blocks (2 / 1 per SM), no straggler wave
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  s = okl_ensure_workspace(ctx, (size_t)gx * gy_ * gz * WG_OB * KT * sizeof(float));
  if (s != OKL_OK) return s;
  float* partial = (float*)ctx->workspace;
  dim3 grid(gx, gy_, gz);
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
  okl_launch(ctx, conv_smallk_wgrad_reduce, dim3((total * 32 + 255) / 256), dim3(256), 0, (const float*)partial, dW, db, g.O, K, KT, gx, WG_OB, gy_);
  OKL_LAUNCHED(ctx, "conv_smallk_wgrad_reduce");
  return OKL_OK;
}

// ================================================================================================ Conv -> ReLU -> MaxPool(2,2) block
// The first block of an MNIST/CIFAR-style CNN is bound by the traffic of the un-pooled activation tensor (22 MB per step at
// batch 256 for the MNIST CNN: written by the conv, read by the pool, read + written by the pool backward, read by wgrad).
// Fused, that tensor never exists:
//   forward : each thread computes the 2x2 group of conv outputs that feed ONE pooled position (bias + ReLU applied), writes the
//             pooled maximum and a 4-bit mask of which of the four positions attain it (and are > 0).  The mask IS the combined
//             backward of ReLU and max-pool, with the reference's tie rule (every maximum receives the gradient, okl.py:1757-1766).
//   backward: dW / db accumulate only over the positions the mask selects (a quarter of the conv outputs), reading the pooled
//             gradient, the mask and the (4 x 4) input windows - never the un-pooled gradient.
namespace {

// One thread per (pooled position, group of OT = 8 output channels): enough threads to fill the chip at MNIST sizes, consecutive
// threads = consecutive pooled positions of one channel group, so the pooled / mask stores are coalesced per channel.
template <int KH, int KW, bool PAD>
__global__ void __launch_bounds__(256, 3) conv_relu_pool2_fwd(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                                             float* __restrict__ yp, unsigned char* __restrict__ mask, const ConvGeom g,
                                                             const int K, const int Opad, const int PH, const int PW, const FastDiv dPW,
                                                             const FastDiv dPH, const FastDiv dPix, const int npix, const int total) {
  constexpr int OT = 8, WH = KH + 1, WW = KW + 1;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                 // [K][Opad]
  float* bs = sm + K * Opad;      // [Opad]
  for (int i = threadIdx.x; i < K * Opad; i += blockDim.x) {
    int k = i / Opad, o = i - k * Opad;
    Ws[i] = o < g.O ? __ldg(W + (size_t)o * K + k) : 0.f;
  }
  for (int i = threadIdx.x; i < Opad; i += blockDim.x) bs[i] = (b && i < g.O) ? __ldg(b + i) : 0.f;
  __syncthreads();
  const long long plane = (long long)PH * PW;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
    unsigned chunk, pix, t, pw, ph;
    dPix.divmod((unsigned)it, chunk, pix);
    dPW.divmod(pix, t, pw); dPH.divmod(t, t, ph);
    const int n = (int)t, oc = (int)chunk * OT;
    const int ih0 = 2 * (int)ph - g.P[1], iw0 = 2 * (int)pw - g.P[2];
    const float* xn = x + n * g.xs[0];
    float a[4][OT];
#pragma unroll
    for (int j = 0; j < OT; j++) { const float bv = bs[oc + j]; a[0][j] = bv; a[1][j] = bv; a[2][j] = bv; a[3][j] = bv; }
    for (int c = 0; c < g.C; c++) {
      const float* xc = xn + c * g.xs[1];
      float win[WH][WW];
#pragma unroll
      for (int r = 0; r < WH; r++)
#pragma unroll
        for (int q = 0; q < WW; q++) {
          const int ih = ih0 + r, iw = iw0 + q;
          const bool ok = !PAD || ((unsigned)ih < (unsigned)g.I[1] && (unsigned)iw < (unsigned)g.I[2]);
          win[r][q] = ok ? __ldg(xc + ih * g.xs[3] + iw) : 0.f;
        }
#pragma unroll
      for (int th = 0; th < KH; th++)
#pragma unroll
        for (int tw = 0; tw < KW; tw++) {
          const float4* wr = reinterpret_cast<const float4*>(Ws + ((c * KH + th) * KW + tw) * Opad + oc);
          const float x00 = win[th][tw], x01 = win[th][tw + 1], x10 = win[th + 1][tw], x11 = win[th + 1][tw + 1];
#pragma unroll
          for (int q = 0; q < OT / 4; q++) {
            const float4 w4 = wr[q];
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              a[0][4 * q + e] = fmaf(x00, wv[e], a[0][4 * q + e]);
              a[1][4 * q + e] = fmaf(x01, wv[e], a[1][4 * q + e]);
              a[2][4 * q + e] = fmaf(x10, wv[e], a[2][4 * q + e]);
              a[3][4 * q + e] = fmaf(x11, wv[e], a[3][4 * q + e]);
            }
          }
        }
    }
    const long long obase = (long long)n * g.O * plane + ph * PW + pw;
#pragma unroll
    for (int j = 0; j < OT; j++) {
      if (oc + j < g.O) {
        const float v0 = fmaxf(a[0][j], 0.f), v1 = fmaxf(a[1][j], 0.f), v2 = fmaxf(a[2][j], 0.f), v3 = fmaxf(a[3][j], 0.f);
        const float m = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
        unsigned mk = 0;
        if (m > 0.f) mk = (v0 == m ? 1u : 0u) | (v1 == m ? 2u : 0u) | (v2 == m ? 4u : 0u) | (v3 == m ? 8u : 0u);
        const long long off = obase + (long long)(oc + j) * plane;
        yp[off] = m;
        mask[off] = (unsigned char)mk;
      }
    }
  }
}

// dW / db from the POOLED gradient.  One 128-thread block per tile of 128 pooled positions and 16 output channels: thread p first
// issues every global load of its position (16 gradients, 16 masks, the C*(KH+1)*(KW+1) input window) back to back, then warp w
// accumulates channels 4w..4w+3 with lane l owning positions 4l..4l+3.
constexpr int CP_OB = 16;
template <int KH, int KW, int CMAX, int KT, bool PAD>
__global__ void __launch_bounds__(128, CMAX == 1 ? 5 : 3)
conv_relu_pool2_wgrad(const float* __restrict__ x, const float* __restrict__ gp, const unsigned char* __restrict__ mask,
                      float* __restrict__ partial, const ConvGeom g, const int K, const int PH, const int PW, const FastDiv dPW,
                      const FastDiv dPH, const int Mpool, const int ntiles) {
  constexpr int WH = KH + 1, WW = KW + 1, NW = CMAX * WH * WW;
  __shared__ __align__(16) float gs[CP_OB][WG_PT];
  __shared__ __align__(16) float xw[NW][WG_PT];
  __shared__ unsigned char ms[CP_OB][WG_PT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o_base = blockIdx.y * CP_OB;
  const int p = tid;
  okl_pdl_sync();
  const long long plane = (long long)PH * PW;
  float acc[4][KT];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < KT; k++) acc[i][k] = 0.f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m = tile * WG_PT + p;
    const bool ok = m < Mpool;
    unsigned t, pw, ph;
    dPW.divmod((unsigned)m, t, pw); dPH.divmod(t, t, ph);
    float gr[CP_OB], wr[NW];
    unsigned char mr[CP_OB];
    {
      const long long gb = (long long)t * g.O * plane + ph * PW + pw;
#pragma unroll
      for (int o = 0; o < CP_OB; o++) {
        const bool v = ok && o_base + o < g.O;
        gr[o] = v ? __ldg(gp + gb + (long long)(o_base + o) * plane) : 0.f;
        mr[o] = v ? __ldg(mask + gb + (long long)(o_base + o) * plane) : (unsigned char)0;
      }
      const int ih0 = 2 * (int)ph - g.P[1], iw0 = 2 * (int)pw - g.P[2];
      const float* xn = x + t * g.xs[0];
#pragma unroll
      for (int r = 0; r < NW; r++) {
        const int c = r / (WH * WW), rr = r % (WH * WW);
        const int ih = ih0 + rr / WW, iw = iw0 + rr % WW;
        const bool in = ok && (!PAD || ((unsigned)ih < (unsigned)g.I[1] && (unsigned)iw < (unsigned)g.I[2]));
        wr[r] = in ? __ldg(xn + c * g.xs[1] + ih * g.xs[3] + iw) : 0.f;
      }
    }
    __syncthreads();                                   // previous tile fully consumed
#pragma unroll
    for (int o = 0; o < CP_OB; o++) { gs[o][p] = gr[o]; ms[o][p] = mr[o]; }
#pragma unroll
    for (int r = 0; r < NW; r++) xw[r][p] = wr[r];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int o = warp * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int pp = lane + 32 * j;                     // consecutive lanes -> consecutive words: conflict-free (lane * 4 + j was 4-way)
        const float gv = gs[o][pp];
        unsigned mk = ms[o][pp];
        while (mk) {                                   // one iteration per maximum of the window (almost always exactly one)
          const int bit = __ffs(mk) - 1;
          mk &= mk - 1;
          const float* wbase = &xw[(bit >> 1) * WW + (bit & 1)][pp];
#pragma unroll
          for (int c = 0; c < CMAX; c++) {
#pragma unroll
            for (int th = 0; th < KH; th++)
#pragma unroll
              for (int tw = 0; tw < KW; tw++) {
                const int k = (c * KH + th) * KW + tw;
                if (k < KT) acc[i][k] = fmaf(gv, wbase[(c * WH * WW + th * WW + tw) * WG_PT], acc[i][k]);
              }
          }
          acc[i][CMAX * KH * KW] += gv;                // bias gradient: the all-ones tap lives in column K
        }
      }
    }
  }
  float* out = partial + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * CP_OB + warp * 4) * KT;
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

}  // namespace

extern "C" int okl_conv_relu_pool_supported(const okl_conv_desc* d) {
  if (!d || d->nd != 2 || d->x_layout != OKL_NCX || d->y_layout != OKL_NCX) return 0;
  if (d->k[0] != 3 || d->k[1] != 3 || d->s[0] != 1 || d->s[1] != 1) return 0;
  if (d->C != 1 && d->C != 3) return 0;
  const int oh = d->in[0] + 2 * d->p[0] - 2, ow = d->in[1] + 2 * d->p[1] - 2;
  if (oh < 2 || ow < 2) return 0;
  const int Opad = (d->O + 15) / 16 * 16;
  return (long long)d->C * 9 * Opad + Opad <= 11000 ? 1 : 0;
}

extern "C" int okl_conv_relu_pool_fwd(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* W, const void* b, void* pooled,
                                      void* mask) {
  OKL_REQUIRE(ctx && d && x && W && pooled && mask, "NULL argument");
  OKL_REQUIRE(okl_conv_relu_pool_supported(d), "okl_conv_relu_pool_fwd: unsupported shape");
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  const int K = g.C * 9, Opad = (g.O + 15) / 16 * 16;
  const int PH = g.Oe[1] / 2, PW = g.Oe[2] / 2;
  const int npix = g.N * PH * PW;
  const int total = npix * (Opad / 8);
  if (total == 0) return OKL_OK;
  const size_t smem = (size_t)(K * Opad + Opad) * sizeof(float);
  int grid = (total + 255) / 256;
  const int cap = ctx->sm_count * 16;
  if (grid > cap) grid = cap;
  const bool pad = g.P[1] || g.P[2];
  if (pad) okl_launch(ctx, conv_relu_pool2_fwd<3, 3, true>, dim3(grid), dim3(256), smem, (const float*)x, (const float*)W, (const float*)b,
                      (float*)pooled, (unsigned char*)mask, g, K, Opad, PH, PW, FastDiv(PW), FastDiv(PH), FastDiv(npix), npix, total);
  else okl_launch(ctx, conv_relu_pool2_fwd<3, 3, false>, dim3(grid), dim3(256), smem, (const float*)x, (const float*)W, (const float*)b,
                  (float*)pooled, (unsigned char*)mask, g, K, Opad, PH, PW, FastDiv(PW), FastDiv(PH), FastDiv(npix), npix, total);
  OKL_LAUNCHED(ctx, "conv_relu_pool2_fwd");
  return OKL_OK;
}

extern "C" int okl_conv_relu_pool_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* pooled_g, const void* mask,
                                        void* dW, void* db) {
  OKL_REQUIRE(ctx && d && x && pooled_g && mask, "NULL argument");
  OKL_REQUIRE(okl_conv_relu_pool_supported(d), "okl_conv_relu_pool_wgrad: unsupported shape");
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  const int K = g.C * 9;
  const int KT = K + 1 <= 12 ? 12 : 28;
  const int PH = g.Oe[1] / 2, PW = g.Oe[2] / 2;
  const int Mpool = g.N * PH * PW;
  const int ntiles = (Mpool + WG_PT - 1) / WG_PT;
  const int gy_ = (g.O + CP_OB - 1) / CP_OB;
  const int resident = g.C == 1 ? 5 : 3;                  // CTAs per SM the register budget allows (__launch_bounds__)
  int gx = (resident * ctx->sm_count) / gy_;              // one wave: every tile gets its own resident block at MNIST sizes
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  s = okl_ensure_workspace(ctx, (size_t)gx * gy_ * CP_OB * KT * sizeof(float));
  if (s != OKL_OK) return s;
  float* partial = (float*)ctx->workspace;
  dim3 grid(gx, gy_);
  const bool pad = g.P[1] || g.P[2];
#define OKL_LAUNCH_CRPW(C_, KT_, PAD_)                                                                                               \
  okl_launch(ctx, conv_relu_pool2_wgrad<3, 3, C_, KT_, PAD_>, grid, dim3(128), 0, (const float*)x, (const float*)pooled_g,            \
             (const unsigned char*)mask, partial, g, K, PH, PW, FastDiv(PW), FastDiv(PH), Mpool, ntiles)
  if (g.C == 1) { if (pad) OKL_LAUNCH_CRPW(1, 12, true); else OKL_LAUNCH_CRPW(1, 12, false); }
  else { if (pad) OKL_LAUNCH_CRPW(3, 28, true); else OKL_LAUNCH_CRPW(3, 28, false); }
#undef OKL_LAUNCH_CRPW
  OKL_LAUNCHED(ctx, "conv_relu_pool2_wgrad");
  const int total = g.O * (K + 1);
  okl_launch(ctx, conv_smallk_wgrad_reduce, dim3((total * 32 + 255) / 256), dim3(256), 0, (const float*)partial, (float*)dW, (float*)db, g.O, K, KT,
             gx, CP_OB, gy_);
  OKL_LAUNCHED(ctx, "conv_smallk_wgrad_reduce");
  return OKL_OK;
}
