// Exact-fp32 (FFMA) GEMM-shaped kernels: the OKL_FP32 parity mode (<= 1e-5 relative vs the fp64 oracle) and the
// fallback for shapes the tcgen05 kernels do not take (strided convs, tiny channel counts, odd sizes).
//
// One tiled kernel, `igemm_ffma`, does C[M,N] = sum_k A(m,k) * B(k,n) with A and B produced by *loader functors*:
// plain strided matrices for DenseLayer (okl.py:29-47) and on-the-fly im2col gathers for Conv1D/2D/3D fprop, dgrad
// and wgrad (okl.py:691-754, 790-883, 948-1039) - the im2col matrix is never materialised (implicit GEMM).
// Tile 64x64x16, 256 threads, 4x4 outputs per thread, register-prefetch double buffering, optional split-K
// through a workspace followed by a fixed-order (deterministic) reduction that applies the epilogue.
#include "common.cuh"
#include "conv_geom.cuh"

using namespace okl_conv;

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case OKL_ACT_RELU: return v > 0.f ? v : 0.f;
    case OKL_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case OKL_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------------ loaders
// Strided matrix: element (i, k) at p[i*ld_i + k*ld_k].  KC = the k index is the contiguous one.
template <bool KC>
struct MatA {
  static constexpr bool kContig = KC;
  const float* p; long long ld_m, ld_k; int M, K;
  struct Row { long long off; bool ok; };
  __device__ __forceinline__ Row row(int m) const { return {m * ld_m, m < M}; }
  __device__ __forceinline__ float get(const Row& r, int k) const { return (r.ok && k < K) ? __ldg(p + r.off + k * ld_k) : 0.f; }
};
template <bool NC>
struct MatB {
  static constexpr bool nContig = NC;
  const float* p; long long ld_k, ld_n; int K, N;
  struct Col { long long off; bool ok; };
  __device__ __forceinline__ Col col(int n) const { return {n * ld_n, n < N}; }
  __device__ __forceinline__ float get(const Col& c, int k) const { return (c.ok && k < K) ? __ldg(p + c.off + k * ld_k) : 0.f; }
};

struct FpropA {                       // A(m = out pixel, k = (c,taps))
  static constexpr bool kContig = false;
  ConvPatch cp;
  using Row = PixCoord;
  __device__ __forceinline__ Row row(int m) const { return cp.pix(m); }
  __device__ __forceinline__ float get(const Row& r, int k) const { return cp.at(r, k); }
};
struct FpropB {                       // B(k, o) = W[o*K + k]
  static constexpr bool nContig = false;
  const float* W; int K, O;
  struct Col { long long off; bool ok; };
  __device__ __forceinline__ Col col(int n) const { return {(long long)n * K, n < O}; }
  __device__ __forceinline__ float get(const Col& c, int k) const { return (c.ok && k < K) ? __ldg(W + c.off + k) : 0.f; }
};
struct WgradA {                       // A'(o, k' = out pixel) = g[n, o, opos]
  static constexpr bool kContig = true;
  const float* gp; ConvGeom g; int Mpix;
  struct Row { long long off; bool ok; };
  __device__ __forceinline__ Row row(int m) const { return {m * g.ys[1], m < g.O}; }
  __device__ __forceinline__ float get(const Row& r, int k) const {
    if (!r.ok || k >= Mpix) return 0.f;
    unsigned t, ow, oh, od;
    g.dOW.divmod((unsigned)k, t, ow); g.dOH.divmod(t, t, oh); g.dOD.divmod(t, t, od);
    return __ldg(gp + r.off + t * g.ys[0] + od * g.ys[2] + oh * g.ys[3] + ow * g.ys[4]);
  }
};
struct WgradB {                       // B'(k' = out pixel, n' = (c,taps)) = im2col gather
  static constexpr bool nContig = false;
  ConvPatch cp;
  struct Col { int kk; };
  __device__ __forceinline__ Col col(int n) const { return {n}; }
  __device__ __forceinline__ float get(const Col& c, int k) const { return cp.at(cp.pix(k), c.kk); }
};
struct DgradA {                       // A(m = in pixel (n,id,ih,iw), k = (o,taps)) = g[n, o, (i + p - t)/s] when divisible
  static constexpr bool kContig = false;
  const float* gp; ConvGeom g; int Mpix, K;
  using Row = PixCoord;
  __device__ __forceinline__ Row row(int m) const {
    PixCoord r; r.ok = m < Mpix;
    unsigned t, iw, ih, id;
    g.dIW.divmod((unsigned)m, t, iw); g.dIH.divmod(t, t, ih); g.dID.divmod(t, t, id);
    r.n = (int)t; r.d = (int)id + g.P[0]; r.h = (int)ih + g.P[1]; r.w = (int)iw + g.P[2];
    return r;
  }
  __device__ __forceinline__ float get(const Row& r, int k) const {
    if (!r.ok || k >= K) return 0.f;
    unsigned o, tw, th, td;
    g.dKW.divmod((unsigned)k, o, tw); g.dKH.divmod(o, o, th); g.dKD.divmod(o, o, td);
    int nd_ = r.d - (int)td, nh = r.h - (int)th, nw = r.w - (int)tw;
    if (nd_ < 0 || nh < 0 || nw < 0) return 0.f;
    unsigned od, oh, ow, rd, rh, rw;
    g.dS0.divmod((unsigned)nd_, od, rd); g.dS1.divmod((unsigned)nh, oh, rh); g.dS2.divmod((unsigned)nw, ow, rw);
    if (rd | rh | rw) return 0.f;
    if (od >= (unsigned)g.Oe[0] || oh >= (unsigned)g.Oe[1] || ow >= (unsigned)g.Oe[2]) return 0.f;
    return __ldg(gp + r.n * g.ys[0] + (long long)o * g.ys[1] + od * g.ys[2] + oh * g.ys[3] + ow * g.ys[4]);
  }
};
struct DgradB {                       // B(k = (o,taps), c) = W[(o*C + c)*T + tap]
  static constexpr bool nContig = false;
  const float* W; ConvGeom g; int K;
  struct Col { long long off; bool ok; };
  __device__ __forceinline__ Col col(int n) const { return {(long long)n * g.T, n < g.C}; }
  __device__ __forceinline__ float get(const Col& c, int k) const {
    if (!c.ok || k >= K) return 0.f;
    unsigned o, tap;
    g.dT.divmod((unsigned)k, o, tap);
    return __ldg(W + (long long)o * g.C * g.T + c.off + tap);
  }
};

// ------------------------------------------------------------------------------------------------ epilogues
struct DenseEpi {                     // Y[m*N + n] = act(v + b[n])
  static constexpr bool mContig = false;
  float* Y; const float* b; int N; int act; const float* mask;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    if (b) v += __ldg(b + n);
    v = apply_act(v, act);
    if (mask && !(__ldg(mask + (long long)m * N + n) > 0.f)) v = 0.f;
    Y[(long long)m * N + n] = v;
  }
};
struct ConvOutEpi {                   // y[n, o, opos] = act(v + b[o]) through the y strides
  static constexpr bool mContig = true;
  float* y; const float* b; ConvGeom g; int act; bool out_is_input_space; const float* mask;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    unsigned t, w, h, d;
    if (out_is_input_space) { g.dIW.divmod((unsigned)m, t, w); g.dIH.divmod(t, t, h); g.dID.divmod(t, t, d); }
    else { g.dOW.divmod((unsigned)m, t, w); g.dOH.divmod(t, t, h); g.dOD.divmod(t, t, d); }
    const long long* st = out_is_input_space ? g.xs : g.ys;
    if (b) v += __ldg(b + n);
    v = apply_act(v, act);
    const long long off = t * st[0] + (long long)n * st[1] + d * st[2] + h * st[3] + w * st[4];
    if (mask && !(__ldg(mask + off) > 0.f)) v = 0.f;
    y[off] = v;
  }
};

// ------------------------------------------------------------------------------------------------ kernel
template <class AL, class BL, class EP>
__global__ void __launch_bounds__(NT) igemm_ffma(AL al, BL bl, EP ep, int M, int N, int K, int kchunk, float* __restrict__ ws) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);

  // loader mappings: 4 elements of A and 4 of B per thread per k-tile
  int a_m[4], a_k[4], b_n[4], b_k[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (AL::kContig) { a_k[i] = tid & 15; a_m[i] = (tid >> 4) + 16 * i; } else { a_m[i] = tid & 63; a_k[i] = (tid >> 6) + 4 * i; }
    if (BL::nContig) { b_n[i] = tid & 63; b_k[i] = (tid >> 6) + 4 * i; } else { b_k[i] = tid & 15; b_n[i] = (tid >> 4) + 16 * i; }
  }
  typename AL::Row arow[AL::kContig ? 4 : 1];
  typename BL::Col bcol[BL::nContig ? 1 : 4];
#pragma unroll
  for (int i = 0; i < (AL::kContig ? 4 : 1); i++) arow[i] = al.row(m0 + a_m[i]);
#pragma unroll
  for (int i = 0; i < (BL::nContig ? 1 : 4); i++) bcol[i] = bl.col(n0 + b_n[i]);

  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int ka = k0 + a_k[i], kb = k0 + b_k[i];
      ra[i] = ka < kend ? al.get(arow[AL::kContig ? i : 0], ka) : 0.f;
      rb[i] = kb < kend ? bl.get(bcol[BL::nContig ? 0 : i], kb) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      As[buf][a_k[i]][a_m[i]] = ra[i];
      Bs[buf][b_k[i]][b_n[i]] = rb[i];
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  const int tm = EP::mContig ? tx : ty, tn = EP::mContig ? ty : tx;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  int buf = 0;
  if (kbeg < kend) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][k][tm * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  if (gridDim.z == 1) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int m = m0 + tm * 4 + i, n = n0 + tn * 4 + j;
        if (m < M && n < N) ep(m, n, acc[i][j]);
      }
  } else {
    float* w = ws + (size_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int m = m0 + tm * 4 + i, n = n0 + tn * 4 + j;
        if (m < M && n < N) w[EP::mContig ? (size_t)n * M + m : (size_t)m * N + n] = acc[i][j];
      }
  }
}

template <class EP>
__global__ void splitk_reduce(EP ep, const float* __restrict__ ws, int M, int N, int splits) {
  size_t total = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; z++) s += ws[(size_t)z * total + i];     // fixed order: deterministic
    int m, n;
    if (EP::mContig) { n = (int)(i / M); m = (int)(i - (size_t)n * M); } else { m = (int)(i / N); n = (int)(i - (size_t)m * N); }
    ep(m, n, s);
  }
}

template <class AL, class BL, class EP>
int launch_igemm(okl_ctx* ctx, const AL& al, const BL& bl, const EP& ep, int M, int N, int K, const char* name) {
  if (M <= 0 || N <= 0) return OKL_OK;
  int gm = (M + BM - 1) / BM, gn = (N + BN - 1) / BN;
  OKL_REQUIRE(gn <= 65535, "%s: N too large", name);
  long long tiles = (long long)gm * gn;
  int ktiles = (K + BK - 1) / BK;
  int splits = 1;
  long long target = 2LL * ctx->sm_count;
  if (tiles < target && ktiles >= 8) {
    splits = (int)((target + tiles - 1) / tiles);
    int max_splits = ktiles / 4;                       // >= 64 k per split
    if (splits > max_splits) splits = max_splits;
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
  }
  int kchunk = ((ktiles + splits - 1) / splits) * BK;
  splits = (K + kchunk - 1) / kchunk;
  if (splits < 1) splits = 1;
  float* ws = nullptr;
  if (splits > 1) {
    int s = okl_ensure_workspace(ctx, (size_t)splits * M * N * sizeof(float));
    if (s != OKL_OK) return s;
    ws = (float*)ctx->workspace;
  }
  dim3 grid(gm, gn, splits);
  igemm_ffma<AL, BL, EP><<<grid, NT, 0, ctx->stream>>>(al, bl, ep, M, N, K, kchunk, ws);
  OKL_LAUNCHED(ctx, name);
  if (splits > 1) {
    size_t total = (size_t)M * N;
    splitk_reduce<EP><<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>(ep, ws, M, N, splits);
    OKL_LAUNCHED(ctx, "splitk_reduce");
  }
  return OKL_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ dense
int okl_ffma_dense_fwd(okl_ctx* ctx, const float* X, const float* W, const float* b, float* Y, int M, int K, int N, int act) {
  MatA<true> a{X, K, 1, M, K};
  MatB<true> bb{W, N, 1, K, N};
  DenseEpi ep{Y, b, N, act, nullptr};
  return launch_igemm(ctx, a, bb, ep, M, N, K, "ffma_dense_fwd");
}
int okl_ffma_dense_dw(okl_ctx* ctx, const float* X, const float* G, float* dW, int M, int K, int N) {
  // dW[K,N] = X^T G : A'(m'=kin, k'=batch) = X[batch*K + kin]
  MatA<false> a{X, 1, K, K, M};
  MatB<true> bb{G, N, 1, M, N};
  DenseEpi ep{dW, nullptr, N, OKL_ACT_NONE, nullptr};
  return launch_igemm(ctx, a, bb, ep, K, N, M, "ffma_dense_dw");
}
int okl_ffma_dense_dx(okl_ctx* ctx, const float* G, const float* W, float* dX, int M, int K, int N, const float* mask) {
  // dX[M,K] = G W^T : B'(k'=nout, n'=kin) = W[kin*N + nout]
  MatA<true> a{G, N, 1, M, N};
  MatB<false> bb{W, 1, N, N, K};
  DenseEpi ep{dX, nullptr, K, OKL_ACT_NONE, mask};
  return launch_igemm(ctx, a, bb, ep, M, K, N, "ffma_dense_dx");
}

// ------------------------------------------------------------------------------------------------ conv
extern "C" int okl_conv_out_extent(const okl_conv_desc* d, int out[3]) {
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  for (int i = 0; i < d->nd; i++) out[i] = g.Oe[3 - d->nd + i];
  return OKL_OK;
}

int okl_ffma_conv_fprop(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* W, const float* b, float* y) {
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  int Mpix = g.N * g.Oe[0] * g.Oe[1] * g.Oe[2], K = g.C * g.T;
  FpropA a{ConvPatch{x, g, Mpix, K}};
  FpropB bb{W, K, g.O};
  ConvOutEpi ep{y, b, g, d->act, false, nullptr};
  return launch_igemm(ctx, a, bb, ep, Mpix, g.O, K, "ffma_conv_fprop");
}

int okl_ffma_conv_dgrad(okl_ctx* ctx, const okl_conv_desc* d, const float* gp, const float* W, float* dx, const float* mask) {
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  int Mpix = g.N * g.I[0] * g.I[1] * g.I[2], K = g.O * g.T;
  DgradA a{gp, g, Mpix, K};
  DgradB bb{W, g, K};
  ConvOutEpi ep{dx, nullptr, g, OKL_ACT_NONE, true, mask};
  return launch_igemm(ctx, a, bb, ep, Mpix, g.C, K, "ffma_conv_dgrad");
}

int okl_ffma_conv_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const float* x, const float* gp, float* dW) {
  ConvGeom g;
  int s = make_geom(d, g);
  if (s != OKL_OK) return s;
  int Mpix = g.N * g.Oe[0] * g.Oe[1] * g.Oe[2], KK = g.C * g.T;
  WgradA a{gp, g, Mpix};
  WgradB bb{ConvPatch{x, g, Mpix, KK}};
  DenseEpi ep{dW, nullptr, KK, OKL_ACT_NONE, nullptr};
  return launch_igemm(ctx, a, bb, ep, g.O, KK, Mpix, "ffma_conv_wgrad");
}
