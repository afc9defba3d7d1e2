// Convolution geometry shared by the FFMA implicit-GEMM kernels and the small-K direct kernels: fast integer division,
// the nd<=3 geometry normalised to three spatial dims, and the on-the-fly im2col gather (never materialised).
#pragma once
#include "common.cuh"

namespace okl_conv {

struct FastDiv {            // n / d for 0 <= n < 2^31, 1 <= d < 2^31
  unsigned d, mul, shift;
  FastDiv() : d(1), mul(0), shift(0) {}
  explicit FastDiv(unsigned dd) : d(dd) {
    if (dd == 1) { mul = 0; shift = 0; return; }
    unsigned s = 0;
    while ((1ull << s) < dd) s++;
    shift = s;
    mul = (unsigned)((((1ull << s) - dd) << 32) / dd + 1);
  }
  __device__ __forceinline__ unsigned div(unsigned n) const { return d == 1 ? n : (__umulhi(n, mul) + n) >> shift; }
  __device__ __forceinline__ void divmod(unsigned n, unsigned& q, unsigned& r) const { q = div(n); r = n - q * d; }
};

// Geometry of an nd<=3 convolution normalised to 3 spatial dims (leading dims of extent 1).
struct ConvGeom {
  int N, C, O;
  int I[3], Kk[3], S[3], P[3], Oe[3];
  long long xs[5], ys[5];           // strides of x / y for (n, c, d, h, w)
  int T;                            // taps = prod(Kk)
  FastDiv dOW, dOH, dOD, dIW, dIH, dID, dKW, dKH, dKD, dT, dS0, dS1, dS2;
};

struct PixCoord { int n, d, h, w; bool ok; };

// im2col gather: row = output pixel (n,od,oh,ow), k = (c,td,th,tw) -> x[n, c, od*s+td-p, ...]   (fprop A, wgrad B)
struct ConvPatch {
  const float* x; ConvGeom g; int Mpix, K;
  __device__ __forceinline__ PixCoord pix(int m) const {
    PixCoord r; r.ok = m < Mpix;
    unsigned t, ow, oh, od;
    g.dOW.divmod((unsigned)m, t, ow); g.dOH.divmod(t, t, oh); g.dOD.divmod(t, t, od);
    r.n = (int)t; r.d = (int)od * g.S[0] - g.P[0]; r.h = (int)oh * g.S[1] - g.P[1]; r.w = (int)ow * g.S[2] - g.P[2];
    return r;
  }
  __device__ __forceinline__ float at(const PixCoord& r, int k) const {
    if (!r.ok || k >= K) return 0.f;
    unsigned u, tw, th, td;
    g.dKW.divmod((unsigned)k, u, tw); g.dKH.divmod(u, u, th); g.dKD.divmod(u, u, td);
    int id = r.d + (int)td, ih = r.h + (int)th, iw = r.w + (int)tw;
    if ((unsigned)id >= (unsigned)g.I[0] || (unsigned)ih >= (unsigned)g.I[1] || (unsigned)iw >= (unsigned)g.I[2]) return 0.f;
    return __ldg(x + r.n * g.xs[0] + (long long)u * g.xs[1] + id * g.xs[2] + ih * g.xs[3] + iw * g.xs[4]);
  }
};

inline int make_geom(const okl_conv_desc* d, ConvGeom& g) {
  OKL_REQUIRE(d, "conv desc is NULL");
  OKL_REQUIRE(d->nd >= 1 && d->nd <= 3, "conv nd must be 1, 2 or 3 (got %d)", d->nd);
  OKL_REQUIRE(d->N >= 0 && d->C >= 1 && d->O >= 1, "conv N/C/O invalid (%d,%d,%d)", d->N, d->C, d->O);
  g.N = d->N; g.C = d->C; g.O = d->O;
  int lead = 3 - d->nd;
  for (int i = 0; i < 3; i++) { g.I[i] = 1; g.Kk[i] = 1; g.S[i] = 1; g.P[i] = 0; }
  for (int i = 0; i < d->nd; i++) {
    OKL_REQUIRE(d->in[i] >= 1 && d->k[i] >= 1 && d->s[i] >= 1 && d->p[i] >= 0, "conv dim %d invalid", i);
    g.I[lead + i] = d->in[i]; g.Kk[lead + i] = d->k[i]; g.S[lead + i] = d->s[i]; g.P[lead + i] = d->p[i];
  }
  for (int i = 0; i < 3; i++) {
    int e = g.I[i] + 2 * g.P[i] - g.Kk[i];
    OKL_REQUIRE(e >= 0, "conv kernel larger than padded input in dim %d", i);
    g.Oe[i] = e / g.S[i] + 1;
  }
  g.T = g.Kk[0] * g.Kk[1] * g.Kk[2];
  long long SI = (long long)g.I[0] * g.I[1] * g.I[2], SO = (long long)g.Oe[0] * g.Oe[1] * g.Oe[2];
  OKL_REQUIRE((long long)g.N * SI < (1LL << 31) && (long long)g.N * SO < (1LL << 31) && (long long)g.C * g.T < (1LL << 31) &&
                  (long long)g.O * g.T < (1LL << 31), "conv index space exceeds 2^31");
  auto strides = [](long long* st, int layout, int C, const int* e) {
    long long S = (long long)e[0] * e[1] * e[2];
    if (layout == OKL_NCX) { st[0] = C * S; st[1] = S; st[2] = (long long)e[1] * e[2]; st[3] = e[2]; st[4] = 1; }
    else { st[0] = C * S; st[1] = 1; st[2] = (long long)e[1] * e[2] * C; st[3] = (long long)e[2] * C; st[4] = C; }
  };
  OKL_REQUIRE(d->x_layout == OKL_NCX || d->x_layout == OKL_NXC, "bad x_layout");
  OKL_REQUIRE(d->y_layout == OKL_NCX || d->y_layout == OKL_NXC, "bad y_layout");
  strides(g.xs, d->x_layout, g.C, g.I);
  strides(g.ys, d->y_layout, g.O, g.Oe);
  g.dOW = FastDiv(g.Oe[2]); g.dOH = FastDiv(g.Oe[1]); g.dOD = FastDiv(g.Oe[0]);
  g.dIW = FastDiv(g.I[2]); g.dIH = FastDiv(g.I[1]); g.dID = FastDiv(g.I[0]);
  g.dKW = FastDiv(g.Kk[2]); g.dKH = FastDiv(g.Kk[1]); g.dKD = FastDiv(g.Kk[0]);
  g.dT = FastDiv(g.T);
  g.dS0 = FastDiv(g.S[0]); g.dS1 = FastDiv(g.S[1]); g.dS2 = FastDiv(g.S[2]);
  return OKL_OK;
}


}  // namespace okl_conv
