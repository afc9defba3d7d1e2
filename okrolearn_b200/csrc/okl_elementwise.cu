// Memory-bound kernels of the train step: activations, pooling, softmax, losses, bias-gradient reductions, layout
// permutation, dataset gather and the parameter updates.  All are coalesced along the contiguous dimension, use
// 128-bit accesses where alignment allows, grid-stride over a grid capped at a few waves of the 148 SMs, and
// reduce with warp shuffles in a fixed order (deterministic).  Reference citations per entry point are in okl.h.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ float act_f(float v, int act) {
  switch (act) {
    case OKL_ACT_RELU: return v > 0.f ? v : 0.f;                       // okl.py:103  max(0, x)
    case OKL_ACT_SIGMOID: return 1.f / (1.f + expf(-v));               // okl.py:509
    case OKL_ACT_TANH: return tanhf(v);                                // okl.py:2088
    default: return v;
  }
}
__device__ __forceinline__ float act_b(float ref, float g, int act) {
  switch (act) {
    case OKL_ACT_RELU: return ref > 0.f ? g : 0.f;                     // okl.py:107  g * (x > 0)
    case OKL_ACT_SIGMOID: return g * (ref * (1.f - ref));              // okl.py:514  ref = y
    case OKL_ACT_TANH: { float t = tanhf(ref); return g * (1.f - t * t); }   // okl.py:2092 ref = x
    default: return g;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the block; result valid in every thread.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = l < nw ? sh[l] : 0.f;
  r = warp_sum(r);
  return r;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------ activations
template <bool VEC>
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int act) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  if (VEC) {
    size_t n4 = n >> 2;
    for (size_t j = i; j < n4; j += st) {
      float4 v = reinterpret_cast<const float4*>(x)[j];
      v.x = act_f(v.x, act); v.y = act_f(v.y, act); v.z = act_f(v.z, act); v.w = act_f(v.w, act);
      reinterpret_cast<float4*>(y)[j] = v;
    }
    for (size_t j = (n4 << 2) + i; j < n; j += st) y[j] = act_f(x[j], act);
  } else {
    for (; i < n; i += st) y[i] = act_f(x[i], act);
  }
}
template <bool VEC>
__global__ void act_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ g, float* __restrict__ dx, size_t n, int act) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  if (VEC) {
    size_t n4 = n >> 2;
    for (size_t j = i; j < n4; j += st) {
      float4 r = reinterpret_cast<const float4*>(ref)[j], q = reinterpret_cast<const float4*>(g)[j], o;
      o.x = act_b(r.x, q.x, act); o.y = act_b(r.y, q.y, act); o.z = act_b(r.z, q.z, act); o.w = act_b(r.w, q.w, act);
      reinterpret_cast<float4*>(dx)[j] = o;
    }
    for (size_t j = (n4 << 2) + i; j < n; j += st) dx[j] = act_b(ref[j], g[j], act);
  } else {
    for (; i < n; i += st) dx[i] = act_b(ref[i], g[i], act);
  }
}
__global__ void binary_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n, size_t nb, int op) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    float bv = b[nb == n ? i : i % nb];
    o[i] = op == 0 ? a[i] + bv : a[i] * bv;
  }
}
__global__ void scale_kernel(float* __restrict__ x, size_t n, float a) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) x[i] *= a;
}

// ------------------------------------------------------------------------------------------------ permute / gather
// per image: src is (R x Cc) row-major, dst is (Cc x R) row-major
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, long long Cc) {
  __shared__ float tile[32][33];
  const float* s = src + (size_t)blockIdx.z * R * Cc;
  float* d = dst + (size_t)blockIdx.z * R * Cc;
  long long c0 = (long long)blockIdx.x * 32;
  int r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j; long long c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = s[(size_t)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    long long c = c0 + j; int r = r0 + threadIdx.x;
    if (r < R && c < Cc) d[(size_t)c * R + r] = tile[threadIdx.x][j];
  }
}
__global__ void gather_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, const int* __restrict__ idx,
                                   long long rows, long long row_elems) {
  okl_pdl_sync();
  size_t total = (size_t)rows * row_elems;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / row_elems, c = i - r * row_elems;
    dst[i] = src[(size_t)idx[r] * row_elems + c];
  }
}

__global__ void gather_rows4_kernel(float4* __restrict__ dst, const float4* __restrict__ src, const int* __restrict__ idx, long long rows,
                                    long long row4) {
  okl_pdl_sync();
  size_t total = (size_t)rows * row4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / row4, c = i - r * row4;
    dst[i] = src[(size_t)idx[r] * row4 + c];
  }
}

// one batch = the rows idx[0..rows) of the inputs (float4 granules) AND of the targets (32-bit words), in one launch.  The
// source may be page-locked HOST memory (zero-copy over PCIe, ~2 us per round trip): every thread keeps GB_U independent loads in
// flight and the grid covers the batch in a single pass, so the transfer is bound by the link, not by latency.
constexpr int GB_U = 8, GB_THREADS = 128;
__global__ void __launch_bounds__(GB_THREADS) gather_batch_kernel(float4* __restrict__ xd, const float4* __restrict__ xs, long long xrow4,
                                                           unsigned* __restrict__ td, const unsigned* __restrict__ ts, long long trow,
                                                           const int* __restrict__ idx, long long rows) {
  okl_pdl_sync();
  const size_t nx = (size_t)rows * xrow4, total = nx + (size_t)rows * trow;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t base = blockIdx.x * (size_t)blockDim.x + threadIdx.x; base < total; base += stride * GB_U) {
    float4 v[GB_U];
    unsigned u[GB_U];
#pragma unroll
    for (int j = 0; j < GB_U; j++) {
      const size_t i = base + j * stride;
      if (i < nx) {
        const size_t r = i / xrow4, c = i - r * xrow4;
        v[j] = xs[(size_t)idx[r] * xrow4 + c];
      } else if (i < total) {
        const size_t q = i - nx, r = q / trow, c = q - r * trow;
        u[j] = ts[(size_t)idx[r] * trow + c];
      }
    }
#pragma unroll
    for (int j = 0; j < GB_U; j++) {
      const size_t i = base + j * stride;
      if (i < nx) xd[i] = v[j];
      else if (i < total) td[i - nx] = u[j];
    }
  }
}

// loss read-back without a DMA copy: {value, sequence number} posted into page-locked host memory; the host polls the number
__global__ void loss_post_kernel(volatile float* __restrict__ host_slot, const float* __restrict__ dev_loss, unsigned seq) {
  host_slot[0] = dev_loss[0];
  __threadfence_system();
  reinterpret_cast<volatile unsigned*>(host_slot)[1] = seq;
}

// ------------------------------------------------------------------------------------------------ Unfold / Fold (okl.py:525-636)
// Unfold.forward: out[n, c*k*k + u*k + v, i*OW + j]; a window that sticks out of the image is CLIPPED and copied into the top-left
// corner of a zero k x k patch (what okl.py:588-603 does for padding > 0 - not true zero padding).
__global__ void unfold_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W, int k, int st, int pad, int OH, int OW,
                                  size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int L = OH * OW, kk = k * k;
    const int col = (int)(idx % L);
    size_t t = idx / L;
    const int row = (int)(t % ((size_t)C * kk));
    const int n = (int)(t / ((size_t)C * kk));
    const int c = row / kk, u = (row - c * kk) / k, v = row - c * kk - u * k;
    const int i = col / OW, j = col - i * OW;
    const int hs = i * st - pad, ws = j * st - pad;
    const int h0 = max(0, hs), h1 = min(H, hs + k), w0 = max(0, ws), w1 = min(W, ws + k);
    const int ih = h0 + u, iw = w0 + v;
    out[idx] = (ih < h1 && iw < w1) ? x[(((size_t)n * C + c) * H + ih) * W + iw] : 0.f;
  }
}
// Unfold.backward for padding == 0 (col2im): dx[n,c,h,w] = sum over the windows (i, j) and taps (u, v) that cover (h, w)
__global__ void unfold_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int C, int H, int W, int k, int st, int OH, int OW,
                                  size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    size_t t = idx / W;
    const int h = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const int L = OH * OW, kk = k * k;
    const float* gn = g + ((size_t)n * C + c) * kk * L;
    float s = 0.f;
    for (int u = 0; u < k; u++) {
      const int hh = h - u;
      if (hh < 0 || hh % st) continue;
      const int i = hh / st;
      if (i >= OH) continue;
      for (int v = 0; v < k; v++) {
        const int ww = w - v;
        if (ww < 0 || ww % st) continue;
        const int j = ww / st;
        if (j < OW) s += gn[(size_t)(u * k + v) * L + i * OW + j];
      }
    }
    dx[idx] = s;
  }
}
// Fold.forward (inverse == 0): (n, c, u, v, oh, ow) -> (n, c, oh, u, ow, v); Fold.backward (inverse == 1): the other way round
__global__ void fold_permute_kernel(const float* __restrict__ in, float* __restrict__ out, int k, int OH, int OW, int inverse, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    // idx enumerates the (n, c, oh, u, ow, v) order
    size_t t = idx;
    const int v = (int)(t % k); t /= k;
    const int ow = (int)(t % OW); t /= OW;
    const int u = (int)(t % k); t /= k;
    const int oh = (int)(t % OH); t /= OH;                  // t = n * C + c
    const size_t src = (((t * k + u) * k + v) * OH + oh) * (size_t)OW + ow;
    if (inverse) out[src] = in[idx]; else out[idx] = in[src];
  }
}

// ------------------------------------------------------------------------------------------------ pooling
struct PoolGeom { int N, C, H, W, OH, OW, ps, st; long long xs[4], ys[4]; bool cl; };   // strides for (n,c,h,w)

__device__ __forceinline__ void pool_decomp(size_t i, int C, int H, int W, bool cl, int& n, int& c, int& h, int& w) {
  if (cl) { c = (int)(i % C); i /= C; w = (int)(i % W); i /= W; h = (int)(i % H); n = (int)(i / H); }
  else { w = (int)(i % W); i /= W; h = (int)(i % H); i /= H; c = (int)(i % C); n = (int)(i / C); }
}

__global__ void pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g, int kind) {
  size_t total = (size_t)g.N * g.C * g.OH * g.OW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n, c, oh, ow;
    pool_decomp(i, g.C, g.OH, g.OW, g.cl, n, c, oh, ow);
    const float* base = x + n * g.xs[0] + c * g.xs[1] + (long long)(oh * g.st) * g.xs[2] + (long long)(ow * g.st) * g.xs[3];
    float acc = kind == OKL_POOL_MAX ? -INFINITY : 0.f;
    for (int a = 0; a < g.ps; a++)
      for (int b = 0; b < g.ps; b++) {
        float v = __ldg(base + a * g.xs[2] + b * g.xs[3]);
        acc = kind == OKL_POOL_MAX ? fmaxf(acc, v) : acc + v;
      }
    if (kind == OKL_POOL_AVG) acc /= (float)(g.ps * g.ps);            // np.mean, okl.py:1803
    y[i] = acc;                                                        // y is dense in its layout order
  }
}

// Gather form of okl.py:1753-1766 [R4] / 1809-1819 [R5]: each input position looks at the windows that contain it,
// in the reference's raster order (i then j).  Max: the last window whose max equals x assigns its gradient
// (ties -> every maximum of a window receives it; overlapping windows overwrite).  Avg: contributions add.
__global__ void pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gr, float* __restrict__ dx, PoolGeom g, int kind,
                                int relu_mask) {
  size_t total = (size_t)g.N * g.C * g.H * g.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n, c, h, w;
    pool_decomp(i, g.C, g.H, g.W, g.cl, n, c, h, w);
    int i_lo = h - g.ps + 1 <= 0 ? 0 : (h - g.ps + 1 + g.st - 1) / g.st, i_hi = min(h / g.st, g.OH - 1);
    int j_lo = w - g.ps + 1 <= 0 ? 0 : (w - g.ps + 1 + g.st - 1) / g.st, j_hi = min(w / g.st, g.OW - 1);
    const float* xb = x + n * g.xs[0] + c * g.xs[1];
    const float* gb = gr + n * g.ys[0] + c * g.ys[1];
    float out = 0.f;
    if (kind == OKL_POOL_AVG) {
      for (int oi = i_lo; oi <= i_hi; oi++)
        for (int oj = j_lo; oj <= j_hi; oj++) out += __ldg(gb + oi * g.ys[2] + oj * g.ys[3]) / (float)(g.ps * g.ps);
    } else {
      float xv = __ldg(xb + h * g.xs[2] + w * g.xs[3]);
      for (int oi = i_lo; oi <= i_hi; oi++)
        for (int oj = j_lo; oj <= j_hi; oj++) {
          const float* wb = xb + (long long)(oi * g.st) * g.xs[2] + (long long)(oj * g.st) * g.xs[3];
          float m = -INFINITY;
          for (int a = 0; a < g.ps; a++)
            for (int b = 0; b < g.ps; b++) m = fmaxf(m, __ldg(wb + a * g.xs[2] + b * g.xs[3]));
          if (xv == m) out = __ldg(gb + oi * g.ys[2] + oj * g.ys[3]);
        }
    }
    if (relu_mask && !(__ldg(xb + h * g.xs[2] + w * g.xs[3]) > 0.f)) out = 0.f;
    dx[i] = out;
  }
}

// 2x2 / stride 2 windows on channels-first tensors with even H, W: one thread per window, 64-bit accesses, no index math
// beyond one division per thread.  relu_mask fuses the backward of a ReLU that precedes the pooling layer
// (dx = 0 where x <= 0), saving one full pass over the largest activation tensor of the network.
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total, int OW, int W) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / OW;                       // r = (n*C + c)*OH + oh
    int ow = (int)(i - r * OW);
    const float* p = x + (r * 2) * (size_t)W + 2 * ow;
    float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + W);
    y[i] = fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
  }
}
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dx, size_t total, int OW,
                                    int W, int relu_mask) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / OW;
    int ow = (int)(i - r * OW);
    size_t off = (r * 2) * (size_t)W + 2 * ow;
    float2 a = *reinterpret_cast<const float2*>(x + off), b = *reinterpret_cast<const float2*>(x + off + W);
    float m = fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
    float gv = g[i];
    if (relu_mask && !(m > 0.f)) gv = 0.f;   // the window maximum is the only candidate; y <= 0 => relu'(x) = 0
    float2 da, db;
    da.x = a.x == m ? gv : 0.f; da.y = a.y == m ? gv : 0.f;
    db.x = b.x == m ? gv : 0.f; db.y = b.y == m ? gv : 0.f;
    *reinterpret_cast<float2*>(dx + off) = da;
    *reinterpret_cast<float2*>(dx + off + W) = db;
  }
}

// channels-last 2x2 / stride 2: one thread per (window, 4 channels), 128-bit accesses along C
__global__ void maxpool2_nxc_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total4, int C4, int OW, int OH, int W, int H) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const size_t n = r / OH;
    const float4* p = reinterpret_cast<const float4*>(x) + ((n * H + 2 * oh) * W + 2 * ow) * C4 + c;
    const float4 a = p[0], b = p[C4], d = p[(size_t)W * C4], e = p[(size_t)W * C4 + C4];
    float4 m;
    m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)); m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y));
    m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)); m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w));
    reinterpret_cast<float4*>(y)[i] = m;
  }
}
__device__ __forceinline__ float4 pool_route(float4 v, float4 m, float4 g) {
  return make_float4(v.x == m.x ? g.x : 0.f, v.y == m.y ? g.y : 0.f, v.z == m.z ? g.z : 0.f, v.w == m.w ? g.w : 0.f);
}
__global__ void maxpool2_nxc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dx, size_t total4, int C4,
                                        int OW, int OH, int W, int H, int relu_mask) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const size_t n = r / OH;
    const size_t off = ((n * H + 2 * oh) * W + 2 * ow) * C4 + c;
    const float4* p = reinterpret_cast<const float4*>(x) + off;
    const float4 a = p[0], b = p[C4], d = p[(size_t)W * C4], e = p[(size_t)W * C4 + C4];
    float4 m;
    m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)); m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y));
    m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)); m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w));
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    if (relu_mask) {
      if (!(m.x > 0.f)) gv.x = 0.f;
      if (!(m.y > 0.f)) gv.y = 0.f;
      if (!(m.z > 0.f)) gv.z = 0.f;
      if (!(m.w > 0.f)) gv.w = 0.f;
    }
    float4* q = reinterpret_cast<float4*>(dx) + off;
    q[0] = pool_route(a, m, gv); q[C4] = pool_route(b, m, gv);
    q[(size_t)W * C4] = pool_route(d, m, gv); q[(size_t)W * C4 + C4] = pool_route(e, m, gv);
  }
}

// ------------------------------------------------------------------------------------------------ softmax / losses
// one warp per row
__global__ void softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ p, int rows, int cols) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const float* xr = x + (size_t)r * cols;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += expf(xr[c] - m);
    s = warp_sum(s);
    for (int c = lane; c < cols; c += 32) {
      float v = expf(xr[c] - m) / s;
      if (isnan(v)) v = 0.f;                                           // np.nan_to_num, okl.py:370
      p[(size_t)r * cols + c] = v;
    }
  }
}
// dx = p * (g - sum_k g_k p_k)    == einsum('ijk,ik->ij', J, g) of okl.py:376-393
__global__ void softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ dx, int rows, int cols) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const float* pr = p + (size_t)r * cols;
    const float* gr = g + (size_t)r * cols;
    float d = 0.f;
    for (int c = lane; c < cols; c += 32) d += gr[c] * pr[c];
    d = warp_sum(d);
    for (int c = lane; c < cols; c += 32) dx[(size_t)r * cols + c] = pr[c] * (gr[c] - d);
  }
}

__device__ __forceinline__ float clip_p(float v) {
  // np.clip(p, 1e-12, 1 - 1e-12) (okl.py:1391); in fp32 the upper bound rounds to 1.0f
  return fminf(fmaxf(v, 1e-12f), 1.0f);
}
// grad over the whole grid; block 0 additionally reduces the loss (rows is at most a few thousand per step)
__global__ void ce_kernel(const float* __restrict__ p, const int* __restrict__ t, float* __restrict__ loss, float* __restrict__ grad,
                          int rows, int cols, float inv_denom) {
  __shared__ float sh[32];
  size_t total = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    float v = clip_p(p[i]);
    if (c == t[r]) v -= 1.f;
    grad[i] = v * inv_denom;
  }
  if (blockIdx.x == 0) {
    float s = 0.f;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
      int tc = t[r];
      float v = (tc >= 0 && tc < cols) ? clip_p(p[(size_t)r * cols + tc]) : 1.f;
      s += -logf(v);
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) loss[0] = rows > 0 ? s / (float)rows : 0.f;
  }
}

// Small problems (rows*cols <= 64K): ONE block does the loss, the CE gradient and - when asked - the backward of the softmax
// layer that produced p (dz = p * (g - sum_k g_k p_k), okl.py:376-393), one warp per row, and adds the batch loss to the
// epoch accumulator.  Fixed reduction order, no second launch.
__global__ void ce_fused_small_kernel(const float* __restrict__ p, const int* __restrict__ t, float* __restrict__ loss, float* __restrict__ grad,
                                      float* __restrict__ dz, float* __restrict__ loss_acc, int rows, int cols, float inv_denom) {
  __shared__ float sh[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float lsum = 0.f;
  for (int r = warp; r < rows; r += nw) {
    const int tc = t[r];
    const float* pr = p + (size_t)r * cols;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float pv = pr[c];
      float gv = clip_p(pv);
      if (c == tc) gv -= 1.f;
      gv *= inv_denom;
      grad[(size_t)r * cols + c] = gv;
      dot += gv * pv;
    }
    if (dz) {
      dot = warp_sum(dot);
      for (int c = lane; c < cols; c += 32) {
        const float pv = pr[c];
        float gv = clip_p(pv);
        if (c == tc) gv -= 1.f;
        gv *= inv_denom;
        dz[(size_t)r * cols + c] = pv * (gv - dot);
      }
    }
    if (lane == 0) lsum += -logf((tc >= 0 && tc < cols) ? clip_p(pr[tc]) : 1.f);
  }
  lsum = block_sum(lsum, sh);
  if (threadIdx.x == 0) {
    const float l = rows > 0 ? lsum / (float)rows : 0.f;
    loss[0] = l;
    if (loss_acc) loss_acc[0] += l;
  }
}

// Same contract for few classes (cols <= 32): one THREAD per row keeps the whole row in registers, so the 256 rows of a batch
// are 256 independent two-load chains instead of 8 sequential rows per warp.
template <int MAXC>
__global__ void ce_fused_rowthread_kernel(const float* __restrict__ p, const int* __restrict__ t, float* __restrict__ loss,
                                          float* __restrict__ grad, float* __restrict__ dz, float* __restrict__ loss_acc, int rows, int cols,
                                          float inv_denom) {
  __shared__ float sh[32];
  float lsum = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const int tc = t[r];
    const float* pr = p + (size_t)r * cols;
    float pv[MAXC], gv[MAXC];
    float dot = 0.f, pt = 1.f;
#pragma unroll
    for (int c = 0; c < MAXC; c++)
      if (c < cols) pv[c] = pr[c];
#pragma unroll
    for (int c = 0; c < MAXC; c++)
      if (c < cols) {
        float g = clip_p(pv[c]);
        if (c == tc) { pt = g; g -= 1.f; }
        g *= inv_denom;
        gv[c] = g;
        dot += g * pv[c];
        grad[(size_t)r * cols + c] = g;
      }
    if (dz) {
#pragma unroll
      for (int c = 0; c < MAXC; c++)
        if (c < cols) dz[(size_t)r * cols + c] = pv[c] * (gv[c] - dot);
    }
    lsum += -logf(pt);
  }
  lsum = block_sum(lsum, sh);
  if (threadIdx.x == 0) {
    const float l = rows > 0 ? lsum / (float)rows : 0.f;
    loss[0] = l;
    if (loss_acc) loss_acc[0] += l;
  }
}

// RELU: the outputs come from a ReLU whose backward is applied right here (g * (o > 0), okl.py:107): one pass over o instead of two
template <bool RELU>
__global__ void mse_kernel(const float* __restrict__ o, const float* __restrict__ t, float* __restrict__ partial, float* __restrict__ grad,
                           size_t n, float scale) {
  __shared__ float sh[32];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float ov = o[i], d = ov - t[i];
    s += d * d;
    grad[i] = (RELU && !(ov > 0.f)) ? 0.f : d * scale;                 // 2 (o - t) / denom, okl.py:1838
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// 128-bit variant (n % 4 == 0, 16-byte aligned pointers): two independent float4 pairs in flight per thread
template <bool RELU>
__global__ void __launch_bounds__(256) mse_kernel4(const float4* __restrict__ o, const float4* __restrict__ t, float* __restrict__ partial,
                                                   float4* __restrict__ grad, size_t n4, float scale) {
  __shared__ float sh[32];
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  auto one = [&](const float4 a, const float4 b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    s += dx * dx + dy * dy + dz * dz + dw * dw;
    return make_float4((RELU && !(a.x > 0.f)) ? 0.f : dx * scale, (RELU && !(a.y > 0.f)) ? 0.f : dy * scale,
                       (RELU && !(a.z > 0.f)) ? 0.f : dz * scale, (RELU && !(a.w > 0.f)) ? 0.f : dw * scale);
  };
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    const float4 a0 = o[i], b0 = t[i], a1 = o[i + stride], b1 = t[i + stride];
    grad[i] = one(a0, b0);
    grad[i + stride] = one(a1, b1);
  }
  if (i < n4) grad[i] = one(o[i], t[i]);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// Outputs channels-last (n, s, c) - as the tensor-core convolutions produce them - against targets in the caller's (n, c, s)
// layout: the target tile is transposed through shared memory, so neither the outputs nor the gradient are ever permuted.
constexpr int MSE_NXC_TILES = 8;
template <bool RELU>
__global__ void mse_nxc_kernel(const float* __restrict__ o, const float* __restrict__ t, float* __restrict__ partial, float* __restrict__ grad,
                               int Cc, long long S, float scale) {
  __shared__ float tile[32][33];
  __shared__ float sh[32];
  const size_t img = (size_t)blockIdx.z * Cc * S;
  const int c0 = blockIdx.x * 32;
  float acc = 0.f;
  for (int st = 0; st < MSE_NXC_TILES; st++) {             // several 32-position tiles per block: fewer partials, less ramp
    const long long s0 = ((long long)blockIdx.y * MSE_NXC_TILES + st) * 32;
    if (s0 >= S) break;
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int c = c0 + j; const long long sp = s0 + threadIdx.x;
      tile[j][threadIdx.x] = (c < Cc && sp < S) ? t[img + (size_t)c * S + sp] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const long long sp = s0 + j; const int c = c0 + threadIdx.x;
      if (sp < S && c < Cc) {
        const size_t i = img + (size_t)sp * Cc + c;
        const float ov = o[i], d = ov - tile[threadIdx.x][j];
        acc += d * d;
        grad[i] = (RELU && !(ov > 0.f)) ? 0.f : d * scale;
      }
    }
  }
  acc = warp_sum(acc);                                   // a warp = one threadIdx.y row of the 32 x 8 block
  if (threadIdx.x == 0) sh[threadIdx.y] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)blockDim.y; w++) tot += sh[w];
    partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
  }
}
__global__ void final_sum_kernel(const float* __restrict__ partial, int count, float* __restrict__ out, float scale, float* __restrict__ acc) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s += partial[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    out[0] = s * scale;
    if (acc) acc[0] += s * scale;
  }
}
__global__ void add_scalar_kernel(float* __restrict__ acc, const float* __restrict__ v) { acc[0] += v[0]; }

// ------------------------------------------------------------------------------------------------ reductions for bias gradients
// stage 1: partial[chunk][c] = sum over the chunk's rows of G[r, c]  (G row-major rows x cols); 32 cols x 8 row-lanes per block
__global__ void colsum_stage1(const float* __restrict__ G, float* __restrict__ partial, long long rows, int cols, long long rows_per_chunk) {
  __shared__ float sh[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += G[(size_t)r * cols + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) a += sh[j][threadIdx.x];
    partial[(size_t)blockIdx.y * cols + c] = a;
  }
}
// single pass (rows small, or enough column blocks to fill the chip): out[c] = sum_r G[r, c]; 32 cols x 32 row-lanes per block
__global__ void colsum_single(const float* __restrict__ G, float* __restrict__ out, long long rows, int cols) {
  __shared__ float sh[32][33];
  okl_pdl_sync();
  int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < cols)
    for (long long r = threadIdx.y; r < rows; r += 32) s += G[(size_t)r * cols + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j++) a += sh[j][threadIdx.x];
    out[c] = a;
  }
}
// stage 1 for channels-first (N, C, S): partial[chunk][c] = sum_{n in chunk, s} g[n, c, s]
__global__ void chansum_stage1(const float* __restrict__ g, float* __restrict__ partial, int N, int C, long long S, int n_per_chunk) {
  __shared__ float sh[32];
  int c = blockIdx.x;
  int n0 = blockIdx.y * n_per_chunk, n1 = min(N, n0 + n_per_chunk);
  float s = 0.f;
  for (int n = n0; n < n1; n++) {
    const float* b = g + ((size_t)n * C + c) * S;
    for (long long i = threadIdx.x; i < S; i += blockDim.x) s += b[i];
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * C + c] = s;
}
__global__ void colsum_stage2(const float* __restrict__ partial, float* __restrict__ out, int chunks, int cols) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int k = 0; k < chunks; k++) s += partial[(size_t)k * cols + c];
  out[c] = s;
}

// ------------------------------------------------------------------------------------------------ parameter updates
template <bool VEC>
__global__ void accum_update_kernel(float* __restrict__ w, float* __restrict__ gacc, const float* __restrict__ g, size_t n, float lr,
                                    const float* __restrict__ lr_dev) {
  okl_pdl_sync();
  if (lr_dev) lr = __ldg(lr_dev);
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  if (VEC) {
    size_t n4 = n >> 2;
    for (size_t j = i; j < n4; j += st) {
      float4 a = reinterpret_cast<float4*>(gacc)[j], q = reinterpret_cast<const float4*>(g)[j], ww = reinterpret_cast<float4*>(w)[j];
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;                   // grad = grad + g      okl.py:39
      ww.x -= lr * a.x; ww.y -= lr * a.y; ww.z -= lr * a.z; ww.w -= lr * a.w;   // w -= lr * grad   okl.py:44
      reinterpret_cast<float4*>(gacc)[j] = a;
      reinterpret_cast<float4*>(w)[j] = ww;
    }
    for (size_t j = (n4 << 2) + i; j < n; j += st) { float a = gacc[j] + g[j]; gacc[j] = a; w[j] -= lr * a; }
  } else {
    for (; i < n; i += st) { float a = gacc[i] + g[i]; gacc[i] = a; w[i] -= lr * a; }
  }
}
struct MultiUpd {
  float* w[16];
  float* gacc[16];
  const float* g[16];
  unsigned long long n[16];
  __nv_bfloat16* w16[16];                 // optional bf16 shadow of the first n16 parameters (the weights the BF16 GEMMs read)
  unsigned long long n16[16];
};
// every parameter bucket of the network in one launch: blockIdx.y selects the bucket
__global__ void accum_update_multi_kernel(const MultiUpd mu, float lr, const float* __restrict__ lr_dev) {
  if (lr_dev) lr = __ldg(lr_dev);
  const int b = blockIdx.y;
  float* __restrict__ w = mu.w[b];
  float* __restrict__ ga = mu.gacc[b];
  const float* __restrict__ g = mu.g[b];
  const size_t n = mu.n[b], n4 = n >> 2;
  __nv_bfloat16* __restrict__ w16 = mu.w16[b];
  const size_t n16_4 = mu.n16[b] >> 2;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (size_t j = i; j < n4; j += st) {
    float4 a = reinterpret_cast<float4*>(ga)[j], q = reinterpret_cast<const float4*>(g)[j], ww = reinterpret_cast<float4*>(w)[j];
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    ww.x -= lr * a.x; ww.y -= lr * a.y; ww.z -= lr * a.z; ww.w -= lr * a.w;
    reinterpret_cast<float4*>(ga)[j] = a;
    reinterpret_cast<float4*>(w)[j] = ww;
    if (w16 && j < n16_4) {                 // the shadow is refreshed by the kernel that changes the master: no separate 6 B/param pass
      __nv_bfloat162 lo = __floats2bfloat162_rn(ww.x, ww.y), hi = __floats2bfloat162_rn(ww.z, ww.w);
      reinterpret_cast<uint2*>(w16)[j] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
  for (size_t j = (n4 << 2) + i; j < n; j += st) { float a = ga[j] + g[j]; ga[j] = a; w[j] -= lr * a; }
}
__global__ void sgd_kernel(float* __restrict__ w, float* __restrict__ v, const float* __restrict__ g, size_t n, float lr, float mu) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) { float nv = mu * v[i] - lr * g[i]; v[i] = nv; w[i] += nv; }   // optimizers.py:143-144
}
__global__ void adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, size_t n,
                            float lr, float b1, float b2, float eps, float c1, float c2) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    float gi = g[i];
    float mi = b1 * m[i] + (1.f - b1) * gi;                             // optimizers.py:35
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;                        // optimizers.py:36
    m[i] = mi; v[i] = vi;
    float mh = mi / c1, vh = vi / c2;                                   // optimizers.py:38-39
    w[i] -= lr * mh / (sqrtf(vh) + eps);                                // optimizers.py:41-42
  }
}

}  // namespace

// ================================================================================================ entry points
// ------------------------------------------------------------------------------------------------ parametric activations
// (okl.py:57-352, 2111-2153).  Formulas follow the reference line by line, including the clips it applies and the
// GELU derivative exactly as written there (0.5 tanh + ..., okl.py:300-301).
namespace {
__device__ __forceinline__ float act2_f(int kind, float x, float alpha, float scale) {
  switch (kind) {
    case OKL_ACT_ELU: return x > 0.f ? x : alpha * (expf(x) - 1.f);                                   // okl.py:70
    case OKL_ACT_SELU: return scale * (x > 0.f ? x : alpha * (expf(x) - 1.f));                        // okl.py:179
    case OKL_ACT_LEAKY: return x > 0.f ? x : alpha * x;                                               // okl.py:212 / 245
    case OKL_ACT_SWISH: return x / (1.f + expf(-fminf(fmaxf(x, -88.f), 88.f)));                       // okl.py:131-137
    case OKL_ACT_GELU: {                                                                              // okl.py:277-285
      const float a = fminf(fmaxf(0.7978845608028654f * (x + 0.044715f * x * x * x), -15.f), 15.f);
      return 0.5f * x * (1.f + tanhf(a));
    }
    case OKL_ACT_SOFTSIGN: return x / (1.f + fabsf(x));                                               // okl.py:322
    case OKL_ACT_HARDTANH: return fminf(fmaxf(x, -1.f), 1.f);                                         // okl.py:2135
    default: return x;
  }
}
__device__ __forceinline__ float act2_d(int kind, float x, float alpha, float scale) {
  switch (kind) {
    case OKL_ACT_ELU: return x > 0.f ? 1.f : alpha * expf(x);                                         // okl.py:74-78
    case OKL_ACT_SELU: return x > 0.f ? scale : scale * alpha * expf(x);                              // okl.py:183-184
    case OKL_ACT_LEAKY: return x > 0.f ? 1.f : alpha;                                                 // okl.py:216 / 249
    case OKL_ACT_SWISH: {                                                                             // okl.py:141-149
      const float sg = 1.f / (1.f + expf(-fminf(fmaxf(x, -88.f), 88.f))), sw = x * sg;
      return fminf(fmaxf(sw + sg * (1.f - sw), -1e9f), 1e9f);
    }
    case OKL_ACT_GELU: {                                                                              // okl.py:291-301 as written
      const float a = fminf(fmaxf(0.7978845608028654f * (x + 0.044715f * x * x * x), -15.f), 15.f);
      const float th = tanhf(a);
      return 0.5f * th + 0.5f * x * (1.f - th * th) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
    }
    case OKL_ACT_SOFTSIGN: { const float d = 1.f + fabsf(x); return 1.f / (d * d); }                 // okl.py:330
    case OKL_ACT_HARDTANH: return (x >= -1.f && x <= 1.f) ? 1.f : 0.f;                                // okl.py:2139
    case OKL_ACT_TANH_OUT: return 1.f - x * x;                // x is the tanh OUTPUT here (RNNLayer, okl.py:1667: 1 - hidden^2)
    default: return 1.f;
  }
}
__global__ void act2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int kind, float alpha, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = act2_f(kind, x[i], alpha, scale);
}
__global__ void act2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dx, size_t n, int kind, float alpha,
                                float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = g[i] * act2_d(kind, x[i], alpha, scale);
}
// PReLU: per-block partial of sum g * x * [x <= 0] (okl.py:250-251); combined in fixed order by final_sum_kernel
__global__ void prelu_alpha_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ partial, size_t n) {
  __shared__ float sh[32];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    if (!(xv > 0.f)) s += g[i] * xv;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// one warp per row: softmin forward (as written: exp(-(x + max)), okl.py:410-416), log-softmax forward / backward (okl.py:455-488)
__global__ void rowwise_kernel(int kind, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int rows, int cols) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* ar = a + (size_t)row * cols;
  float* orow = out + (size_t)row * cols;
  if (kind == 2) {                                         // log-softmax backward: g_j - sum_k exp(ls_k) g_k
    const float* gr = b + (size_t)row * cols;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += expf(ar[c]) * gr[c];
    s = warp_sum(s);
    for (int c = lane; c < cols; c += 32) orow[c] = gr[c] - s;
    return;
  }
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, ar[c]);
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
  float s = 0.f;
  if (kind == 0) {                                         // softmin
    for (int c = lane; c < cols; c += 32) s += expf(-(ar[c] + mx));
    s = warp_sum(s);
    for (int c = lane; c < cols; c += 32) {
      float v = expf(-(ar[c] + mx)) / s;
      if (isnan(v)) v = 0.f; else if (isinf(v)) v = v > 0.f ? 1.f : 0.f;      // nan_to_num(nan=0, posinf=1, neginf=0)
      orow[c] = v;
    }
  } else {                                                 // log-softmax
    for (int c = lane; c < cols; c += 32) s += expf(fminf(fmaxf(ar[c] - mx, -500.f), 500.f));
    s = warp_sum(s);
    const float ls = logf(s);
    for (int c = lane; c < cols; c += 32) orow[c] = (ar[c] - mx) - ls;
  }
}
}  // namespace

// ------------------------------------------------------------------------------------------------ LSTM cell pointwise parts (okl.py:1412-1534)
namespace {
// out[r, :] = [a[r, :ca] | b[r, :cb]]   (np.concatenate((inputs, prev_hidden), axis=1), okl.py:1452)
__global__ void concat_cols_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int ca, int cb, size_t total) {
  const int cw = ca + cb;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cw; const int c = (int)(i - r * cw);
    out[i] = c < ca ? a[r * ca + c] : b[r * cb + (c - ca)];
  }
}
// cell = f * c_prev + i * c~ ; hidden = o * tanh(cell)   (okl.py:1456-1458)
__global__ void lstm_fwd_kernel(const float* __restrict__ f, const float* __restrict__ ig, const float* __restrict__ cc, const float* __restrict__ o,
                                const float* __restrict__ cprev, float* __restrict__ cell, float* __restrict__ hidden, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float c = f[i] * cprev[i] + ig[i] * cc[i];
    cell[i] = c;
    hidden[i] = o[i] * tanhf(c);
  }
}
// gate pre-activation gradients and dL/dc_prev exactly as okl.py:1464-1483, 1510 write them
__global__ void lstm_bwd_kernel(const float* __restrict__ f, const float* __restrict__ ig, const float* __restrict__ cc, const float* __restrict__ o,
                                const float* __restrict__ cprev, const float* __restrict__ cell, const float* __restrict__ dh,
                                const float* __restrict__ dc, float* __restrict__ dzf, float* __restrict__ dzi, float* __restrict__ dzc,
                                float* __restrict__ dzo, float* __restrict__ dcprev, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float th = tanhf(cell[i]);
    const float d_o = dh[i] * th;
    const float dcell = dc[i] + dh[i] * o[i] * (1.f - th * th);
    dzf[i] = dcell * cprev[i] * f[i] * (1.f - f[i]);
    dzi[i] = dcell * cc[i] * ig[i] * (1.f - ig[i]);
    dzc[i] = dcell * ig[i] * (1.f - cc[i] * cc[i]);
    dzo[i] = d_o * o[i] * (1.f - o[i]);
    dcprev[i] = dcell * f[i];
  }
}
// out[r, c] = a0[r, off + c] + a1[...] + a2[...] + a3[...] for c < cols  (dL/dh_prev: the hidden-state columns of the four dX, okl.py:1506-1509)
__global__ void sum4_cols_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ a3,
                                 float* __restrict__ out, int ld, int off, int cols, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols; const int c = (int)(i - r * cols);
    const size_t j = r * ld + off + c;
    out[i] = ((a0[j] + a1[j]) + a2[j]) + a3[j];
  }
}
}  // namespace

namespace {
// ---- GRU cell pointwise parts (okl.py:1537-1643)
// mode 0: hidden = (1 - z) h_prev + z h~                                                     (okl.py:1569)
// mode 1: dzh = dh z (1 - h~^2) -> out0 ; dzz = dh (h~ - h_prev) z (1 - z) -> out1            (okl.py:1574-1575, 1579, 1583)
__global__ void gru_pointwise_kernel(int mode, const float* __restrict__ z, const float* __restrict__ hc, const float* __restrict__ hprev,
                                     const float* __restrict__ dh, float* __restrict__ out0, float* __restrict__ out1, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (mode == 0) out0[i] = (1.f - z[i]) * hprev[i] + z[i] * hc[i];
    else {
      out0[i] = dh[i] * z[i] * (1.f - hc[i] * hc[i]);
      out1[i] = dh[i] * (hc[i] - hprev[i]) * z[i] * (1.f - z[i]);
    }
  }
}
// dzr = dcomb_h[:, I:] * h_prev * r (1 - r)                                                   (okl.py:1576-1577, 1581)
__global__ void gru_dr_kernel(const float* __restrict__ dcomb_h, const float* __restrict__ hprev, const float* __restrict__ r, float* __restrict__ dzr,
                              int I, int H, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / H; const int c = (int)(i - row * H);
    dzr[i] = dcomb_h[row * (I + H) + I + c] * hprev[i] * r[i] * (1.f - r[i]);
  }
}
// dL/dh_prev and dL/dx from the three [x | h] gradients, in the reference's order of additions (okl.py:1597-1605)
__global__ void gru_dinputs_kernel(const float* __restrict__ dcz, const float* __restrict__ dcr, const float* __restrict__ dch,
                                   const float* __restrict__ r, const float* __restrict__ z, const float* __restrict__ dh, float* __restrict__ dx,
                                   float* __restrict__ dprev, int I, int H, size_t rows) {
  const int cw = I + H;
  const size_t total = rows * cw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / cw; const int c = (int)(i - row * cw);
    if (c < I) dx[row * I + c] = (dcz[i] + dcr[i]) + dch[i];
    else {
      const size_t j = row * H + (c - I);
      dprev[j] = ((dcz[i] + dcr[i]) + dch[i] * r[j]) + dh[j] * (1.f - z[j]);
    }
  }
}
}  // namespace

extern "C" int okl_gru_pointwise(okl_ctx* ctx, int mode, const void* z, const void* h_cand, const void* h_prev, const void* dh, void* out0,
                                 void* out1, size_t n) {
  OKL_REQUIRE(ctx && (mode == 0 || mode == 1), "bad gru_pointwise mode");
  if (n == 0) return OKL_OK;
  OKL_REQUIRE(z && h_cand && h_prev && out0 && (mode == 0 || (dh && out1)), "NULL argument");
  gru_pointwise_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>(mode, (const float*)z, (const float*)h_cand, (const float*)h_prev,
                                                                         (const float*)dh, (float*)out0, (float*)out1, n);
  OKL_LAUNCHED(ctx, "gru_pointwise");
  return OKL_OK;
}

extern "C" int okl_gru_dr(okl_ctx* ctx, const void* dcomb_h, const void* h_prev, const void* r, void* dzr, long long rows, int I, int H) {
  OKL_REQUIRE(ctx && rows >= 0 && I >= 0 && H >= 1, "bad gru_dr arguments");
  const size_t n = (size_t)rows * H;
  if (n == 0) return OKL_OK;
  OKL_REQUIRE(dcomb_h && h_prev && r && dzr, "NULL argument");
  gru_dr_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)dcomb_h, (const float*)h_prev, (const float*)r, (float*)dzr, I, H, n);
  OKL_LAUNCHED(ctx, "gru_dr");
  return OKL_OK;
}

extern "C" int okl_gru_dinputs(okl_ctx* ctx, const void* dcomb_z, const void* dcomb_r, const void* dcomb_h, const void* r, const void* z,
                               const void* dh, void* dx, void* dprev, long long rows, int I, int H) {
  OKL_REQUIRE(ctx && rows >= 0 && I >= 1 && H >= 1, "bad gru_dinputs arguments");
  if (rows == 0) return OKL_OK;
  OKL_REQUIRE(dcomb_z && dcomb_r && dcomb_h && r && z && dh && dx && dprev, "NULL argument");
  gru_dinputs_kernel<<<okl_grid_1d(ctx, (size_t)rows * (I + H), 256), 256, 0, ctx->stream>>>(
      (const float*)dcomb_z, (const float*)dcomb_r, (const float*)dcomb_h, (const float*)r, (const float*)z, (const float*)dh, (float*)dx,
      (float*)dprev, I, H, (size_t)rows);
  OKL_LAUNCHED(ctx, "gru_dinputs");
  return OKL_OK;
}

extern "C" int okl_concat_cols(okl_ctx* ctx, const void* a, const void* b, void* out, long long rows, int ca, int cb) {
  OKL_REQUIRE(ctx && rows >= 0 && ca >= 0 && cb >= 0, "bad concat arguments");
  const size_t total = (size_t)rows * (ca + cb);
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(out && (ca == 0 || a) && (cb == 0 || b), "NULL argument");
  concat_cols_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)a, (const float*)b, (float*)out, ca, cb, total);
  OKL_LAUNCHED(ctx, "concat_cols");
  return OKL_OK;
}

extern "C" int okl_lstm_pointwise_fwd(okl_ctx* ctx, const void* f, const void* i, const void* c_cand, const void* o, const void* c_prev,
                                      void* cell, void* hidden, size_t n) {
  OKL_REQUIRE(ctx && (n == 0 || (f && i && c_cand && o && c_prev && cell && hidden)), "NULL argument");
  if (n == 0) return OKL_OK;
  lstm_fwd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)f, (const float*)i, (const float*)c_cand, (const float*)o,
                                                                    (const float*)c_prev, (float*)cell, (float*)hidden, n);
  OKL_LAUNCHED(ctx, "lstm_fwd");
  return OKL_OK;
}

extern "C" int okl_lstm_pointwise_bwd(okl_ctx* ctx, const void* f, const void* i, const void* c_cand, const void* o, const void* c_prev,
                                      const void* cell, const void* dh, const void* dc, void* dzf, void* dzi, void* dzc, void* dzo,
                                      void* dc_prev, size_t n) {
  OKL_REQUIRE(ctx && (n == 0 || (f && i && c_cand && o && c_prev && cell && dh && dc && dzf && dzi && dzc && dzo && dc_prev)), "NULL argument");
  if (n == 0) return OKL_OK;
  lstm_bwd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)f, (const float*)i, (const float*)c_cand, (const float*)o,
                                                                    (const float*)c_prev, (const float*)cell, (const float*)dh, (const float*)dc,
                                                                    (float*)dzf, (float*)dzi, (float*)dzc, (float*)dzo, (float*)dc_prev, n);
  OKL_LAUNCHED(ctx, "lstm_bwd");
  return OKL_OK;
}

extern "C" int okl_sum4_cols(okl_ctx* ctx, const void* a0, const void* a1, const void* a2, const void* a3, void* out, long long rows, int ld,
                             int off, int cols) {
  OKL_REQUIRE(ctx && rows >= 0 && ld >= 1 && off >= 0 && cols >= 0 && off + cols <= ld, "bad sum4 arguments");
  const size_t total = (size_t)rows * cols;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(a0 && a1 && a2 && a3 && out, "NULL argument");
  sum4_cols_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)a0, (const float*)a1, (const float*)a2, (const float*)a3,
                                                                        (float*)out, ld, off, cols, total);
  OKL_LAUNCHED(ctx, "sum4_cols");
  return OKL_OK;
}

// ------------------------------------------------------------------------------------------------ BatchNorm (okl.py:1145-1222)
// (B, F) inputs, statistics over the batch axis.  One block per 32 features, 32 row lanes: column sums are formed in a fixed order.
namespace {
// stats[0:F] = mean, stats[F:2F] = population variance (np.mean / np.var, okl.py:1171-1172)
__global__ void bn_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, int B, int F) {
  __shared__ float sh[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < F) for (int r = threadIdx.y; r < B; r += 32) s += x[(size_t)r * F + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  float mean = 0.f;
  for (int q = 0; q < 32; q++) mean += sh[q][threadIdx.x];
  mean /= (float)B;
  __syncthreads();
  float v = 0.f;
  if (c < F) for (int r = threadIdx.y; r < B; r += 32) { const float d = x[(size_t)r * F + c] - mean; v += d * d; }
  sh[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0 && c < F) {
    float var = 0.f;
    for (int q = 0; q < 32; q++) var += sh[q][threadIdx.x];
    stats[c] = mean;
    stats[F + c] = var / (float)B;
  }
}
// running = momentum * batch + (1 - momentum) * running for mean and var (okl.py:1175-1176); n = 2F
__global__ void bn_running_kernel(float* __restrict__ running, const float* __restrict__ stats, int n, float momentum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) running[i] = momentum * stats[i] + (1.f - momentum) * running[i];
}
// y = gamma * (x - mean) / sqrt(max(var, eps) + eps) + beta, nan_to_num(nan = 0, +inf = 1, -inf = 0)  (okl.py:1182-1190)
__global__ void bn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mv, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ y, int F, size_t n, float eps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % F);
    float v = gamma[c] * ((x[i] - mv[c]) / sqrtf(fmaxf(mv[F + c], eps) + eps)) + beta[c];
    if (isnan(v)) v = 0.f; else if (isinf(v)) v = v > 0.f ? 1.f : 0.f;
    y[i] = v;
  }
}
// column sums the backward needs: red[0:F] = sum g, [F:2F] = sum g * xnorm, [2F:3F] = sum g * xc, [3F:4F] = sum xc
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mv, float* __restrict__ red,
                                     int B, int F, float eps) {
  __shared__ float sh[4][32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < F) {
    const float mean = mv[c], rstd = 1.f / sqrtf(fmaxf(mv[F + c], eps) + eps);
    for (int r = threadIdx.y; r < B; r += 32) {
      const float gv = g[(size_t)r * F + c], xc = x[(size_t)r * F + c] - mean;
      a0 += gv; a1 += gv * (xc * rstd); a2 += gv * xc; a3 += xc;
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = a0; sh[1][threadIdx.y][threadIdx.x] = a1;
  sh[2][threadIdx.y][threadIdx.x] = a2; sh[3][threadIdx.y][threadIdx.x] = a3;
  __syncthreads();
  if (threadIdx.y < 4 && c < F) {
    float t = 0.f;
    for (int q = 0; q < 32; q++) t += sh[threadIdx.y][q][threadIdx.x];
    red[threadIdx.y * F + c] = t;
  }
}
// dx exactly as okl.py:1197-1204 writes it, with the RUNNING variance in the chain rule (reference behaviour)
__global__ void bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mv, const float* __restrict__ running,
                              const float* __restrict__ gamma, const float* __restrict__ red, float* __restrict__ dx, int B, int F, size_t n, float eps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % F);
    const float rv = running[F + c] + eps, ga = gamma[c];
    const float inv = 1.f / sqrtf(rv);
    const float dvar = ga * red[2 * F + c] * -0.5f * powf(rv, -1.5f);
    const float dmean = ga * red[c] * -inv + dvar * (-2.f * red[3 * F + c] / (float)B);
    const float xc = x[i] - mv[c];
    dx[i] = g[i] * ga * inv + dvar * 2.f * xc / (float)B + dmean / (float)B;
  }
}
}  // namespace

extern "C" int okl_bn_fwd(okl_ctx* ctx, const void* x, const void* gamma, const void* beta, void* running_mean_var, void* batch_mean_var,
                          void* y, int B, int F, float eps, float momentum, int training) {
  OKL_REQUIRE(ctx && x && gamma && beta && running_mean_var && batch_mean_var && y && B >= 1 && F >= 1, "bad batch-norm arguments");
  const size_t n = (size_t)B * F;
  if (training) {
    bn_stats_kernel<<<(F + 31) / 32, dim3(32, 32), 0, ctx->stream>>>((const float*)x, (float*)batch_mean_var, B, F);
    OKL_LAUNCHED(ctx, "bn_stats");
    bn_running_kernel<<<(2 * F + 255) / 256, 256, 0, ctx->stream>>>((float*)running_mean_var, (const float*)batch_mean_var, 2 * F, momentum);
    OKL_LAUNCHED(ctx, "bn_running");
  } else {
    int s = okl_d2d(ctx, batch_mean_var, running_mean_var, (size_t)2 * F * sizeof(float));
    if (s != OKL_OK) return s;
  }
  bn_fwd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)x, (const float*)batch_mean_var, (const float*)gamma,
                                                                  (const float*)beta, (float*)y, F, n, eps);
  OKL_LAUNCHED(ctx, "bn_fwd");
  return OKL_OK;
}

extern "C" int okl_bn_bwd(okl_ctx* ctx, const void* x, const void* g, const void* gamma, const void* running_mean_var, const void* batch_mean_var,
                          void* dx, void* dgamma, void* dbeta, int B, int F, float eps) {
  OKL_REQUIRE(ctx && x && g && gamma && running_mean_var && batch_mean_var && dx && dgamma && dbeta && B >= 1 && F >= 1, "bad batch-norm arguments");
  int s = okl_ensure_workspace(ctx, (size_t)4 * F * sizeof(float));
  if (s != OKL_OK) return s;
  float* red = (float*)ctx->workspace;
  bn_bwd_reduce_kernel<<<(F + 31) / 32, dim3(32, 32), 0, ctx->stream>>>((const float*)x, (const float*)g, (const float*)batch_mean_var, red, B, F, eps);
  OKL_LAUNCHED(ctx, "bn_bwd_reduce");
  const size_t n = (size_t)B * F;
  bn_bwd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)x, (const float*)g, (const float*)batch_mean_var,
                                                                  (const float*)running_mean_var, (const float*)gamma, red, (float*)dx, B, F, n, eps);
  OKL_LAUNCHED(ctx, "bn_bwd");
  s = okl_d2d(ctx, dbeta, red, (size_t)F * sizeof(float));
  if (s != OKL_OK) return s;
  return okl_d2d(ctx, dgamma, red + F, (size_t)F * sizeof(float));
}

extern "C" int okl_act2_fwd(okl_ctx* ctx, int act, const void* x, void* y, size_t n, float alpha, float scale) {
  OKL_REQUIRE(ctx && (n == 0 || (x && y)), "NULL argument");
  OKL_REQUIRE(act >= OKL_ACT_ELU && act <= OKL_ACT_HARDTANH, "bad activation %d", act);
  if (n == 0) return OKL_OK;
  act2_fwd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)x, (float*)y, n, act, alpha, scale);
  OKL_LAUNCHED(ctx, "act2_fwd");
  return OKL_OK;
}

extern "C" int okl_act2_bwd(okl_ctx* ctx, int act, const void* x, const void* g, void* dx, size_t n, float alpha, float scale) {
  OKL_REQUIRE(ctx && (n == 0 || (x && g && dx)), "NULL argument");
  OKL_REQUIRE(act >= OKL_ACT_ELU && act <= OKL_ACT_TANH_OUT, "bad activation %d", act);
  if (n == 0) return OKL_OK;
  act2_bwd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)x, (const float*)g, (float*)dx, n, act, alpha, scale);
  OKL_LAUNCHED(ctx, "act2_bwd");
  return OKL_OK;
}

extern "C" int okl_prelu_alpha_grad(okl_ctx* ctx, const void* x, const void* g, size_t n, void* out_scalar) {
  OKL_REQUIRE(ctx && out_scalar && (n == 0 || (x && g)), "NULL argument");
  int grid = n ? okl_grid_1d(ctx, n, 256, 4) : 1;
  if (grid > 1024) grid = 1024;
  int s = okl_ensure_workspace(ctx, (size_t)grid * sizeof(float));
  if (s != OKL_OK) return s;
  prelu_alpha_kernel<<<grid, 256, 0, ctx->stream>>>((const float*)x, (const float*)g, (float*)ctx->workspace, n);
  OKL_LAUNCHED(ctx, "prelu_alpha");
  final_sum_kernel<<<1, 256, 0, ctx->stream>>>((const float*)ctx->workspace, grid, (float*)out_scalar, 1.f, nullptr);
  OKL_LAUNCHED(ctx, "final_sum");
  return OKL_OK;
}

extern "C" int okl_rowwise(okl_ctx* ctx, int kind, const void* a, const void* b, void* out, int rows, int cols) {
  OKL_REQUIRE(ctx && kind >= 0 && kind <= 2 && rows >= 0 && cols >= 1, "bad rowwise arguments");
  if (rows == 0) return OKL_OK;
  OKL_REQUIRE(a && out && (kind != 2 || b), "NULL argument");
  rowwise_kernel<<<(rows * 32 + 255) / 256, 256, 0, ctx->stream>>>(kind, (const float*)a, (const float*)b, (float*)out, rows, cols);
  OKL_LAUNCHED(ctx, "rowwise");
  return OKL_OK;
}

extern "C" int okl_act_fwd(okl_ctx* ctx, int act, const void* x, void* y, size_t n) {
  OKL_REQUIRE(ctx && (n == 0 || (x && y)), "NULL argument");
  OKL_REQUIRE(act >= OKL_ACT_NONE && act <= OKL_ACT_TANH, "bad activation %d", act);
  if (n == 0) return OKL_OK;
  bool vec = aligned16(x) && aligned16(y);
  int grid = okl_grid_1d(ctx, vec ? (n + 3) / 4 : n, 256);
  if (vec) act_fwd_kernel<true><<<grid, 256, 0, ctx->stream>>>((const float*)x, (float*)y, n, act);
  else act_fwd_kernel<false><<<grid, 256, 0, ctx->stream>>>((const float*)x, (float*)y, n, act);
  OKL_LAUNCHED(ctx, "act_fwd");
  return OKL_OK;
}

extern "C" int okl_act_bwd(okl_ctx* ctx, int act, const void* ref, const void* g, void* dx, size_t n) {
  OKL_REQUIRE(ctx && (n == 0 || (ref && g && dx)), "NULL argument");
  OKL_REQUIRE(act >= OKL_ACT_NONE && act <= OKL_ACT_TANH, "bad activation %d", act);
  if (n == 0) return OKL_OK;
  bool vec = aligned16(ref) && aligned16(g) && aligned16(dx);
  int grid = okl_grid_1d(ctx, vec ? (n + 3) / 4 : n, 256);
  if (vec) act_bwd_kernel<true><<<grid, 256, 0, ctx->stream>>>((const float*)ref, (const float*)g, (float*)dx, n, act);
  else act_bwd_kernel<false><<<grid, 256, 0, ctx->stream>>>((const float*)ref, (const float*)g, (float*)dx, n, act);
  OKL_LAUNCHED(ctx, "act_bwd");
  return OKL_OK;
}

extern "C" int okl_binary(okl_ctx* ctx, int op, const void* a, const void* b, void* out, size_t n, size_t nb) {
  OKL_REQUIRE(ctx && (n == 0 || (a && b && out)), "NULL argument");
  OKL_REQUIRE(op == 0 || op == 1, "bad binary op %d", op);
  OKL_REQUIRE(n == 0 || (nb >= 1 && nb <= n), "bad broadcast length");
  if (n == 0) return OKL_OK;
  binary_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((const float*)a, (const float*)b, (float*)out, n, nb, op);
  OKL_LAUNCHED(ctx, "binary");
  return OKL_OK;
}

extern "C" int okl_scale(okl_ctx* ctx, void* x, size_t n, float alpha) {
  OKL_REQUIRE(ctx && (n == 0 || x), "NULL argument");
  if (n == 0) return OKL_OK;
  scale_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((float*)x, n, alpha);
  OKL_LAUNCHED(ctx, "scale");
  return OKL_OK;
}

extern "C" int okl_permute(okl_ctx* ctx, void* dst, const void* src, int N, int C, long long S, int src_layout) {
  OKL_REQUIRE(ctx && dst && src, "NULL argument");
  OKL_REQUIRE(src_layout == OKL_NCX || src_layout == OKL_NXC, "bad layout");
  if (N <= 0 || C <= 0 || S <= 0) return OKL_OK;
  if (C == 1 || S == 1) return okl_d2d(ctx, dst, src, (size_t)N * C * S * sizeof(float));
  OKL_REQUIRE(N <= 65535, "okl_permute: N > 65535");
  // NCX -> NXC: per image (C x S) -> (S x C);  NXC -> NCX: (S x C) -> (C x S)
  int R = src_layout == OKL_NCX ? C : 0;
  long long Cc = src_layout == OKL_NCX ? S : C;
  long long Rl = src_layout == OKL_NCX ? C : S;
  OKL_REQUIRE(Rl < (1LL << 31), "okl_permute: dimension too large");
  R = (int)Rl;
  dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((R + 31) / 32), N);
  OKL_REQUIRE(grid.y <= 65535, "okl_permute: too many row tiles");
  transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>((const float*)src, (float*)dst, R, Cc);
  OKL_LAUNCHED(ctx, "permute");
  return OKL_OK;
}

extern "C" int okl_gather_rows(okl_ctx* ctx, void* dst, const void* src, const void* idx, long long rows, long long row_elems) {
  OKL_REQUIRE(ctx && (rows == 0 || (dst && src && idx)), "NULL argument");
  if (rows <= 0 || row_elems <= 0) return OKL_OK;
  if ((row_elems & 3) == 0 && aligned16(dst) && aligned16(src)) {
    gather_rows4_kernel<<<okl_grid_1d(ctx, (size_t)rows * (row_elems / 4), 256, 16), 256, 0, ctx->stream>>>(
        (float4*)dst, (const float4*)src, (const int*)idx, rows, row_elems / 4);
    OKL_LAUNCHED(ctx, "gather_rows4");
    return OKL_OK;
  }
  gather_rows_kernel<<<okl_grid_1d(ctx, (size_t)rows * row_elems, 256), 256, 0, ctx->stream>>>((float*)dst, (const float*)src,
                                                                                           (const int*)idx, rows, row_elems);
  OKL_LAUNCHED(ctx, "gather_rows");
  return OKL_OK;
}

extern "C" int okl_gather_batch(okl_ctx* ctx, void* x_dst, const void* x_src, long long x_row_elems, void* t_dst, const void* t_src,
                                long long t_row_elems, const void* idx, long long rows) {
  OKL_REQUIRE(ctx && (rows == 0 || idx), "NULL argument");
  if (rows <= 0) return OKL_OK;
  // a part whose consumer reads the dataset rows through the index itself (okl_*_rows entry points) is not gathered: x_dst / t_dst NULL
  if (!x_dst || !t_dst) {
    if (x_dst) { OKL_REQUIRE(x_src, "NULL argument"); return okl_gather_rows(ctx, x_dst, x_src, idx, rows, x_row_elems); }
    if (t_dst) { OKL_REQUIRE(t_src, "NULL argument"); return okl_gather_rows(ctx, t_dst, t_src, idx, rows, t_row_elems); }
    return OKL_OK;
  }
  OKL_REQUIRE(x_src && t_src, "NULL argument");
  if (x_row_elems > 0 && t_row_elems > 0 && (x_row_elems & 3) == 0 && aligned16(x_dst) && aligned16(x_src)) {
    const size_t total = (size_t)rows * (x_row_elems / 4 + t_row_elems);
    // the gather runs NEXT TO the compute kernels of the previous step (copy stream): few, small blocks with many loads in flight
    // per thread, so it does not take a register-file slot away from a whole wave of a compute kernel on every SM
    size_t blocks = (total + (size_t)GB_THREADS * GB_U * 4 - 1) / ((size_t)GB_THREADS * GB_U * 4);     // ~4 rounds per thread
    if (blocks < 16) blocks = 16;
    if (blocks > (size_t)ctx->sm_count * 4) blocks = (size_t)ctx->sm_count * 4;
    if (blocks * GB_THREADS * GB_U > total + GB_THREADS * GB_U) blocks = (total + GB_THREADS * GB_U - 1) / (GB_THREADS * GB_U);
    okl_launch(ctx, gather_batch_kernel, dim3((unsigned)blocks), dim3(GB_THREADS), 0, (float4*)x_dst, (const float4*)x_src,
               x_row_elems / 4, (unsigned*)t_dst, (const unsigned*)t_src, t_row_elems, (const int*)idx, rows);
    OKL_LAUNCHED(ctx, "gather_batch");
    return OKL_OK;
  }
  int s = okl_gather_rows(ctx, x_dst, x_src, idx, rows, x_row_elems);
  if (s != OKL_OK) return s;
  return okl_gather_rows(ctx, t_dst, t_src, idx, rows, t_row_elems);
}

int okl_loss_post(okl_ctx* ctx, float* host_slot, const void* dev_loss, unsigned seq) {
  loss_post_kernel<<<1, 1, 0, ctx->stream>>>(host_slot, (const float*)dev_loss, seq);
  OKL_LAUNCHED(ctx, "loss_post");
  return OKL_OK;
}

namespace {
__global__ void bias_add_kernel(float* __restrict__ y, const float* __restrict__ b, int C, long long S, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += b[(i / S) % C];
}
}  // namespace

// y[n, c, s] += b[c] in place (transposed convolution's bias, okl.py:838)
extern "C" int okl_bias_add(okl_ctx* ctx, void* y, const void* b, int N, int C, long long S) {
  OKL_REQUIRE(ctx && N >= 0 && C >= 1 && S >= 1, "bad bias_add shape");
  const size_t n = (size_t)N * C * S;
  if (n == 0) return OKL_OK;
  OKL_REQUIRE(y && b, "NULL argument");
  bias_add_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((float*)y, (const float*)b, C, S, n);
  OKL_LAUNCHED(ctx, "bias_add");
  return OKL_OK;
}

// db[c] = sum over n and s of g[n, c, s] (okl.py:917)
extern "C" int okl_bias_grad(okl_ctx* ctx, const void* g, void* db, int N, int C, long long S) {
  OKL_REQUIRE(ctx && g && db && N >= 1 && C >= 1 && S >= 1, "bad bias_grad arguments");
  return okl_channel_sum(ctx, (const float*)g, (float*)db, N, C, S, OKL_NCX);
}

// ------------------------------------------------------------------------------------------------ un-pooling / padding
namespace {
// Max / AverageUnpoolingLayer.forward (okl.py:2624-2637 / 2666-2679): every input pixel stamps a ps x ps block (value, or value /
// ps^2) at (i * stride, j * stride) with plain assignment in (i, j) loop order - where stamps overlap the LAST one wins, where
// none reaches the output stays 0.
__global__ void unpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int OH, int OW, int ps, int st, float scale,
                                  size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % OW);
    size_t t = idx / OW;
    const int h = (int)(t % OH);
    const size_t nc = t / OH;
    const int i = min(H - 1, h / st), j = min(W - 1, w / st);
    y[idx] = (h < i * st + ps && w < j * st + ps) ? x[(nc * H + i) * W + j] * scale : 0.f;
  }
}
// backward (okl.py:2640-2649): dx[i, j] = scale * sum of the ps x ps block of g at (i * stride, j * stride)
__global__ void unpool_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int H, int W, int OH, int OW, int ps, int st, float scale,
                                  size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W);
    size_t t = idx / W;
    const int i = (int)(t % H);
    const size_t nc = t / H;
    float s = 0.f;
    for (int u = 0; u < ps; u++)
      for (int v = 0; v < ps; v++) s += g[(nc * OH + i * st + u) * OW + j * st + v];
    dx[idx] = s * scale;
  }
}
// Zero / ReflectionPaddingLayer.forward (okl.py:1946-1950, 1966-1969): np.pad over the two spatial axes, mode constant 0 / reflect
__global__ void pad2d_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int pt, int pl, int OH, int OW, int reflect,
                             size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % OW);
    size_t t = idx / OW;
    const int h = (int)(t % OH);
    const size_t nc = t / OH;
    int ih = h - pt, iw = w - pl;
    float v = 0.f;
    if (reflect) {                                            // np.pad 'reflect': mirror without repeating the edge, periodically
      const int ph = 2 * (H - 1), pw = 2 * (W - 1);
      if (ph > 0) { ih %= ph; if (ih < 0) ih += ph; if (ih >= H) ih = ph - ih; } else ih = 0;
      if (pw > 0) { iw %= pw; if (iw < 0) iw += pw; if (iw >= W) iw = pw - iw; } else iw = 0;
      v = x[(nc * H + ih) * W + iw];
    } else if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
      v = x[(nc * H + ih) * W + iw];
    }
    y[idx] = v;
  }
}
// backward of both padding layers (okl.py:1953-1955): the crop g[:, :, top : top + h, left : left + w]
__global__ void crop2d_kernel(const float* __restrict__ g, float* __restrict__ dx, int GH, int GW, int top, int left, int h, int w, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % w);
    size_t t = idx / w;
    const int r = (int)(t % h);
    const size_t nc = t / h;
    dx[idx] = g[(nc * GH + top + r) * GW + left + c];
  }
}
}  // namespace

extern "C" int okl_unpool_fwd(okl_ctx* ctx, int kind, const void* x, void* y, long long NC, int H, int W, int pool, int stride) {
  OKL_REQUIRE(ctx && NC >= 0 && H >= 1 && W >= 1 && pool >= 1 && stride >= 1 && (kind == 0 || kind == 1), "bad unpool arguments");
  const int OH = (H - 1) * stride + pool, OW = (W - 1) * stride + pool;
  const size_t total = (size_t)NC * OH * OW;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(x && y, "NULL argument");
  unpool_fwd_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)x, (float*)y, H, W, OH, OW, pool, stride,
                                                                         kind == 0 ? 1.f : 1.f / (float)(pool * pool), total);
  OKL_LAUNCHED(ctx, "unpool_fwd");
  return OKL_OK;
}

extern "C" int okl_unpool_bwd(okl_ctx* ctx, int kind, const void* g, void* dx, long long NC, int H, int W, int pool, int stride) {
  OKL_REQUIRE(ctx && NC >= 0 && H >= 1 && W >= 1 && pool >= 1 && stride >= 1 && (kind == 0 || kind == 1), "bad unpool arguments");
  const int OH = (H - 1) * stride + pool, OW = (W - 1) * stride + pool;
  const size_t total = (size_t)NC * H * W;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(g && dx, "NULL argument");
  unpool_bwd_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)g, (float*)dx, H, W, OH, OW, pool, stride,
                                                                         kind == 0 ? 1.f : 1.f / (float)(pool * pool), total);
  OKL_LAUNCHED(ctx, "unpool_bwd");
  return OKL_OK;
}

extern "C" int okl_pad2d(okl_ctx* ctx, const void* x, void* y, long long NC, int H, int W, int top, int bottom, int left, int right, int reflect) {
  OKL_REQUIRE(ctx && NC >= 0 && H >= 1 && W >= 1 && top >= 0 && bottom >= 0 && left >= 0 && right >= 0, "bad pad2d arguments");
  OKL_REQUIRE(!reflect || ((top < H && bottom < H && left < W && right < W) || (H > 1 && W > 1)), "reflect padding needs extent > 1");
  const int OH = H + top + bottom, OW = W + left + right;
  const size_t total = (size_t)NC * OH * OW;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(x && y, "NULL argument");
  pad2d_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)x, (float*)y, H, W, top, left, OH, OW, reflect, total);
  OKL_LAUNCHED(ctx, "pad2d");
  return OKL_OK;
}

extern "C" int okl_crop2d(okl_ctx* ctx, const void* g, void* dx, long long NC, int GH, int GW, int top, int left, int h, int w) {
  OKL_REQUIRE(ctx && NC >= 0 && top >= 0 && left >= 0 && h >= 0 && w >= 0 && top + h <= GH && left + w <= GW, "bad crop2d arguments");
  const size_t total = (size_t)NC * h * w;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(g && dx, "NULL argument");
  crop2d_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)g, (float*)dx, GH, GW, top, left, h, w, total);
  OKL_LAUNCHED(ctx, "crop2d");
  return OKL_OK;
}

extern "C" int okl_unfold_fwd(okl_ctx* ctx, const void* x, void* out, int N, int C, int H, int W, int k, int stride, int pad) {
  OKL_REQUIRE(ctx && N >= 0 && C >= 1 && H >= 1 && W >= 1 && k >= 1 && stride >= 1 && pad >= 0, "bad unfold shape");
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  OKL_REQUIRE(H + 2 * pad >= k && W + 2 * pad >= k, "unfold: kernel larger than the padded input");
  const size_t total = (size_t)N * C * k * k * OH * OW;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(x && out, "NULL argument");
  unfold_fwd_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)x, (float*)out, C, H, W, k, stride, pad, OH, OW, total);
  OKL_LAUNCHED(ctx, "unfold_fwd");
  return OKL_OK;
}

extern "C" int okl_unfold_bwd(okl_ctx* ctx, const void* g, void* dx, int N, int C, int H, int W, int k, int stride) {
  OKL_REQUIRE(ctx && N >= 0 && C >= 1 && H >= k && W >= k && k >= 1 && stride >= 1, "bad unfold shape");
  const int OH = (H - k) / stride + 1, OW = (W - k) / stride + 1;
  const size_t total = (size_t)N * C * H * W;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(g && dx, "NULL argument");
  unfold_bwd_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)g, (float*)dx, C, H, W, k, stride, OH, OW, total);
  OKL_LAUNCHED(ctx, "unfold_bwd");
  return OKL_OK;
}

extern "C" int okl_fold_permute(okl_ctx* ctx, const void* in, void* out, long long NC, int k, int OH, int OW, int inverse) {
  OKL_REQUIRE(ctx && NC >= 0 && k >= 1 && OH >= 1 && OW >= 1, "bad fold shape");
  const size_t total = (size_t)NC * k * k * OH * OW;
  if (total == 0) return OKL_OK;
  OKL_REQUIRE(in && out, "NULL argument");
  fold_permute_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)in, (float*)out, k, OH, OW, inverse, total);
  OKL_LAUNCHED(ctx, "fold_permute");
  return OKL_OK;
}

static int make_pool_geom(PoolGeom& g, int N, int C, int H, int W, int pool, int stride, int layout) {
  OKL_REQUIRE(N >= 0 && C >= 1 && H >= 1 && W >= 1 && pool >= 1 && stride >= 1, "bad pooling shape");
  OKL_REQUIRE(H >= pool && W >= pool, "pool window larger than input");
  OKL_REQUIRE(layout == OKL_NCX || layout == OKL_NXC, "bad layout");
  g.N = N; g.C = C; g.H = H; g.W = W; g.ps = pool; g.st = stride;
  g.OH = (H - pool) / stride + 1; g.OW = (W - pool) / stride + 1;       // okl.py:1736-1737
  g.cl = layout == OKL_NXC;
  if (!g.cl) {
    g.xs[0] = (long long)C * H * W; g.xs[1] = (long long)H * W; g.xs[2] = W; g.xs[3] = 1;
    g.ys[0] = (long long)C * g.OH * g.OW; g.ys[1] = (long long)g.OH * g.OW; g.ys[2] = g.OW; g.ys[3] = 1;
  } else {
    g.xs[0] = (long long)C * H * W; g.xs[1] = 1; g.xs[2] = (long long)W * C; g.xs[3] = C;
    g.ys[0] = (long long)C * g.OH * g.OW; g.ys[1] = 1; g.ys[2] = (long long)g.OW * C; g.ys[3] = C;
  }
  return OKL_OK;
}

extern "C" int okl_pool_fwd(okl_ctx* ctx, int kind, const void* x, void* y, int N, int C, int H, int W, int pool, int stride, int layout) {
  OKL_REQUIRE(ctx && x && y, "NULL argument");
  OKL_REQUIRE(kind == OKL_POOL_MAX || kind == OKL_POOL_AVG, "bad pool kind");
  PoolGeom g;
  int s = make_pool_geom(g, N, C, H, W, pool, stride, layout);
  if (s != OKL_OK) return s;
  size_t total = (size_t)N * C * g.OH * g.OW;
  if (total == 0) return OKL_OK;
  if (kind == OKL_POOL_MAX && layout == OKL_NCX && pool == 2 && stride == 2 && (H & 1) == 0 && (W & 1) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 7) == 0) {
    maxpool2_fwd_kernel<<<okl_grid_1d(ctx, total, 256, 16), 256, 0, ctx->stream>>>((const float*)x, (float*)y, total, g.OW, W);
    OKL_LAUNCHED(ctx, "maxpool2_fwd");
    return OKL_OK;
  }
  if (kind == OKL_POOL_MAX && layout == OKL_NXC && pool == 2 && stride == 2 && (H & 1) == 0 && (W & 1) == 0 && (C & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    maxpool2_nxc_fwd_kernel<<<okl_grid_1d(ctx, total / 4, 256, 16), 256, 0, ctx->stream>>>((const float*)x, (float*)y, total / 4, C / 4, g.OW,
                                                                                      g.OH, W, H);
    OKL_LAUNCHED(ctx, "maxpool2_nxc_fwd");
    return OKL_OK;
  }
  pool_fwd_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)x, (float*)y, g, kind);
  OKL_LAUNCHED(ctx, "pool_fwd");
  return OKL_OK;
}

extern "C" int okl_pool_bwd(okl_ctx* ctx, int kind, const void* x, const void* gr, void* dx, int N, int C, int H, int W, int pool,
                            int stride, int layout, int relu_mask) {
  OKL_REQUIRE(ctx && x && gr && dx, "NULL argument");
  OKL_REQUIRE(kind == OKL_POOL_MAX || kind == OKL_POOL_AVG, "bad pool kind");
  PoolGeom g;
  int s = make_pool_geom(g, N, C, H, W, pool, stride, layout);
  if (s != OKL_OK) return s;
  size_t total = (size_t)N * C * H * W;
  if (total == 0) return OKL_OK;
  if (kind == OKL_POOL_MAX && layout == OKL_NCX && pool == 2 && stride == 2 && (H & 1) == 0 && (W & 1) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(dx) & 7) == 0) {
    size_t nwin = (size_t)N * C * g.OH * g.OW;
    maxpool2_bwd_kernel<<<okl_grid_1d(ctx, nwin, 256, 16), 256, 0, ctx->stream>>>((const float*)x, (const float*)gr, (float*)dx, nwin, g.OW,
                                                                              W, relu_mask);
    OKL_LAUNCHED(ctx, "maxpool2_bwd");
    return OKL_OK;
  }
  if (kind == OKL_POOL_MAX && layout == OKL_NXC && pool == 2 && stride == 2 && (H & 1) == 0 && (W & 1) == 0 && (C & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(gr) & 15) == 0) {
    size_t nwin4 = (size_t)N * g.OH * g.OW * (C / 4);
    maxpool2_nxc_bwd_kernel<<<okl_grid_1d(ctx, nwin4, 256, 16), 256, 0, ctx->stream>>>((const float*)x, (const float*)gr, (float*)dx, nwin4,
                                                                                  C / 4, g.OW, g.OH, W, H, relu_mask);
    OKL_LAUNCHED(ctx, "maxpool2_nxc_bwd");
    return OKL_OK;
  }
  pool_bwd_kernel<<<okl_grid_1d(ctx, total, 256), 256, 0, ctx->stream>>>((const float*)x, (const float*)gr, (float*)dx, g, kind, relu_mask);
  OKL_LAUNCHED(ctx, "pool_bwd");
  return OKL_OK;
}

extern "C" int okl_softmax_fwd(okl_ctx* ctx, const void* x, void* p, int rows, int cols) {
  OKL_REQUIRE(ctx && (rows == 0 || (x && p)), "NULL argument");
  OKL_REQUIRE(rows >= 0 && cols >= 1, "bad softmax shape");
  if (rows == 0) return OKL_OK;
  softmax_fwd_kernel<<<okl_grid_1d(ctx, (size_t)rows * 32, 256), 256, 0, ctx->stream>>>((const float*)x, (float*)p, rows, cols);
  OKL_LAUNCHED(ctx, "softmax_fwd");
  return OKL_OK;
}

extern "C" int okl_softmax_bwd(okl_ctx* ctx, const void* p, const void* g, void* dx, int rows, int cols) {
  OKL_REQUIRE(ctx && (rows == 0 || (p && g && dx)), "NULL argument");
  OKL_REQUIRE(rows >= 0 && cols >= 1, "bad softmax shape");
  if (rows == 0) return OKL_OK;
  softmax_bwd_kernel<<<okl_grid_1d(ctx, (size_t)rows * 32, 256), 256, 0, ctx->stream>>>((const float*)p, (const float*)g, (float*)dx,
                                                                                    rows, cols);
  OKL_LAUNCHED(ctx, "softmax_bwd");
  return OKL_OK;
}

extern "C" int okl_ce_fwd_bwd(okl_ctx* ctx, const void* p, const void* t, void* loss, void* grad, int rows, int cols, float denom,
                              void* loss_accum, void* softmax_dx) {
  OKL_REQUIRE(ctx && p && t && loss && grad, "NULL argument");
  OKL_REQUIRE(rows >= 1 && cols >= 1 && denom > 0.f, "bad cross-entropy shape");
  if (cols <= 16 && rows <= 16384) {
    int threads = rows >= 1024 ? 1024 : ((rows + 31) / 32) * 32;
    ce_fused_rowthread_kernel<16><<<1, threads, 0, ctx->stream>>>((const float*)p, (const int*)t, (float*)loss, (float*)grad,
                                                                (float*)softmax_dx, (float*)loss_accum, rows, cols, 1.f / denom);
    OKL_LAUNCHED(ctx, "ce_fused_rowthread");
    return OKL_OK;
  }
  if ((long long)rows * cols <= 65536) {
    ce_fused_small_kernel<<<1, 1024, 0, ctx->stream>>>((const float*)p, (const int*)t, (float*)loss, (float*)grad, (float*)softmax_dx,
                                                      (float*)loss_accum, rows, cols, 1.f / denom);
    OKL_LAUNCHED(ctx, "ce_fused_small");
    return OKL_OK;
  }
  ce_kernel<<<okl_grid_1d(ctx, (size_t)rows * cols, 256), 256, 0, ctx->stream>>>((const float*)p, (const int*)t, (float*)loss,
                                                                             (float*)grad, rows, cols, 1.f / denom);
  OKL_LAUNCHED(ctx, "ce_fwd_bwd");
  if (loss_accum) {
    add_scalar_kernel<<<1, 1, 0, ctx->stream>>>((float*)loss_accum, (const float*)loss);
    OKL_LAUNCHED(ctx, "add_scalar");
  }
  if (softmax_dx) return okl_softmax_bwd(ctx, p, grad, softmax_dx, rows, cols);
  return OKL_OK;
}

extern "C" int okl_mse_fwd_bwd(okl_ctx* ctx, const void* o, const void* t, void* loss, void* grad, size_t n, float denom,
                               void* loss_accum, int relu_outputs) {
  OKL_REQUIRE(ctx && o && t && loss && grad, "NULL argument");
  OKL_REQUIRE(n >= 1 && denom > 0.f, "bad mse shape");
  const bool vec = (n & 3) == 0 && aligned16(o) && aligned16(t) && aligned16(grad);
  int grid = vec ? okl_grid_1d(ctx, n / 4, 256, 2) : okl_grid_1d(ctx, n, 256, 4);
  if (vec && grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  int s = okl_ensure_workspace(ctx, (size_t)grid * sizeof(float));
  if (s != OKL_OK) return s;
  float* partial = (float*)ctx->workspace;
  if (vec) {
    if (relu_outputs) mse_kernel4<true><<<grid, 256, 0, ctx->stream>>>((const float4*)o, (const float4*)t, partial, (float4*)grad, n / 4, 2.f / denom);
    else mse_kernel4<false><<<grid, 256, 0, ctx->stream>>>((const float4*)o, (const float4*)t, partial, (float4*)grad, n / 4, 2.f / denom);
  } else {
    if (relu_outputs) mse_kernel<true><<<grid, 256, 0, ctx->stream>>>((const float*)o, (const float*)t, partial, (float*)grad, n, 2.f / denom);
    else mse_kernel<false><<<grid, 256, 0, ctx->stream>>>((const float*)o, (const float*)t, partial, (float*)grad, n, 2.f / denom);
  }
  OKL_LAUNCHED(ctx, "mse_fwd_bwd");
  final_sum_kernel<<<1, 256, 0, ctx->stream>>>((const float*)partial, grid, (float*)loss, 1.f / (float)n, (float*)loss_accum);
  OKL_LAUNCHED(ctx, "final_sum");
  return OKL_OK;
}

extern "C" int okl_mse_fwd_bwd_nxc(okl_ctx* ctx, const void* o_nxc, const void* t_ncx, void* loss, void* grad_nxc, int N, int Cc, long long S,
                                   float denom, void* loss_accum, int relu_outputs) {
  OKL_REQUIRE(ctx && o_nxc && t_ncx && loss && grad_nxc, "NULL argument");
  OKL_REQUIRE(N >= 1 && Cc >= 1 && S >= 1 && denom > 0.f, "bad mse shape");
  dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((S + 32 * MSE_NXC_TILES - 1) / (32 * MSE_NXC_TILES)), (unsigned)N);
  OKL_REQUIRE(grid.y <= 65535 && N <= 65535, "okl_mse_fwd_bwd_nxc: extent too large");
  const size_t nblk = (size_t)grid.x * grid.y * grid.z;
  OKL_REQUIRE(nblk < (1u << 30), "okl_mse_fwd_bwd_nxc: too many tiles");
  int s = okl_ensure_workspace(ctx, nblk * sizeof(float));
  if (s != OKL_OK) return s;
  float* partial = (float*)ctx->workspace;
  if (relu_outputs) mse_nxc_kernel<true><<<grid, dim3(32, 8), 0, ctx->stream>>>((const float*)o_nxc, (const float*)t_ncx, partial, (float*)grad_nxc, Cc, S, 2.f / denom);
  else mse_nxc_kernel<false><<<grid, dim3(32, 8), 0, ctx->stream>>>((const float*)o_nxc, (const float*)t_ncx, partial, (float*)grad_nxc, Cc, S, 2.f / denom);
  OKL_LAUNCHED(ctx, "mse_fwd_bwd_nxc");
  final_sum_kernel<<<1, 1024, 0, ctx->stream>>>((const float*)partial, (int)nblk, (float*)loss, 1.f / ((float)N * (float)Cc * (float)S), (float*)loss_accum);
  OKL_LAUNCHED(ctx, "final_sum");
  return OKL_OK;
}

int okl_colsum(okl_ctx* ctx, const float* G, float* out, long long rows, int cols) {
  if (cols <= 0) return OKL_OK;
  int col_blocks = (cols + 31) / 32;
  if (rows <= 4096 || col_blocks >= ctx->sm_count) {
    okl_launch(ctx, colsum_single, dim3(col_blocks), dim3(32, 32), 0, G, out, rows, cols);
    OKL_LAUNCHED(ctx, "colsum_single");
    return OKL_OK;
  }
  long long want = (2LL * ctx->sm_count + col_blocks - 1) / col_blocks;
  long long chunks = rows / 64 < 1 ? 1 : rows / 64;
  if (chunks > want) chunks = want;
  if (chunks > 65535) chunks = 65535;
  long long rpc = (rows + chunks - 1) / chunks;
  if (rpc < 1) rpc = 1;
  chunks = rows > 0 ? (rows + rpc - 1) / rpc : 1;
  int s = okl_ensure_workspace(ctx, (size_t)chunks * cols * sizeof(float));
  if (s != OKL_OK) return s;
  colsum_stage1<<<dim3(col_blocks, (unsigned)chunks), dim3(32, 8), 0, ctx->stream>>>(G, (float*)ctx->workspace, rows, cols, rpc);
  OKL_LAUNCHED(ctx, "colsum_stage1");
  colsum_stage2<<<(cols + 255) / 256, 256, 0, ctx->stream>>>((const float*)ctx->workspace, out, (int)chunks, cols);
  OKL_LAUNCHED(ctx, "colsum_stage2");
  return OKL_OK;
}

int okl_channel_sum(okl_ctx* ctx, const float* g, float* out, int N, int C, long long S, int layout) {
  if (layout == OKL_NXC) return okl_colsum(ctx, g, out, (long long)N * S, C);
  int want = (2 * ctx->sm_count + C - 1) / C;
  int chunks = N < want ? N : want;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  int npc = (N + chunks - 1) / chunks;
  if (npc < 1) npc = 1;
  chunks = N > 0 ? (N + npc - 1) / npc : 1;
  int s = okl_ensure_workspace(ctx, (size_t)chunks * C * sizeof(float));
  if (s != OKL_OK) return s;
  chansum_stage1<<<dim3(C, chunks), 256, 0, ctx->stream>>>(g, (float*)ctx->workspace, N, C, S, npc);
  OKL_LAUNCHED(ctx, "chansum_stage1");
  colsum_stage2<<<(C + 255) / 256, 256, 0, ctx->stream>>>((const float*)ctx->workspace, out, chunks, C);
  OKL_LAUNCHED(ctx, "colsum_stage2");
  return OKL_OK;
}

extern "C" int okl_accum_update(okl_ctx* ctx, void* w, void* gacc, const void* g, size_t n, float lr, const void* lr_dev) {
  OKL_REQUIRE(ctx && (n == 0 || (w && gacc && g)), "NULL argument");
  if (n == 0) return OKL_OK;
  bool vec = aligned16(w) && aligned16(gacc) && aligned16(g);
  int grid = okl_grid_1d(ctx, vec ? (n + 3) / 4 : n, 256);
  if (vec) okl_launch(ctx, accum_update_kernel<true>, dim3(grid), dim3(256), 0, (float*)w, (float*)gacc, (const float*)g, n, lr, (const float*)lr_dev);
  else okl_launch(ctx, accum_update_kernel<false>, dim3(grid), dim3(256), 0, (float*)w, (float*)gacc, (const float*)g, n, lr, (const float*)lr_dev);
  OKL_LAUNCHED(ctx, "accum_update");
  return OKL_OK;
}

extern "C" int okl_accum_update_multi16(okl_ctx* ctx, int count, void* const* w, void* const* gacc, const void* const* g, const size_t* n,
                                        void* const* w16, const size_t* n16, float lr, const void* lr_dev);

extern "C" int okl_accum_update_multi(okl_ctx* ctx, int count, void* const* w, void* const* gacc, const void* const* g, const size_t* n,
                                      float lr, const void* lr_dev) {
  return okl_accum_update_multi16(ctx, count, w, gacc, g, n, nullptr, nullptr, lr, lr_dev);
}

// the same, and bucket i's first n16[i] parameters (a multiple of 4) are also written as bf16 to w16[i] (NULL entries: no shadow)
extern "C" int okl_accum_update_multi16(okl_ctx* ctx, int count, void* const* w, void* const* gacc, const void* const* g, const size_t* n,
                                        void* const* w16, const size_t* n16, float lr, const void* lr_dev) {
  OKL_REQUIRE(ctx && (count == 0 || (w && gacc && g && n)), "NULL argument");
  OKL_REQUIRE((w16 == nullptr) == (n16 == nullptr), "w16 and n16 go together");
  OKL_REQUIRE(count >= 0, "bad count");
  for (int base = 0; base < count; base += 16) {
    MultiUpd mu;
    int c = count - base < 16 ? count - base : 16;
    size_t nmax = 0;
    for (int i = 0; i < 16; i++) {
      int j = base + (i < c ? i : 0);
      mu.w[i] = (float*)w[j]; mu.gacc[i] = (float*)gacc[j]; mu.g[i] = (const float*)g[j]; mu.n[i] = i < c ? n[j] : 0;
      mu.w16[i] = (w16 && i < c) ? (__nv_bfloat16*)w16[j] : nullptr;
      mu.n16[i] = mu.w16[i] ? n16[j] : 0;
      OKL_REQUIRE(mu.n16[i] % 4 == 0 && mu.n16[i] <= mu.n[i] && (reinterpret_cast<uintptr_t>(mu.w16[i]) & 7) == 0,
                  "okl_accum_update_multi16: shadow extent must be a multiple of 4 inside the bucket, 8-byte aligned");
      OKL_REQUIRE(aligned16(mu.w[i]) && aligned16(mu.gacc[i]) && aligned16(mu.g[i]), "okl_accum_update_multi: buckets must be 16-byte aligned");
      if (mu.n[i] > nmax) nmax = mu.n[i];
    }
    if (nmax == 0) continue;
    dim3 grid(okl_grid_1d(ctx, (nmax + 3) / 4, 256, 4), c);
    accum_update_multi_kernel<<<grid, 256, 0, ctx->stream>>>(mu, lr, (const float*)lr_dev);
    OKL_LAUNCHED(ctx, "accum_update_multi");
  }
  return OKL_OK;
}

extern "C" int okl_sgd_update(okl_ctx* ctx, void* w, void* v, const void* g, size_t n, float lr, float momentum) {
  OKL_REQUIRE(ctx && (n == 0 || (w && v && g)), "NULL argument");
  if (n == 0) return OKL_OK;
  sgd_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((float*)w, (float*)v, (const float*)g, n, lr, momentum);
  OKL_LAUNCHED(ctx, "sgd_update");
  return OKL_OK;
}

extern "C" int okl_adam_update(okl_ctx* ctx, void* w, void* m, void* v, const void* g, size_t n, float lr, float beta1, float beta2,
                               float eps, int t) {
  OKL_REQUIRE(ctx && (n == 0 || (w && m && v && g)), "NULL argument");
  OKL_REQUIRE(t >= 1, "adam step t must be >= 1");
  if (n == 0) return OKL_OK;
  float c1 = (float)(1.0 - pow((double)beta1, (double)t)), c2 = (float)(1.0 - pow((double)beta2, (double)t));
  adam_kernel<<<okl_grid_1d(ctx, n, 256), 256, 0, ctx->stream>>>((float*)w, (float*)m, (float*)v, (const float*)g, n, lr, beta1, beta2,
                                                              eps, c1, c2);
  OKL_LAUNCHED(ctx, "adam_update");
  return OKL_OK;
}
