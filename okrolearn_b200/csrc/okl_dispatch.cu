// Public Dense / Conv entry points: argument validation and the choice between the exact-fp32 FFMA kernels
// (okl_ffma.cu) and the tcgen05 tensor-core kernels (okl_tc_*.cu) according to the context's precision mode.
#include "common.cuh"

extern "C" int okl_dense_fwd(okl_ctx* ctx, const void* X, const void* W, const void* b, void* Y, int M, int K, int N, int act) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_REQUIRE(M >= 0 && K >= 1 && N >= 1, "okl_dense_fwd: bad shape M=%d K=%d N=%d", M, K, N);
  OKL_REQUIRE(act >= OKL_ACT_NONE && act <= OKL_ACT_SOFTMAX, "bad activation %d", act);
  if (M == 0) return OKL_OK;
  OKL_REQUIRE(X && W && Y, "okl_dense_fwd: NULL tensor");
  if (okl_dense_small_ok(M, K, N))
    return okl_dense_small_fwd(ctx, (const float*)X, (const float*)W, (const float*)b, (float*)Y, M, K, N, act);
  OKL_REQUIRE(act != OKL_ACT_SOFTMAX, "okl_dense_fwd: the fused softmax epilogue needs an output width <= 16");
  if (ctx->precision != OKL_FP32 && okl_tc_dense_supported(ctx, M, K, N))
    return okl_tc_dense_fwd(ctx, (const float*)X, (const float*)W, (const float*)b, (float*)Y, M, K, N, act);
  return okl_ffma_dense_fwd(ctx, (const float*)X, (const float*)W, (const float*)b, (float*)Y, M, K, N, act);
}

extern "C" int okl_dense_bwd(okl_ctx* ctx, const void* X, const void* W, const void* G, void* dX, void* dW, void* db,
                             int M, int K, int N, const void* dx_relu_mask) {
  OKL_REQUIRE(ctx, "ctx is NULL");
  OKL_REQUIRE(M >= 0 && K >= 1 && N >= 1, "okl_dense_bwd: bad shape M=%d K=%d N=%d", M, K, N);
  OKL_REQUIRE(X && W && G, "okl_dense_bwd: NULL tensor");
  if (M > 0 && okl_dense_small_ok(M, K, N))
    return okl_dense_small_bwd(ctx, (const float*)X, (const float*)W, (const float*)G, (float*)dX, (float*)dW, (float*)db,
                               (const float*)dx_relu_mask, M, K, N);
  int s;
  const bool tc = ctx->precision != OKL_FP32 && M > 0 && okl_tc_dense_supported(ctx, M, K, N);
  if (dW) {
    if (M == 0) s = okl_memset(ctx, dW, 0, (size_t)K * N * sizeof(float));
    else if (tc) s = okl_tc_dense_dw(ctx, (const float*)X, (const float*)G, (float*)dW, M, K, N);
    else s = okl_ffma_dense_dw(ctx, (const float*)X, (const float*)G, (float*)dW, M, K, N);
    if (s != OKL_OK) return s;
  }
  if (db) {
    s = okl_colsum(ctx, (const float*)G, (float*)db, M, N);
    if (s != OKL_OK) return s;
  }
  if (dX && M > 0) {
    if (tc) s = okl_tc_dense_dx(ctx, (const float*)G, (const float*)W, (float*)dX, M, K, N, (const float*)dx_relu_mask);
    else s = okl_ffma_dense_dx(ctx, (const float*)G, (const float*)W, (float*)dX, M, K, N, (const float*)dx_relu_mask);
    if (s != OKL_OK) return s;
  }
  return OKL_OK;
}

static int conv_sizes(const okl_conv_desc* d, long long* in_sp, long long* out_sp) {
  int oe[3] = {1, 1, 1};
  int s = okl_conv_out_extent(d, oe);
  if (s != OKL_OK) return s;
  long long a = 1, b = 1;
  for (int i = 0; i < d->nd; i++) { a *= d->in[i]; b *= oe[i]; }
  *in_sp = a; *out_sp = b;
  return OKL_OK;
}

extern "C" int okl_conv_fprop(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* W, const void* b, void* y) {
  OKL_REQUIRE(ctx && d, "NULL argument");
  long long si, so;
  int s = conv_sizes(d, &si, &so);
  if (s != OKL_OK) return s;
  OKL_REQUIRE(d->act >= OKL_ACT_NONE && d->act <= OKL_ACT_TANH, "bad activation %d", d->act);
  if (d->N == 0) return OKL_OK;
  OKL_REQUIRE(x && W && y, "okl_conv_fprop: NULL tensor");
  // small C, many taps: im2col tile built in shared memory in the UMMA layout, tcgen05 (video layer: 0.43 ms vs 0.59 ms for the
  // direct FFMA kernel); OKL_TC_SMALLC_FPROP=0 selects the FFMA kernel
  if (ctx->precision != OKL_FP32 && okl_tc_conv_smallc_supported(ctx, d)) {
    const char* e = getenv("OKL_TC_SMALLC_FPROP");
    if (!e || atoi(e)) return okl_tc_conv_smallc_fprop(ctx, d, (const float*)x, (const float*)W, (const float*)b, (float*)y);
  }
  if (okl_conv_small_fprop_ok(d)) return okl_conv_small_fprop(ctx, d, (const float*)x, (const float*)W, (const float*)b, (float*)y);
  if (ctx->precision != OKL_FP32 && okl_tc_conv_supported(ctx, d))
    return okl_tc_conv_fprop(ctx, d, (const float*)x, (const float*)W, (const float*)b, (float*)y);
  return okl_ffma_conv_fprop(ctx, d, (const float*)x, (const float*)W, (const float*)b, (float*)y);
}

extern "C" int okl_conv_dgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* g, const void* W, void* dx, const void* dx_relu_mask) {
  OKL_REQUIRE(ctx && d, "NULL argument");
  long long si, so;
  int s = conv_sizes(d, &si, &so);
  if (s != OKL_OK) return s;
  if (d->N == 0) return OKL_OK;
  OKL_REQUIRE(g && W && dx, "okl_conv_dgrad: NULL tensor");
  if (ctx->precision != OKL_FP32 && okl_tc_conv_supported(ctx, d))
    return okl_tc_conv_dgrad(ctx, d, (const float*)g, (const float*)W, (float*)dx, (const float*)dx_relu_mask);
  return okl_ffma_conv_dgrad(ctx, d, (const float*)g, (const float*)W, (float*)dx, (const float*)dx_relu_mask);
}

extern "C" int okl_conv_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* g, void* dW, void* db) {
  OKL_REQUIRE(ctx && d, "NULL argument");
  long long si, so;
  int s = conv_sizes(d, &si, &so);
  if (s != OKL_OK) return s;
  long long taps = 1;
  for (int i = 0; i < d->nd; i++) taps *= d->k[i];
  if (d->N == 0) {
    if (dW) { s = okl_memset(ctx, dW, 0, (size_t)d->O * d->C * taps * sizeof(float)); if (s != OKL_OK) return s; }
    if (db) { s = okl_memset(ctx, db, 0, (size_t)d->O * sizeof(float)); if (s != OKL_OK) return s; }
    return OKL_OK;
  }
  OKL_REQUIRE(x && g, "okl_conv_wgrad: NULL tensor");
  if (ctx->precision != OKL_FP32 && okl_tc_conv_smallc_supported(ctx, d))
    return okl_tc_conv_smallc_wgrad(ctx, d, (const float*)x, (const float*)g, (float*)dW, (float*)db);
  if (okl_conv_small_wgrad_ok(d)) return okl_conv_small_wgrad(ctx, d, (const float*)x, (const float*)g, (float*)dW, (float*)db);
  if (dW) {
    if (ctx->precision != OKL_FP32 && okl_tc_conv_supported(ctx, d))
      s = okl_tc_conv_wgrad(ctx, d, (const float*)x, (const float*)g, (float*)dW);
    else
      s = okl_ffma_conv_wgrad(ctx, d, (const float*)x, (const float*)g, (float*)dW);
    if (s != OKL_OK) return s;
  }
  if (db) {
    s = okl_channel_sum(ctx, (const float*)g, (float*)db, d->N, d->O, so, d->y_layout);
    if (s != OKL_OK) return s;
  }
  return OKL_OK;
}
