// Placeholder until the tcgen05 kernels land: reports "not supported" so the dispatcher stays on the FFMA kernels.
#include "common.cuh"
bool okl_tc_dense_supported(okl_ctx*, int, int, int) { return false; }
int okl_tc_dense_fwd(okl_ctx*, const float*, const float*, const float*, float*, int, int, int, int) { return OKL_ERR_UNSUPPORTED; }
int okl_tc_dense_dw(okl_ctx*, const float*, const float*, float*, int, int, int) { return OKL_ERR_UNSUPPORTED; }
int okl_tc_dense_dx(okl_ctx*, const float*, const float*, float*, int, int, int) { return OKL_ERR_UNSUPPORTED; }
bool okl_tc_conv_supported(okl_ctx*, const okl_conv_desc*) { return false; }
int okl_tc_conv_fprop(okl_ctx*, const okl_conv_desc*, const float*, const float*, const float*, float*) { return OKL_ERR_UNSUPPORTED; }
int okl_tc_conv_dgrad(okl_ctx*, const okl_conv_desc*, const float*, const float*, float*) { return OKL_ERR_UNSUPPORTED; }
int okl_tc_conv_wgrad(okl_ctx*, const okl_conv_desc*, const float*, const float*, float*) { return OKL_ERR_UNSUPPORTED; }
