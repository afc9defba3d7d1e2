// Placeholder until the tcgen05 implicit-GEMM convolution lands: the dispatcher stays on the FFMA conv kernels.
#include "common.cuh"
bool okl_tc_conv_supported(okl_ctx*, const okl_conv_desc*) { return false; }
int okl_tc_conv_fprop(okl_ctx*, const okl_conv_desc*, const float*, const float*, const float*, float*) { return OKL_ERR_UNSUPPORTED; }
int okl_tc_conv_dgrad(okl_ctx*, const okl_conv_desc*, const float*, const float*, float*) { return OKL_ERR_UNSUPPORTED; }
int okl_tc_conv_wgrad(okl_ctx*, const okl_conv_desc*, const float*, const float*, float*) { return OKL_ERR_UNSUPPORTED; }
