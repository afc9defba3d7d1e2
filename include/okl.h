/*
 * okl.h - C ABI of libokl_b200.so: the B200-native (sm_100a) implementation of okrolearn's
 * Conv/Dense train-step hot path (BASELINE.json north_star).
 *
 * The reference (okrolearn 0.2.7, pure Python + NumPy) has no FFI; its boundary is the duck-typed
 * layer protocol `forward(Tensor)->Tensor`, `backward(Tensor, lr)->Tensor` (okrolearn.py:2887-2888,
 * 2910-2911).  The Python package okrolearn_b200 re-implements that protocol and calls ONLY the
 * functions below through ctypes.  Each entry point cites the reference code it replaces
 * ("okl.py" = src/okrolearn/okrolearn.py, "tensor.py", "optimizers.py" under /root/reference).
 *
 * Conventions
 *   - every function returns an okl_status (0 = OK, < 0 = error) and never throws;
 *     okl_last_error() returns a NUL-terminated message owned by the library (thread-local).
 *   - device pointers are plain `void*` obtained from okl_malloc (sub-ranges are allowed);
 *     all device tensors are dense fp32 unless stated (bf16 shadows are internal).
 *   - host pointers are borrowed for the duration of the call only.
 *   - all work is enqueued on the context's compute stream; only okl_d2h, okl_ctx_sync,
 *     okl_timer_stop and okl_read_scalar block the host.
 *   - one okl_ctx <-> one CUDA device <-> one compute stream + one comm stream; a ctx is used from
 *     one host thread at a time.  Multi-GPU = one process per GPU (okl_comm_*).
 *   - tensors use the reference's layouts: Dense X (M,K) row-major, W (K,N) row-major, b (N);
 *     conv x (N,C,*spatial), filters (O,C,*k), biases (O); `layout` OKL_NCX is that channels-first
 *     layout, OKL_NXC is the channels-last layout the tensor-core kernels use internally.
 */
#ifndef OKL_H
#define OKL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct okl_ctx okl_ctx;
typedef struct okl_graph okl_graph;

typedef enum {
  OKL_OK = 0,
  OKL_ERR_INVALID = -1,      /* bad argument / shape (Python raises ValueError) */
  OKL_ERR_CUDA = -2,         /* CUDA runtime / driver failure (Python raises RuntimeError) */
  OKL_ERR_NCCL = -3,
  OKL_ERR_OOM = -4,
  OKL_ERR_UNSUPPORTED = -5
} okl_status;

/* arithmetic mode of the GEMM-shaped kernels (north_star: 1e-5 rel in FP32, 2e-2 in TF32/BF16) */
typedef enum { OKL_FP32 = 0, OKL_TF32 = 1, OKL_BF16 = 2 } okl_precision;
/* OKL_ACT_SOFTMAX: row softmax (okl.py:364-373) as the epilogue of okl_dense_fwd, for output widths <= 16 only */
typedef enum { OKL_ACT_NONE = 0, OKL_ACT_RELU = 1, OKL_ACT_SIGMOID = 2, OKL_ACT_TANH = 3, OKL_ACT_SOFTMAX = 4,
               /* okl_act2_fwd / okl_act2_bwd only: */
               OKL_ACT_ELU = 10, OKL_ACT_SELU = 11, OKL_ACT_LEAKY = 12, OKL_ACT_SWISH = 13, OKL_ACT_GELU = 14, OKL_ACT_SOFTSIGN = 15,
               OKL_ACT_HARDTANH = 16,
               /* okl_act2_bwd only: dx = g * (1 - y^2) with the tanh OUTPUT y passed as x (RNNLayer, okl.py:1667) */
               OKL_ACT_TANH_OUT = 17 } okl_act;
typedef enum { OKL_F32 = 0, OKL_F64 = 1, OKL_I64 = 2, OKL_I32 = 3, OKL_U8 = 4 } okl_dtype;
typedef enum { OKL_NCX = 0, OKL_NXC = 1 } okl_layout;
typedef enum { OKL_POOL_MAX = 0, OKL_POOL_AVG = 1 } okl_pool_kind;

/* ------------------------------------------------------------------ context / errors */
const char* okl_version(void);
const char* okl_last_error(void);
int okl_ctx_create(int device, int precision, okl_ctx** out);
int okl_ctx_destroy(okl_ctx* ctx);
int okl_ctx_sync(okl_ctx* ctx);
int okl_ctx_set_precision(okl_ctx* ctx, int precision);
int okl_ctx_get_precision(okl_ctx* ctx, int* precision);
/* sm_count, total HBM bytes, compute capability major*10+minor */
int okl_device_info(okl_ctx* ctx, int* sm_count, size_t* total_mem, int* cc);
/* number of kernels this library has launched on ctx (graph replays count their captured nodes) */
int okl_launch_count(okl_ctx* ctx, unsigned long long* out);

/* ------------------------------------------------------------------ memory (replaces Tensor.data ndarray storage, tensor.py:6-10) */
int okl_malloc(okl_ctx* ctx, size_t bytes, void** out);      /* caching pool; safe during graph capture on a hit */
int okl_free(okl_ctx* ctx, void* p);
int okl_pool_reserve(okl_ctx* ctx, size_t bytes);            /* pre-grow the pool with one block of `bytes` */
int okl_memset(okl_ctx* ctx, void* p, int byte, size_t bytes);
int okl_pinned_alloc(size_t bytes, void** out);
int okl_pinned_free(void* p);
/* host -> device fp32, converting from src_dtype on the device; n = element count */
int okl_h2d(okl_ctx* ctx, void* dst_f32, const void* host, size_t n, int src_dtype);
/* same, host buffer must be pinned (okl_pinned_alloc), fully asynchronous, f32 / i32 only (raw copy) */
int okl_h2d_async_raw(okl_ctx* ctx, void* dst, const void* pinned_host, size_t bytes);
/* device fp32 -> host dst_dtype (F32 or F64); blocks until the data is in `host` */
int okl_d2h(okl_ctx* ctx, void* host, const void* src_f32, size_t n, int dst_dtype);
int okl_d2h_async_raw(okl_ctx* ctx, void* pinned_host, const void* src, size_t bytes);
int okl_d2d(okl_ctx* ctx, void* dst, const void* src, size_t bytes);
/* host int64/int32 class targets -> device int32 */
int okl_h2d_i32(okl_ctx* ctx, void* dst_i32, const void* host, size_t n, int src_dtype);
/* channels-first (N,C,S) <-> channels-last (N,S,C) permutation, S = prod(spatial) */
int okl_permute(okl_ctx* ctx, void* dst, const void* src, int N, int C, long long S, int src_layout);
/* dst[r,:] = src[idx[r],:]  (epoch shuffle okl.py:2957-2960 as a device gather); idx int32 on device */
int okl_gather_rows(okl_ctx* ctx, void* dst, const void* src, const void* idx_i32, long long rows, long long row_elems);
/* the same for one training batch: rows idx[0..rows) of the inputs (fp32) and of the targets (32-bit words) in ONE launch
 * (batch slicing okl.py:2966-2967); either source may be device memory or page-locked host memory */
int okl_gather_batch(okl_ctx* ctx, void* x_dst, const void* x_src, long long x_row_elems, void* t_dst, const void* t_src,
                     long long t_row_elems, const void* idx_i32, long long rows);

/* Batch pipeline of the train loop (okl.py:2963-2967 slices the next batch on the host between steps; here the gather of step
 * s+1's rows - from HBM, or over PCIe from page-locked host memory - runs on a copy stream WHILE step s computes).
 * okl_batch_prefetch: okl_gather_batch into input slot 0/1 on the copy stream, ordered after everything enqueued so far on the
 * compute stream (so the slot's previous consumer is finished); okl_batch_wait: the compute stream waits for that slot.
 * x_nxc_dst (optional): additionally a channels-last (N, S, C) copy of the gathered inputs, for a first layer that runs
 * channels-last on the tensor cores - its layout permutation then overlaps the previous step too; x_nxc_bf16 != 0: that copy
 * is written as bf16 (okl_to_bf16_nxc: permutation and conversion in one pass), the operand format of the BF16 convolutions.
 * Gather-free parts: a NULL x_dst / t_dst skips that gather - its consumer reads the dataset rows through the index itself (the
 * *_rows_i32 arguments of okl_to_bf16_nxc, okl_conv_first_*, okl_mse_fwd_bwd_nxc_bf16); idx_dst (optional) receives the batch's
 * index slice at a fixed per-slot address for those consumers (and for graph replay). */
/* Stream mode (dataset in page-locked host memory): a host copy of the NEXT okl_batch_prefetch's row indices (int32, valid until
 * that call returns).  With it the rows are gathered by host threads into a page-locked staging ring and cross PCIe as ONE DMA copy
 * per part, instead of a zero-copy gather kernel that competes with the step's kernels for SM resources. */
int okl_batch_host_index(okl_ctx* ctx, const void* idx_host_i32);
int okl_batch_prefetch(okl_ctx* ctx, int slot, void* x_dst, const void* x_src, long long x_row_elems, void* t_dst, const void* t_src,
                       long long t_row_elems, const void* idx_i32, long long rows, void* x_nxc_dst, int x_channels, int x_nxc_bf16,
                       void* idx_dst);
int okl_batch_wait(okl_ctx* ctx, int slot);
/* Per-step loss read-back (okl.py:2973 reads loss.data every batch) without stalling the device: _async copies the device scalar
 * to page-locked host memory in stream order (slot 0/1), _wait blocks the host until that copy has landed and returns it - called
 * for step s-1 after step s has been enqueued, so the device never idles. */
int okl_loss_fetch_async(okl_ctx* ctx, int slot, const void* dev_f32);
int okl_loss_fetch_wait(okl_ctx* ctx, int slot, float* out);

/* ------------------------------------------------------------------ Dense (okl.py:18-54; tensor.py:57-67)
 * fwd: Y[M,N] = act(X[M,K] . W[K,N] + b[N])                       okl.py:31 (+ fused activation epilogue)
 * bwd: dW[K,N] = X^T . G ; db[N] = sum_rows G ; dX[M,K] = G . W^T  okl.py:35-37   (dX may be NULL)            */
int okl_dense_fwd(okl_ctx* ctx, const void* X, const void* W, const void* b, void* Y, int M, int K, int N, int act);
int okl_dense_bwd(okl_ctx* ctx, const void* X, const void* W, const void* G, void* dX, void* dW, void* db,
                  int M, int K, int N, const void* dx_relu_mask);
/* dx_relu_mask (optional, shape of dX): dX is zeroed where mask <= 0 - the fused backward `g * (x > 0)` (okl.py:107) of a ReLU
 * that sits between the previous layer and this one, evaluated in the GEMM epilogue instead of a separate pass. */

/* BF16 mode with caller-kept shadows: X16 / W16 / G16 are bf16 copies of the fp32 tensors (okl_to_bf16, or written by a producing
 * kernel as Y16 / dX16), so the GEMMs read 2-byte operands without a per-call conversion; outputs stay fp32 (+ optional bf16). */
int okl_to_bf16(okl_ctx* ctx, const void* src_f32, void* dst_bf16, size_t n);
int okl_dense_fwd_bf16(okl_ctx* ctx, const void* X16, const void* W16, const void* b, void* Y, void* Y16, int M, int K, int N, int act);
int okl_dense_bwd_bf16(okl_ctx* ctx, const void* X16, const void* W16, const void* G16, const void* G, void* dX, void* dX16, void* dW,
                       void* db, int M, int K, int N, const void* dx_relu_mask);

/* The tcgen05 GEMM the Dense products run on in TF32/BF16 mode, exposed for tests and benchmarks (reference: np.dot in
 * Tensor.dot, tensor.py:58).  D[M,N] (row-major, ldd) = act(A . B + bias[N]); A, B, D, bias are fp32 device tensors.
 *   a_mn = 0: A stored [M][K] (lda);  a_mn = 1: A stored [K][M] (lda)   -  b_mn = 0: B stored [N][K];  b_mn = 1: B stored [K][N]
 *   kind 0 = TF32 operands, 1 = BF16 operands (converted on the device); accumulation is fp32 in tensor memory. */
int okl_tc_gemm(okl_ctx* ctx, int kind, int a_mn, int b_mn, int M, int N, int K, const void* A, long long lda, const void* B,
                long long ldb, void* D, long long ldd, const void* bias, int act);

/* ------------------------------------------------------------------ Conv1D/2D/3D (okl.py:669-764, 767-883, 936-1049)
 * nd = 1,2,3; unused trailing entries of in/k/s/p are ignored.  Cross-correlation, zero padding both sides,
 * output extent (in + 2p - k) / s + 1 (floor).  filters (O,C,*k) and biases (O) are always in the reference layout. */
typedef struct {
  int nd;
  int N, C, O;
  int in[3];
  int k[3];
  int s[3];
  int p[3];
  int act;          /* okl_act fused into fprop's epilogue (bias first) */
  int x_layout;     /* okl_layout of x / dx */
  int y_layout;     /* okl_layout of y / g  */
} okl_conv_desc;

int okl_conv_out_extent(const okl_conv_desc* d, int out[3]);
/* y = act(conv(x, W) + b)                      okl.py:691-715 / 790-813 / 948-985 */
int okl_conv_fprop(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* W, const void* b, void* y);
/* dx = conv_transpose(g, W)                    okl.py:736-739 / 867-868 / 1017-1020 (+ repairs R1,R3) */
int okl_conv_dgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* g, const void* W, void* dx, const void* dx_relu_mask);
/* dW = x (*) g ; db = sum_{n,pos} g            okl.py:729-735 / 865-866,870 / 1013-1016,991 (+ repairs R1,R2); db may be NULL */
int okl_conv_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* g, void* dW, void* db);

/* BF16 shadows for the implicit-GEMM convolution (channels-last, unit stride, C and O multiples of 64): activations / gradients are read
 * through their bf16 shadows, results are written as fp32 + bf16 */
int okl_conv_bf16_supported(okl_ctx* ctx, const okl_conv_desc* d);
int okl_conv_fprop_bf16(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* W, const void* b, void* y, void* y16);
int okl_conv_dgrad_bf16(okl_ctx* ctx, const okl_conv_desc* d, const void* g16, const void* W, void* dx, void* dx16, const void* dx_relu_mask);
int okl_conv_wgrad_bf16(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* g16, const void* g, void* dW, void* db);

/* Fused first block Conv2D(3x3, stride 1) -> ReLU -> MaxPool(2,2) for tiny input channel counts (C = 1 or 3): the un-pooled
 * activation is never written.  fwd writes the pooled maxima (N,O,PH,PW) and a 4-bit mask per pooled element (which of the four
 * window positions attain the maximum and are > 0 = the combined backward of okl.py:107 and okl.py:1757-1766 with the reference's
 * tie rule); wgrad consumes the POOLED gradient and that mask and yields dW (O,C,3,3) and db (O) (okl.py:865-866, 870). */
int okl_conv_relu_pool_supported(const okl_conv_desc* d);
int okl_conv_relu_pool_fwd(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* W, const void* b, void* pooled, void* mask_u8);
int okl_conv_relu_pool_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* pooled_g, const void* mask_u8, void* dW,
                             void* db);

/* ------------------------------------------------------------------ activations (okl.py:94-120, 500-522, 2070-2108) */
int okl_act_fwd(okl_ctx* ctx, int act, const void* x, void* y, size_t n);
/* relu: ref = x (g * (x>0), okl.py:107) [y works too]; sigmoid: ref = y (okl.py:514); tanh: ref = x (okl.py:2092) */
int okl_act_bwd(okl_ctx* ctx, int act, const void* ref, const void* g, void* dx, size_t n);
/* the reference's other elementwise activations (okl.py:57-352, 2111-2153): y = f(x), dx = g * f'(x) evaluated from the INPUT x.
 * alpha: ELU / SELU / LeakyReLU / PReLU slope; scale: SELU.  Swish, GELU (tanh form, derivative as written in okl.py:300) and the
 * clips the reference applies are reproduced literally. */
int okl_act2_fwd(okl_ctx* ctx, int act, const void* x, void* y, size_t n, float alpha, float scale);
int okl_act2_bwd(okl_ctx* ctx, int act, const void* x, const void* g, void* dx, size_t n, float alpha, float scale);
/* PReLU (okl.py:250-251): out_scalar[0] = sum(g * x * [x <= 0]), fixed summation order */
int okl_prelu_alpha_grad(okl_ctx* ctx, const void* x, const void* g, size_t n, void* out_scalar);
/* row-wise classifiers next to softmax: kind 0 = Softmin forward (okl.py:410-416: exp(-(x + max)) normalised, nan_to_num),
 * 1 = LogSoftmax forward (455-468), 2 = LogSoftmax backward a = log-softmax output, b = g: g_j - sum_k exp(a_k) g_k (470-488).
 * Softmin's backward has the softmax form and uses okl_softmax_bwd. */
int okl_rowwise(okl_ctx* ctx, int kind, const void* a, const void* b, void* out, int rows, int cols);
/* out[i] = a[i] (op) b[i % nb]; op 0 = add (Tensor.__add__, tensor.py:12-22), 1 = mul (Tensor.__mul__, tensor.py:24-39).
 * nb = n for same-shape operands, the row length for a (1,N) row broadcast, 1 for a scalar held on the device. */
int okl_binary(okl_ctx* ctx, int op, const void* a, const void* b, void* out, size_t n, size_t nb);
/* x *= alpha   (final `inputs.data / temperature`, okl.py:2890) */
int okl_scale(okl_ctx* ctx, void* x, size_t n, float alpha);

/* Conv2DLayer(transposed=True) (okl.py:815-840, 891-926) needs no kernel of its own: its forward IS
 * okl_conv_dgrad of the mirrored ordinary convolution (C = out_channels, O = in_channels, input extent = the transposed layer's
 * OUTPUT extent), its input gradient IS okl_conv_fprop of the output gradient, its filter gradient IS okl_conv_wgrad with the two
 * activations swapped.  Only the bias needs these two: y[n, c, s] += b[c] and db[c] = sum_{n, s} g[n, c, s]. */
int okl_bias_add(okl_ctx* ctx, void* y, const void* b, int N, int C, long long S);
int okl_bias_grad(okl_ctx* ctx, const void* g, void* db, int N, int C, long long S);

/* Max / AverageUnpoolingLayer (okl.py:2619-2705; kind 0 = max: the value, 1 = average: value / pool^2) and Zero / Reflection
 * PaddingLayer (okl.py:1920-1986; reflect = np.pad mode 'reflect'; backward of both = the crop) - SURVEY 8(f) N3 */
int okl_unpool_fwd(okl_ctx* ctx, int kind, const void* x, void* y, long long NC, int H, int W, int pool, int stride);
int okl_unpool_bwd(okl_ctx* ctx, int kind, const void* g, void* dx, long long NC, int H, int W, int pool, int stride);
int okl_pad2d(okl_ctx* ctx, const void* x, void* y, long long NC, int H, int W, int top, int bottom, int left, int right, int reflect);
int okl_crop2d(okl_ctx* ctx, const void* g, void* dx, long long NC, int GH, int GW, int top, int left, int h, int w);

/* LSTMLayer (okl.py:1412-1534): the four gate products are okl_dense_fwd (sigmoid / tanh fused) on
 * [inputs | prev_hidden] (okl_concat_cols), their gradients four okl_dense_bwd; these entry points are the pointwise parts:
 * cell = f c_prev + i c~, hidden = o tanh(cell); the gate pre-activation gradients and dL/dc_prev of okl.py:1464-1483, 1510;
 * dL/dh_prev = the hidden-state columns of the four input gradients summed in the reference's order (okl.py:1506-1509). */
int okl_concat_cols(okl_ctx* ctx, const void* a, const void* b, void* out, long long rows, int ca, int cb);
int okl_lstm_pointwise_fwd(okl_ctx* ctx, const void* f, const void* i, const void* c_cand, const void* o, const void* c_prev, void* cell,
                           void* hidden, size_t n);
int okl_lstm_pointwise_bwd(okl_ctx* ctx, const void* f, const void* i, const void* c_cand, const void* o, const void* c_prev,
                           const void* cell, const void* dh, const void* dc, void* dzf, void* dzi, void* dzc, void* dzo, void* dc_prev,
                           size_t n);
int okl_sum4_cols(okl_ctx* ctx, const void* a0, const void* a1, const void* a2, const void* a3, void* out, long long rows, int ld, int off,
                  int cols);

/* GRULayer (okl.py:1537-1643): gates z, r on [x | h_prev], candidate on [x | r h_prev] (okl_dense_fwd, activations fused); these
 * are its pointwise parts.  okl_gru_pointwise mode 0: hidden = (1 - z) h_prev + z h~; mode 1: out0 = dh z (1 - h~^2) (candidate
 * pre-activation gradient), out1 = dh (h~ - h_prev) z (1 - z) (update-gate one).  okl_gru_dr: the reset-gate pre-activation
 * gradient from the hidden columns of the candidate's [x | r h] gradient.  okl_gru_dinputs: dL/dx and dL/dh_prev assembled from the
 * three [x | h] gradients in the reference's order of additions (okl.py:1597-1605). */
int okl_gru_pointwise(okl_ctx* ctx, int mode, const void* z, const void* h_cand, const void* h_prev, const void* dh, void* out0, void* out1,
                      size_t n);
int okl_gru_dr(okl_ctx* ctx, const void* dcomb_h, const void* h_prev, const void* r, void* dzr, long long rows, int I, int H);
int okl_gru_dinputs(okl_ctx* ctx, const void* dcomb_z, const void* dcomb_r, const void* dcomb_h, const void* r, const void* z, const void* dh,
                    void* dx, void* dprev, long long rows, int I, int H);

/* ------------------------------------------------------------------ BatchNormLayer (okl.py:1145-1222), (B, F) inputs
 * fwd: batch mean / population variance over axis 0 (training) or the running statistics; running = momentum * batch +
 * (1 - momentum) * running; y = gamma (x - mean) / sqrt(max(var, eps) + eps) + beta, nan_to_num.  *_mean_var are 2F floats
 * [mean | var].  bwd reproduces okl.py:1193-1204 literally - the chain rule uses the RUNNING variance, x_norm the batch one. */
int okl_bn_fwd(okl_ctx* ctx, const void* x, const void* gamma, const void* beta, void* running_mean_var, void* batch_mean_var, void* y,
               int B, int F, float eps, float momentum, int training);
int okl_bn_bwd(okl_ctx* ctx, const void* x, const void* g, const void* gamma, const void* running_mean_var, const void* batch_mean_var,
               void* dx, void* dgamma, void* dbeta, int B, int F, float eps);

/* ------------------------------------------------------------------ Unfold / Fold (okl.py:525-636; im2col / fold as standalone layers)
 * unfold_fwd: out (N, C*k*k, OH*OW), row c*k*k + u*k + v, column i*OW + j (okl.py:584-605, including the reference's handling of
 * padding > 0: the clipped window goes to the top-left of a zero patch); unfold_bwd: col2im scatter-add for padding == 0
 * (okl.py:625-634); fold_permute: Fold.forward's reshape/transpose (N*C, k, k, OH, OW) -> (N*C, OH, k, OW, k) (543-547), inverse = 1
 * for Fold.backward (559-563). */
int okl_unfold_fwd(okl_ctx* ctx, const void* x, void* out, int N, int C, int H, int W, int k, int stride, int pad);
int okl_unfold_bwd(okl_ctx* ctx, const void* g, void* dx, int N, int C, int H, int W, int k, int stride);
int okl_fold_permute(okl_ctx* ctx, const void* in, void* out, long long NC, int k, int OH, int OW, int inverse);

/* ------------------------------------------------------------------ pooling (okl.py:1715-1831 + repairs R4,R5) */
int okl_pool_fwd(okl_ctx* ctx, int kind, const void* x, void* y, int N, int C, int H, int W, int pool, int stride, int layout);
/* relu_mask != 0 additionally zeroes dx where x <= 0: the fused backward of a ReLU placed right before the pooling layer
 * (g * (x > 0), okl.py:107), so that Conv -> ReLU -> MaxPool needs one backward pass over the conv output, not two. */
int okl_pool_bwd(okl_ctx* ctx, int kind, const void* x, const void* g, void* dx, int N, int C, int H, int W, int pool,
                 int stride, int layout, int relu_mask);

/* ------------------------------------------------------------------ softmax + losses (okl.py:355-399, 1380-1409, 1833-1839) */
int okl_softmax_fwd(okl_ctx* ctx, const void* x, void* p, int rows, int cols);
int okl_softmax_bwd(okl_ctx* ctx, const void* p, const void* g, void* dx, int rows, int cols);
/* loss[0] = mean_r(-log(clip(p)[r,t_r])) over `rows`; grad = (clip(p) - onehot) / denom  (denom = global batch).
 * loss_accum (optional): device scalar that gets += loss[0] (per-epoch loss sum, okl.py:2973, without a host round trip).
 * softmax_dx (optional): also writes p * (grad - sum_k grad_k p_k), the backward of the softmax layer that produced p
 * (okl.py:376-393), so a Softmax + CrossEntropy head costs one kernel for loss, loss gradient and softmax gradient. */
int okl_ce_fwd_bwd(okl_ctx* ctx, const void* p, const void* targets_i32, void* loss, void* grad, int rows, int cols, float denom,
                   void* loss_accum, void* softmax_dx);
/* Classifier head DenseLayer(K <= 256, N <= 16) -> SoftmaxActivationLayer -> CrossEntropyLoss in one launch (okl.py:26-29, 364-373,
 * 1388-1403, 376-393, 35-37): P = softmax(X W + b); loss / loss_accum / grad as okl_ce_fwd_bwd; dZ = softmax backward of grad
 * (the gradient w.r.t. the logits, from which okl_dense_bwd later forms dW, db); dX (optional) = dZ W^T, zeroed where
 * dx_relu_mask <= 0 when that pointer is given.  Same summation orders as the separate entry points. */
int okl_dense_softmax_ce(okl_ctx* ctx, const void* X, const void* W, const void* b, const void* targets_i32, void* P, void* loss,
                         void* grad, void* dZ, void* dX, const void* dx_relu_mask, int M, int K, int N, float denom,
                         void* loss_accum);
/* loss[0] = sum((o-t)^2) / n ; grad = 2 (o - t) / denom   (denom = global element count).
 * relu_outputs != 0: o is the output of a ReLU whose backward (g * (x > 0), okl.py:107; y > 0 <=> x > 0) is applied here, i.e.
 * grad is already the gradient w.r.t. the ReLU's input - one pass over the outputs instead of two. */
int okl_mse_fwd_bwd(okl_ctx* ctx, const void* o, const void* t, void* loss, void* grad, size_t n, float denom, void* loss_accum,
                    int relu_outputs);
/* the same for outputs (and gradient) in channels-last layout (N, S, C) - what the tensor-core convolutions produce - against
 * targets in the caller's (N, C, S) layout: the target tile is transposed on the fly, nothing is permuted in memory */
int okl_mse_fwd_bwd_nxc(okl_ctx* ctx, const void* o_nxc, const void* t_ncx, void* loss, void* grad_nxc, int N, int C, long long S,
                        float denom, void* loss_accum, int relu_outputs);
int okl_read_scalar(okl_ctx* ctx, const void* dev_f32, float* out);

/* ------------------------------------------------------------------ parameter updates
 * reference rule (okl.py:39-45, 744-748, 872-878, 1029-1033): gacc += g ; w -= lr * gacc  (gacc never reset).
 * lr_dev, when non-NULL, is a device fp32 scalar read at run time (lets a captured CUDA graph follow the scheduler). */
int okl_accum_update(okl_ctx* ctx, void* w, void* gacc, const void* g, size_t n, float lr, const void* lr_dev);
/* the same rule applied to `count` parameter buckets in one launch (host arrays of device pointers / element counts) */
int okl_accum_update_multi(okl_ctx* ctx, int count, void* const* w, void* const* gacc, const void* const* g, const size_t* n,
                           float lr, const void* lr_dev);
/* ... and, in the same pass, the first n16[i] parameters of bucket i (the weight matrix the BF16 GEMMs read) are written as bf16 to
 * w16[i] (NULL entry: none): the shadow is refreshed by the kernel that changes the master instead of a separate conversion pass */
int okl_accum_update_multi16(okl_ctx* ctx, int count, void* const* w, void* const* gacc, const void* const* g, const size_t* n,
                             void* const* w16, const size_t* n16, float lr, const void* lr_dev);
/* optimizers.py:140-144: v = mu v - lr g ; w += v */
int okl_sgd_update(okl_ctx* ctx, void* w, void* v, const void* g, size_t n, float lr, float momentum);
/* optimizers.py:28-42 with t supplied by the caller (it increments per update call) */
int okl_adam_update(okl_ctx* ctx, void* w, void* m, void* v, const void* g, size_t n, float lr, float beta1, float beta2,
                    float eps, int t);

/* ------------------------------------------------------------------ CUDA graphs (train-step capture) */
int okl_graph_begin(okl_ctx* ctx);
int okl_graph_end(okl_ctx* ctx, okl_graph** out);
int okl_graph_launch(okl_ctx* ctx, okl_graph* g);
int okl_graph_destroy(okl_graph* g);

/* ------------------------------------------------------------------ auxiliary lanes (concurrent branches of one step)
 * Weight / bias gradients are leaves of the backward chain: okl_aux_begin(lane) routes the following launches to an auxiliary
 * stream (with its own scratch), ordered after what is already enqueued; okl_aux_end returns to the main lane while that work
 * keeps running; okl_aux_join makes the main lane wait for all auxiliary work (done before the parameter update). */
int okl_aux_mark(okl_ctx* ctx);      /* optional: the next okl_aux_begin forks from this point instead of from "now" */
int okl_aux_begin(okl_ctx* ctx, int lane);
int okl_aux_end(okl_ctx* ctx);
int okl_aux_join(okl_ctx* ctx);

/* ------------------------------------------------------------------ timing on the compute stream */
int okl_timer_start(okl_ctx* ctx);
int okl_timer_stop(okl_ctx* ctx, float* ms);     /* records, synchronises, returns elapsed ms */
/* write `bytes` of a scratch buffer larger than L2 (between timed iterations) */
int okl_flush_l2(okl_ctx* ctx);

/* Classifier head after a BF16 conv stack: FlattenLayer -> DenseLayer(K = C * S, N <= 16) -> SoftmaxActivationLayer ->
 * CrossEntropyLoss (okl.py:643-651, 29-47, 364-393, 1389-1409) on the channels-last bf16 activation x16 (M, S, C).  Forward (P),
 * mean CE (loss, also added to *loss_accum), dL/dP (grad), dL/dlogits (dZ), the input gradient dx16 as channels-last bf16
 * (optional), dW (K, N) in the reference's row order k = c * S + s and db; wt_scratch = N * K bf16 of device scratch.  aux_lane 1 / 2:
 * the leaves (loss value, dW, db) are enqueued on that auxiliary lane (okl_aux_join before they are consumed); 0: everything in order. */
int okl_head_big_supported(int M, int C, int S, int N);
int okl_head_big(okl_ctx* ctx, const void* x16, const void* W, const void* b, const void* targets_i32, void* P, void* loss, void* grad, void* dZ,
                 void* dx16, void* dW, void* db, void* wt_scratch, int M, int C, int S, int N, float denom, void* loss_accum, int aux_lane);

/* Memory-bound companions of the BF16 tensor-core pipeline (csrc/okl_bf16_ops.cu): between tcgen05 layers activations and
 * gradients exist as channels-last bf16 only.  okl_from_bf16: bf16 -> fp32 when a host mirror / fp32 consumer asks.
 * okl_pool2_bf16_*: MaxPoolingLayer(2, 2) forward / backward on (N, H, W, C) bf16 (okl.py:1731-1772, tie rule R4); the backward
 * optionally applies the preceding ReLU's backward (okl.py:107-112) and reduces the gradient it writes to db[C] (fp32), the bias
 * gradient of the convolution in front (okl.py:870, R2).  okl_channel_sum_bf16: out[c] = sum_rows g[row, c] (bias gradients). */
int okl_from_bf16(okl_ctx* ctx, const void* src_bf16, void* dst_f32, size_t n);
int okl_pool2_bf16_supported(int C, int H, int W);
int okl_pool2_bf16_fwd(okl_ctx* ctx, const void* x16, void* y16, int N, int C, int H, int W);
int okl_pool2_bf16_bwd(okl_ctx* ctx, const void* x16, const void* g16, void* dx16, int N, int C, int H, int W, int relu_mask, void* db,
                       int aux_lane);   /* aux_lane 1 / 2: the second stage of db runs on that auxiliary lane (okl_aux_join before db is read) */
int okl_channel_sum_bf16(okl_ctx* ctx, const void* g16, void* out, long long rows, int C);
/* Crossing layout AND precision in one pass (32-channel x 128-position shared-memory tiles).  okl_to_bf16_nxc: the caller's fp32
 * (N, C, S) tensor -> the bf16 (N, S, C) operand of the tensor-core convolutions (replaces okl_permute + okl_to_bf16: the im2col +
 * cast of okl.py:691-699, 948-957 never exists).  okl_mse_fwd_bwd_nxc_bf16: MSELoss (okl.py:1833-1839) of bf16 (N, S, C) outputs
 * against fp32 (N, C, S) targets; loss (fp32 scalar, mean over N*C*S), gradient 2 (o - t) / denom as bf16 (N, S, C), optionally
 * with the backward of the ReLU that produced the outputs (okl.py:107).  C % 8 == 0.  *_rows_i32 (optional): sample n is row
 * rows[n] of the fp32 tensor - the batch of a shuffled epoch (okl.py:2957-2967) is read straight from the dataset, no gathered copy. */
int okl_to_bf16_nxc(okl_ctx* ctx, const void* src_ncx_f32, const void* src_rows_i32, void* dst_nxc_bf16, int N, int C, long long S);
int okl_mse_fwd_bwd_nxc_bf16(okl_ctx* ctx, const void* o16_nxc, const void* t_ncx, const void* t_rows_i32, void* loss, void* grad16_nxc, int N, int C,
                             long long S, float denom, void* loss_accum, int relu_outputs);

/* BF16 implicit-GEMM convolution on tcgen05, second generation (csrc/okl_conv_tc.cu): replaces the contraction of
 * Conv1DLayer/Conv2DLayer.forward (okl.py:700-711, 801-811) and of their input-gradient loops (okl.py:722-740, 854-868 with repairs
 * R1/R3).  Channels-last bf16 activations (N, *spatial, C), unit stride, C and O multiples of 64.  okl_conv2_filter_banks converts
 * the fp32 master filters (O, C, *k) into the two K-major bf16 banks once per update: w2 [O][tap][C] (forward) and wf [C][tap][O],
 * taps flipped (input gradient as a convolution of the output gradient).  y / dx (fp32) and y16 / dx16 (bf16) are both optional
 * outputs, at least one is required; d->act may be OKL_ACT_NONE or OKL_ACT_RELU; relu_mask16 (bf16, indexed like dx) zeroes dx where
 * it is <= 0 (backward of the ReLU that produced this layer's input). */
int okl_conv2_supported(okl_ctx* ctx, const okl_conv_desc* d);
int okl_conv2_filter_banks(okl_ctx* ctx, const okl_conv_desc* d, const void* W, void* w2_16, void* wf_16);
int okl_conv2_fprop(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* w2_16, const void* bias, void* y, void* y16);
int okl_conv2_dgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* g16, const void* wf_16, void* dx, void* dx16, const void* relu_mask16);
/* Small-K convolution (C * taps < 128): the first layer of an image network, e.g. Conv2DLayer(3, 64, 3, 1, 1) (okl.py:790-813,
 * 869-871), and the video block Conv3DLayer(3, 64, 3, 1, 1) (okl.py:948-985, 1018-1031).  x is the fp32 (N, C, *spatial) batch -
 * or, with x_rows_i32, the rows x_rows[n] of the whole dataset; y16 / g16 are channels-last bf16; the im2col operand is built in
 * shared memory in the UMMA layout from a staged input patch and consumed by tcgen05.mma; bias and bias gradient ride in the same
 * product as a constant-one tap.  2-D and 3-D, unit stride, O in {64, 128, 256}. */
int okl_conv_first_supported(okl_ctx* ctx, const okl_conv_desc* d);
int okl_conv_first_fprop(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* x_rows_i32, const void* W, const void* bias, void* y16);
int okl_conv_first_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* x_rows_i32, const void* g16, void* dW, void* db);
/* filter gradient (okl.py:739-741, 869-871) and, optionally, the bias gradient (okl.py:721, 870 with repair R2): dW fp32 in the
 * reference layout (O, C, *k); both operands channels-last bf16; split-K over the SMs with a fixed-order second stage. */
/* First layer with the horizontal taps pre-expanded into channels (C * KW <= 16, 2..4 filter rows, unit stride; okl.py:790-813 /
 * 869-871): okl_conv_vexp_expand writes xe16[n, h, ow, c * KW + v] = x[n, c, h, ow + v - pw] (bf16, 16 channels) from the fp32
 * (N, C, H, W) images (optionally rows x_rows_i32[n] of a dataset); okl_conv_vexp_desc gives the channels-last convolution the layer
 * then is (16 channels, KH x 1 filter) for okl_conv2_fprop / okl_conv2_wgrad; okl_conv_vexp_bank its filter bank (O, KH, 16) from the
 * (O, C, KH, KW) masters; okl_conv_vexp_dw maps that convolution's filter gradient (O, 16, KH, 1) back to (O, C, KH, KW). */
int okl_conv_vexp_supported(okl_ctx* ctx, const okl_conv_desc* d);
int okl_conv_vexp_desc(okl_ctx* ctx, const okl_conv_desc* d, okl_conv_desc* out);
int okl_conv_vexp_expand(okl_ctx* ctx, const okl_conv_desc* d, const void* x, const void* x_rows_i32, void* xe16);
int okl_conv_vexp_bank(okl_ctx* ctx, const okl_conv_desc* d, const void* W, void* w2e16);
int okl_conv_vexp_dw(okl_ctx* ctx, const okl_conv_desc* d, const void* dWe, void* dW);
int okl_conv2_wgrad_supported(okl_ctx* ctx, const okl_conv_desc* d);
int okl_conv2_wgrad(okl_ctx* ctx, const okl_conv_desc* d, const void* x16, const void* g16, void* dW, void* db);

/* ------------------------------------------------------------------ data-parallel gradient exchange (NCCL over NVLink) */
int okl_comm_unique_id(void* out128);
int okl_comm_init(okl_ctx* ctx, const void* id128, int rank, int nranks);
int okl_comm_destroy(okl_ctx* ctx);
/* Peer-memory path for small buckets: every rank allocates a symmetric buffer (okl_p2p_alloc), the 64-byte cudaIpc handles are
 * exchanged (okl_p2p_export / okl_p2p_import for every peer), gradient storage is carved out of it (okl_p2p_suballoc, same call
 * sequence on every rank); okl_allreduce_async on such memory is then ONE two-shot kernel over NVLink instead of an NCCL call. */
int okl_p2p_alloc(okl_ctx* ctx, size_t bytes);
int okl_p2p_export(okl_ctx* ctx, void* out64);
int okl_p2p_import(okl_ctx* ctx, int rank, const void* handle64);
int okl_p2p_suballoc(okl_ctx* ctx, size_t bytes, void** out);
/* sum-allreduce `count` fp32 at p on the comm stream, ordered after everything enqueued so far on the compute stream */
/* NVLS (in-switch reduction): a gradient region bound to ONE multicast object per node.  Rank 0 okl_mc_create -> POSIX fd, passed to
 * the other ranks (SCM_RIGHTS) -> okl_mc_import; every rank okl_mc_add_device, all-rank barrier, okl_mc_bind; gradient storage is
 * carved out with okl_mc_suballoc (same call sequence on every rank).  okl_allreduce_async / _fork on such memory is ONE kernel of
 * multimem.ld_reduce + multimem.st over 1/R of the bucket per rank.  Every call returns an error instead of aborting when the box
 * has no multicast support; the caller then stays on the peer-memory path. */
int okl_mc_supported(okl_ctx* ctx);
int okl_mc_create(okl_ctx* ctx, size_t bytes, int* out_fd);
int okl_mc_import(okl_ctx* ctx, int fd, size_t bytes);
int okl_mc_add_device(okl_ctx* ctx);
int okl_mc_bind(okl_ctx* ctx);
int okl_mc_suballoc(okl_ctx* ctx, size_t bytes, void** out);
int okl_allreduce_async(okl_ctx* ctx, void* p, size_t count);
/* the same, but also ordered after the auxiliary lanes (where weight gradients are produced) and ALWAYS on the comm stream -
 * peer-memory kernel included - so that it overlaps what the caller enqueues next on the compute stream (early exchange of the
 * gradients that are already complete while the first layers' backward still runs) */
int okl_allreduce_fork(okl_ctx* ctx, void* p, size_t count);
/* make the compute stream wait for every allreduce issued so far */
int okl_comm_wait(okl_ctx* ctx);
int okl_allreduce_max_host(okl_ctx* ctx, double* value);   /* small host-value reduction for max-over-ranks timing */
int okl_barrier(okl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* OKL_H */
