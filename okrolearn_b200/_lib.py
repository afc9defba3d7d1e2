"""ctypes binding of libokl_b200.so (include/okl.h) and the per-process device context.

There is NO fallback: if the shared library is missing the import raises, and if no CUDA device is present
the first operation that needs the device raises ``RuntimeError``.  Nothing here imports torch, cupy or triton.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libokl_b200.so")

OKL_OK, OKL_ERR_INVALID, OKL_ERR_CUDA, OKL_ERR_NCCL, OKL_ERR_OOM, OKL_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
FP32, TF32, BF16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_SOFTMAX = 0, 1, 2, 3, 4
ACT_ELU, ACT_SELU, ACT_LEAKY, ACT_SWISH, ACT_GELU, ACT_SOFTSIGN, ACT_HARDTANH, ACT_TANH_OUT = 10, 11, 12, 13, 14, 15, 16, 17
F32, F64, I64, I32, U8 = 0, 1, 2, 3, 4
NCX, NXC = 0, 1
POOL_MAX, POOL_AVG = 0, 1
_PRECISIONS = {"fp32": FP32, "tf32": TF32, "bf16": BF16}
_NP2OKL = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int64): I64, np.dtype(np.int32): I32,
           np.dtype(np.uint8): U8}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C okrolearn_b200/csrc` (or `python -c 'import __graft_entry__ as g; "
        "g.build()'`).  okrolearn_b200 has no CPU / PyTorch fallback.")
lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


class ConvDesc(C.Structure):
    _fields_ = [("nd", C.c_int), ("N", C.c_int), ("C", C.c_int), ("O", C.c_int), ("inp", C.c_int * 3), ("k", C.c_int * 3),
                ("s", C.c_int * 3), ("p", C.c_int * 3), ("act", C.c_int), ("x_layout", C.c_int), ("y_layout", C.c_int)]


_vp, _i, _sz, _f, _ll = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_longlong
_P = C.POINTER
# name -> (restype, argtypes); every symbol include/okl.h declares (tests/test_abi.py checks the two lists agree)
SIGNATURES = {
    "okl_version": (C.c_char_p, []),
    "okl_last_error": (C.c_char_p, []),
    "okl_ctx_create": (_i, [_i, _i, _P(_vp)]),
    "okl_ctx_destroy": (_i, [_vp]),
    "okl_ctx_sync": (_i, [_vp]),
    "okl_ctx_set_precision": (_i, [_vp, _i]),
    "okl_ctx_get_precision": (_i, [_vp, _P(_i)]),
    "okl_device_info": (_i, [_vp, _P(_i), _P(_sz), _P(_i)]),
    "okl_launch_count": (_i, [_vp, _P(C.c_ulonglong)]),
    "okl_malloc": (_i, [_vp, _sz, _P(_vp)]),
    "okl_free": (_i, [_vp, _vp]),
    "okl_pool_reserve": (_i, [_vp, _sz]),
    "okl_memset": (_i, [_vp, _vp, _i, _sz]),
    "okl_pinned_alloc": (_i, [_sz, _P(_vp)]),
    "okl_pinned_free": (_i, [_vp]),
    "okl_h2d": (_i, [_vp, _vp, _vp, _sz, _i]),
    "okl_h2d_async_raw": (_i, [_vp, _vp, _vp, _sz]),
    "okl_d2h": (_i, [_vp, _vp, _vp, _sz, _i]),
    "okl_d2h_async_raw": (_i, [_vp, _vp, _vp, _sz]),
    "okl_d2d": (_i, [_vp, _vp, _vp, _sz]),
    "okl_h2d_i32": (_i, [_vp, _vp, _vp, _sz, _i]),
    "okl_permute": (_i, [_vp, _vp, _vp, _i, _i, _ll, _i]),
    "okl_gather_rows": (_i, [_vp, _vp, _vp, _vp, _ll, _ll]),
    "okl_bias_add": (_i, [_vp, _vp, _vp, _i, _i, _ll]),
    "okl_bias_grad": (_i, [_vp, _vp, _vp, _i, _i, _ll]),
    "okl_unpool_fwd": (_i, [_vp, _i, _vp, _vp, _ll, _i, _i, _i, _i]),
    "okl_unpool_bwd": (_i, [_vp, _i, _vp, _vp, _ll, _i, _i, _i, _i]),
    "okl_pad2d": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i]),
    "okl_crop2d": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i]),
    "okl_concat_cols": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i]),
    "okl_lstm_pointwise_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    "okl_lstm_pointwise_bwd": (_i, [_vp] + [_vp] * 13 + [_sz]),
    "okl_sum4_cols": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i]),
    "okl_gru_pointwise": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    "okl_gru_dr": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i]),
    "okl_gru_dinputs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i]),
    "okl_bn_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i]),
    "okl_bn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f]),
    "okl_unfold_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i]),
    "okl_unfold_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "okl_fold_permute": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _i]),
    "okl_gather_batch": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _ll]),
    "okl_batch_host_index": (_i, [_vp, _vp]),
    "okl_batch_prefetch": (_i, [_vp, _i, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _i, _i, _vp]),
    "okl_batch_wait": (_i, [_vp, _i]),
    "okl_loss_fetch_async": (_i, [_vp, _i, _vp]),
    "okl_loss_fetch_wait": (_i, [_vp, _i, C.POINTER(C.c_float)]),
    "okl_dense_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "okl_dense_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "okl_to_bf16": (_i, [_vp, _vp, _vp, _sz]),
    "okl_dense_fwd_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "okl_dense_bwd_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "okl_conv_bf16_supported": (_i, [_vp, _P(ConvDesc)]),
    "okl_conv_fprop_bf16": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv_dgrad_bf16": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv_wgrad_bf16": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_head_big_supported": (_i, [_i, _i, _i, _i]),
    "okl_head_big": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _i]),
    "okl_mc_supported": (_i, [_vp]),
    "okl_mc_create": (_i, [_vp, _sz, _P(_i)]),
    "okl_mc_import": (_i, [_vp, _i, _sz]),
    "okl_mc_add_device": (_i, [_vp]),
    "okl_mc_bind": (_i, [_vp]),
    "okl_mc_suballoc": (_i, [_vp, _sz, _P(_vp)]),
    "okl_from_bf16": (_i, [_vp, _vp, _vp, _sz]),
    "okl_pool2_bf16_supported": (_i, [_i, _i, _i]),
    "okl_pool2_bf16_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i]),
    "okl_pool2_bf16_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i]),
    "okl_channel_sum_bf16": (_i, [_vp, _vp, _vp, _ll, _i]),
    "okl_to_bf16_nxc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _ll]),
    "okl_mse_fwd_bwd_nxc_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _f, _vp, _i]),
    "okl_conv2_supported": (_i, [_vp, _P(ConvDesc)]),
    "okl_conv2_filter_banks": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp]),
    "okl_conv2_fprop": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv2_dgrad": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv_first_supported": (_i, [_vp, _P(ConvDesc)]),
    "okl_conv_first_fprop": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv_first_wgrad": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv_vexp_supported": (_i, [_vp, _P(ConvDesc)]),
    "okl_conv_vexp_desc": (_i, [_vp, _P(ConvDesc), _P(ConvDesc)]),
    "okl_conv_vexp_expand": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp]),
    "okl_conv_vexp_bank": (_i, [_vp, _P(ConvDesc), _vp, _vp]),
    "okl_conv_vexp_dw": (_i, [_vp, _P(ConvDesc), _vp, _vp]),
    "okl_conv2_wgrad_supported": (_i, [_vp, _P(ConvDesc)]),
    "okl_conv2_wgrad": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp]),
    "okl_tc_gemm": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _i]),
    "okl_conv_out_extent": (_i, [_P(ConvDesc), _P(_i)]),
    "okl_conv_fprop": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp]),
    "okl_conv_dgrad": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp]),
    "okl_conv_wgrad": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp]),
    "okl_conv_relu_pool_supported": (_i, [_P(ConvDesc)]),
    "okl_conv_relu_pool_fwd": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_conv_relu_pool_wgrad": (_i, [_vp, _P(ConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "okl_act_fwd": (_i, [_vp, _i, _vp, _vp, _sz]),
    "okl_act_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _sz]),
    "okl_act2_fwd": (_i, [_vp, _i, _vp, _vp, _sz, _f, _f]),
    "okl_act2_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _f, _f]),
    "okl_prelu_alpha_grad": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "okl_rowwise": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i]),
    "okl_binary": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _sz]),
    "okl_scale": (_i, [_vp, _vp, _sz, _f]),
    "okl_pool_fwd": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i]),
    "okl_pool_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i]),
    "okl_softmax_fwd": (_i, [_vp, _vp, _vp, _i, _i]),
    "okl_softmax_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "okl_ce_fwd_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "okl_dense_softmax_ce": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "okl_mse_fwd_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _f, _vp, _i]),
    "okl_mse_fwd_bwd_nxc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _f, _vp, _i]),
    "okl_read_scalar": (_i, [_vp, _vp, _P(_f)]),
    "okl_accum_update": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _vp]),
    "okl_accum_update_multi": (_i, [_vp, _i, _P(_vp), _P(_vp), _P(_vp), _P(_sz), _f, _vp]),
    "okl_accum_update_multi16": (_i, [_vp, _i, _P(_vp), _P(_vp), _P(_vp), _P(_sz), _P(_vp), _P(_sz), _f, _vp]),
    "okl_sgd_update": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f]),
    "okl_adam_update": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _i]),
    "okl_graph_begin": (_i, [_vp]),
    "okl_graph_end": (_i, [_vp, _P(_vp)]),
    "okl_graph_launch": (_i, [_vp, _vp]),
    "okl_graph_destroy": (_i, [_vp]),
    "okl_aux_mark": (_i, [_vp]),
    "okl_aux_begin": (_i, [_vp, _i]),
    "okl_aux_end": (_i, [_vp]),
    "okl_aux_join": (_i, [_vp]),
    "okl_timer_start": (_i, [_vp]),
    "okl_timer_stop": (_i, [_vp, _P(_f)]),
    "okl_flush_l2": (_i, [_vp]),
    "okl_comm_unique_id": (_i, [_vp]),
    "okl_comm_init": (_i, [_vp, _vp, _i, _i]),
    "okl_comm_destroy": (_i, [_vp]),
    "okl_p2p_alloc": (_i, [_vp, _sz]),
    "okl_p2p_export": (_i, [_vp, _vp]),
    "okl_p2p_import": (_i, [_vp, _i, _vp]),
    "okl_p2p_suballoc": (_i, [_vp, _sz, _P(_vp)]),
    "okl_allreduce_async": (_i, [_vp, _vp, _sz]),
    "okl_allreduce_fork": (_i, [_vp, _vp, _sz]),
    "okl_comm_wait": (_i, [_vp]),
    "okl_allreduce_max_host": (_i, [_vp, _P(C.c_double)]),
    "okl_barrier": (_i, [_vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the .so does not export what okl.h declares
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.okl_last_error().decode("utf-8", "replace")


def check(status: int):
    """Map okl_status to the exception the reference's users would see (SURVEY 8(b) error conventions)."""
    if status == OKL_OK:
        return
    msg = last_error()
    if status in (OKL_ERR_INVALID, OKL_ERR_UNSUPPORTED):
        raise ValueError(msg)
    if status == OKL_ERR_OOM:
        raise MemoryError(msg)
    low = msg.lower()
    if any(k in low for k in ("unspecified launch failure", "illegal memory access", "illegal instruction", "launch failed", "trap",
                              "device-side assert", "misaligned address")):
        # a kernel fault (incl. the bounded-wait traps of the tensor-core kernels) poisons the CUDA context: every later call fails too
        msg += "  [the CUDA context of this process is no longer usable after a kernel fault: restart the process]"
    raise RuntimeError(msg)


class Context:
    """One CUDA device + compute stream + comm stream (okl_ctx).  One per process."""

    def __init__(self, device: int | None = None, precision: str | int | None = None):
        if device is None:
            device = int(os.environ.get("OKL_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if precision is None:
            precision = os.environ.get("OKL_PRECISION", "fp32")
        if isinstance(precision, str):
            if precision.lower() not in _PRECISIONS:
                raise ValueError(f"unknown precision {precision!r}; use fp32, tf32 or bf16")
            precision = _PRECISIONS[precision.lower()]
        h = _vp()
        check(lib.okl_ctx_create(device, precision, C.byref(h)))
        self.h = h
        self.device = device
        self.epoch = 0            # bumped whenever device memory may have changed behind host mirrors (graph replay)
        self.capturing = False
        self.rank, self.nranks = 0, 1
        self.p2p_ready = False
        sm, mem, cc = _i(), _sz(), _i()
        check(lib.okl_device_info(self.h, C.byref(sm), C.byref(mem), C.byref(cc)))
        self.sm_count, self.total_mem, self.cc = sm.value, mem.value, cc.value

    # -- memory
    def malloc(self, nbytes: int) -> int:
        p = _vp()
        check(lib.okl_malloc(self.h, max(int(nbytes), 1), C.byref(p)))
        return p.value

    def free(self, ptr: int):
        if ptr and self.h:
            lib.okl_free(self.h, ptr)

    def sync(self):
        check(lib.okl_ctx_sync(self.h))

    @property
    def precision(self) -> int:
        v = _i()
        check(lib.okl_ctx_get_precision(self.h, C.byref(v)))
        return v.value

    @precision.setter
    def precision(self, p):
        if isinstance(p, str):
            p = _PRECISIONS[p.lower()]
        check(lib.okl_ctx_set_precision(self.h, int(p)))

    def launch_count(self) -> int:
        v = C.c_ulonglong()
        check(lib.okl_launch_count(self.h, C.byref(v)))
        return v.value

    # -- graphs
    def graph_begin(self):
        check(lib.okl_graph_begin(self.h))
        self.capturing = True

    def graph_end(self) -> int:
        g = _vp()
        self.capturing = False
        check(lib.okl_graph_end(self.h, C.byref(g)))
        return g.value

    def graph_launch(self, g: int):
        check(lib.okl_graph_launch(self.h, g))
        self.epoch += 1

    # -- timing
    def timer_start(self):
        check(lib.okl_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = _f()
        check(lib.okl_timer_stop(self.h, C.byref(ms)))
        return ms.value

    # -- comm
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        check(lib.okl_comm_init(self.h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib.okl_comm_unique_id(buf))
        return buf.raw


_ctx: Context | None = None


def get_ctx() -> Context:
    """The process-wide context (created on first use; raises RuntimeError when there is no CUDA device)."""
    global _ctx
    if _ctx is None:
        _ctx = Context()
    return _ctx


HAVE_CONV2_WGRAD = os.environ.get("OKL_CONV_TC2_WGRAD", "1") != "0"


def set_precision(p):
    get_ctx().precision = p


def np_to_okl_dtype(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _NP2OKL:
        raise ValueError(f"unsupported host dtype {dt}")
    return _NP2OKL[dt]
