"""Layers of the hot path with the reference's names, constructor signatures, attributes and ``forward`` /
``backward(dL_dout, lr)`` protocol (reference: src/okrolearn/okrolearn.py, cited per class as okl.py:lines).  Every
``forward``/``backward`` enqueues CUDA kernels through the C ABI (include/okl.h) on device-resident tensors; nothing is
computed on the host.  Where the shipped reference backward raises ``ValueError`` for every shape (Conv1D/2D/3D, pooling)
the semantics are those of the repaired loops R1-R5, which is what oracle/okl_oracle.py states.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, get_ctx, ConvDesc, NCX, NXC, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, POOL_MAX, POOL_AVG
from .tensor import Tensor, _DevBuf, _numel

_ALIGN = 64          # floats: parameter sub-tensors start on 256-byte boundaries inside a bucket
# BF16 mode: tensors between tensor-core layers are stored as channels-last bf16 only (OKL_BF16_ONLY=0 restores fp32 + bf16 shadow)
_BF16_ONLY = __import__("os").environ.get("OKL_BF16_ONLY", "1") != "0"
# first layer with the horizontal taps expanded into channels (okl_conv_vexp_*): parity-tested, but measured SLOWER than the small-K
# kernels at the CIFAR shape (expand 25 + fprop 97 + wgrad 63 + remap 6 us against 61 + 70 us: the implicit-GEMM kernel has a per-tile
# floor of ~2.5 us whatever the reduction length - see DESIGN.md) - off unless asked for
_VEXP = __import__("os").environ.get("OKL_CONV_VEXP_LAYER", "0") != "0"
_CT_3D = __import__("os").environ.get("OKL_CT_3D", "1") != "0"          # Conv3D with >= 8 channels on the BF16 tcgen05 kernels
_ANY_C = __import__("os").environ.get("OKL_CT_ANYC", "1") != "0"       # BF16 tcgen05 convolutions for channel counts that are multiples of 8


def _as_tensor(x):
    return x if isinstance(x, Tensor) else Tensor(x)


def _lr_args(lr):
    """(float lr, device pointer or None): a captured CUDA graph reads the learning rate from device memory."""
    ctx = get_ctx()
    dev = getattr(ctx, "lr_dev", None)
    return float(lr), (dev if (ctx.capturing and dev) else None)


def _set_input_grad(inputs, dx):
    """Several reference layers also accumulate dL/dinputs into ``inputs.grad`` (okl.py:110, 751, 1035, 1768, 2094)."""
    if not inputs.requires_grad:
        return
    if inputs._grad is None:
        inputs._grad = dx
    else:
        prev = inputs._grad if isinstance(inputs._grad, Tensor) else Tensor(inputs._grad)
        inputs._grad = prev + dx


class _SubBuf:
    """A slice of a larger device block (kept alive through ``owner``)."""
    __slots__ = ("ptr", "nbytes", "owner")

    def __init__(self, owner, off_bytes, nbytes):
        self.ptr, self.nbytes, self.owner = owner.ptr + off_bytes, nbytes, owner


class _BorrowedPtr:
    __slots__ = ("ptr", "nbytes")

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes


class ParamArena:
    """All parameter buckets of a small network in ONE device block each for params / accumulated gradients / step gradients:
    the data-parallel exchange is a single all-reduce and the update a single kernel, whatever the number of layers."""

    def __init__(self, buckets):
        ctx = get_ctx()
        self.buckets = list(buckets)
        self.total = sum(b.total for b in self.buckets)
        self.param, self.gacc = _DevBuf(ctx, self.total * 4), _DevBuf(ctx, self.total * 4)
        self.grad = None
        if ctx.nranks > 1 and getattr(ctx, "p2p_ready", False):
            # gradients live in the rank's multicast-bound region (NVLS: the switch reduces them) or, without multicast support, in
            # its symmetric peer-mapped buffer: either way the all-reduce is one NVLink kernel
            p = C.c_void_p()
            if getattr(ctx, "mc_ready", False) and lib.okl_mc_suballoc(ctx.h, self.total * 4, C.byref(p)) == 0:
                self.grad = _BorrowedPtr(p.value, self.total * 4)
            elif lib.okl_p2p_suballoc(ctx.h, self.total * 4, C.byref(p)) == 0:
                self.grad = _BorrowedPtr(p.value, self.total * 4)
        if self.grad is None:
            self.grad = _DevBuf(ctx, self.total * 4)
        check(lib.okl_memset(ctx.h, self.grad.ptr, 0, self.total * 4))
        off = 0
        for b in self.buckets:
            b.rehome(self, off)
            off += b.total
        self._tail_exchanged = False

    def exchange_tail(self):
        """Data parallel: start the all-reduce of every bucket but the first (their gradients are complete once the backward pass
        reaches the first parametric layer) on the comm stream, overlapping that layer's backward."""
        ctx = get_ctx()
        head = self.buckets[0].total
        if ctx.nranks > 1 and len(self.buckets) > 1 and not self._tail_exchanged:
            check(lib.okl_allreduce_fork(ctx.h, self.grad.ptr + head * 4, self.total - head))
            self._tail_exchanged = True


class _ParamBucket:
    """All parameters of one layer in ONE device block (params | gacc | grads share the same offsets), so the layer's
    gradient exchange is one NCCL allreduce and its update is one kernel.

    The update is the reference's effective rule (okl.py:39-45, 744-748, 872-878, 1029-1033):
    ``grad = grad + g`` (never reset) ; ``w -= lr * grad``."""

    def __init__(self, tensors):
        self.tensors = tensors           # list of host-initialised Tensor objects (weights, biases)
        self.ready = False

    def materialize(self):
        if self.ready:
            return
        ctx = get_ctx()
        offs, total = [], 0
        for t in self.tensors:
            offs.append(total)
            total += (t.size + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = total
        self.offs = offs
        self.param = _DevBuf(ctx, total * 4)
        self.gacc = _DevBuf(ctx, total * 4)
        self.grad = _DevBuf(ctx, total * 4)
        for b in (self.param, self.gacc, self.grad):
            check(lib.okl_memset(ctx.h, b.ptr, 0, total * 4))
        self.gviews = []
        for t, off in zip(self.tensors, offs):
            pending = None if t._dev_valid else t._host
            t._buf, t._off, t._bound = self.param, off * 4, True
            t._dev_valid = False if pending is not None else True
            if pending is not None:
                t._upload()
            self.gviews.append(Tensor._view(self.gacc, off * 4, t._shape, t._np_dtype))
        self.ready = True

    arena = None

    def rehome(self, arena, elem_off):
        """Move this bucket's storage into ``arena`` at element offset ``elem_off`` (contents preserved)."""
        ctx = get_ctx()
        for t in self.tensors:
            t._dptr(NCX)                                   # flush pending host-side assignments into the old location first
        nb = self.total * 4
        newp, newa, newg = (_SubBuf(buf, elem_off * 4, nb) for buf in (arena.param, arena.gacc, arena.grad))
        check(lib.okl_d2d(ctx.h, newp.ptr, self.param.ptr, nb))
        check(lib.okl_d2d(ctx.h, newa.ptr, self.gacc.ptr, nb))
        self.param, self.gacc, self.grad, self.arena = newp, newa, newg, arena
        for t, gv, off in zip(self.tensors, self.gviews, self.offs):
            t._buf, t._off = self.param, off * 4
            gv._buf, gv._off = self.gacc, off * 4

    def ptr(self, i):
        """Device pointer of parameter i (re-uploading if the user assigned ``.data``)."""
        return self.tensors[i]._dptr(NCX)

    def grad_ptr(self, i):
        return self.grad.ptr + self.offs[i] * 4

    def ptr16(self, i):
        """bf16 shadow of parameter i, refreshed when the fp32 master changed (after every update)."""
        return self.tensors[i]._b16()

    def _live_shadow(self):
        """The weight tensor (parameter 0, at the bucket's start) if its bf16 shadow is current and can be rewritten in place by the
        update kernel; None otherwise."""
        t = self.tensors[0]
        if (self.offs[0] == 0 and t.size % 4 == 0 and t.size > 0 and not t._b16_primary and getattr(t, "_bf16", None) is not None and
                t._dev_valid and t._bf16_ver == (t._dev_ver, t._layout) and t._bf16.nbytes >= t.size * 2 and len(t._shape) == 2):
            return t
        return None

    def _sync_user_grads(self):
        """``layer.weights.grad = None`` / ``= ndarray`` by the user replaces the accumulator."""
        ctx = get_ctx()
        for t, gv, off in zip(self.tensors, self.gviews, self.offs):
            if t._grad is gv:
                continue
            if t._grad is None:
                if getattr(self, "_stepped", False):
                    check(lib.okl_memset(ctx.h, self.gacc.ptr + off * 4, 0, t.size * 4))
            else:
                arr = np.ascontiguousarray(np.asarray(t._grad, dtype=np.float64).reshape(-1))
                if arr.size != t.size:
                    raise ValueError("assigned .grad has the wrong number of elements")
                check(lib.okl_h2d(ctx.h, self.gacc.ptr + off * 4, arr.ctypes.data, arr.size, _lib.F64))
            gv._touch()

    def step(self, lr, deferred=False):
        """Exchange (data parallel) and apply the accumulated-gradient update for the whole bucket."""
        ctx = get_ctx()
        self._sync_user_grads()
        if ctx.nranks > 1 and not (deferred and self.arena is not None):     # an arena exchanges all its buckets at once
            check(lib.okl_allreduce_async(ctx.h, self.grad.ptr, self.total))
        if deferred:
            self._pending_lr = lr
            return
        self.apply(lr)

    def apply(self, lr):
        ctx = get_ctx()
        if ctx.nranks > 1:
            check(lib.okl_comm_wait(ctx.h))
        lrf, lrp = _lr_args(lr)
        for t in self.tensors:
            t._dptr(NCX)                         # flush a pending host-side assignment before updating in place
        check(lib.okl_accum_update(ctx.h, self.param.ptr, self.gacc.ptr, self.grad.ptr, self.total, lrf, lrp))
        self._stepped = True
        for t, gv in zip(self.tensors, self.gviews):
            t._touch()
            gv._touch()
            t._grad = gv
        self._pending_lr = None


class _ParamLayer:
    _aux_lane = 0                  # NeuralNetwork planner: >0 = compute dW/db (leaves of the backward chain) on that auxiliary lane
    _dx_relu_mask = False          # NeuralNetwork planner: a fused ReLU precedes this layer; its backward is applied in our dX epilogue
    _defer_update = False          # set by NeuralNetwork when updates are applied after the whole backward (DP overlap)
    _fused_act = ACT_NONE          # set by NeuralNetwork's planner: activation fused into this layer's epilogue
    _fused_head = False            # NeuralNetwork planner: last Dense(N<=16) -> Softmax; the loss's kernel computes the head
    _big_head = False              # ... and that head follows Flatten of a channels-last bf16 conv stack (okl_head_big)

    def _finish_update(self):
        b = self._bucket
        if getattr(b, "_pending_lr", None) is not None:
            b.apply(b._pending_lr)


def apply_pending_updates(layers):
    """All deferred bucket updates of a network (after joining the auxiliary lanes and the data-parallel exchange): ONE all-reduce
    and ONE kernel when the buckets live in a ParamArena, else one multi-tensor launch."""
    ctx = get_ctx()
    check(lib.okl_aux_join(ctx.h))
    bs = [l._bucket for l in layers if isinstance(l, _ParamLayer) and getattr(l._bucket, "_pending_lr", None) is not None]
    if not bs:
        return
    lrf, lrp = _lr_args(bs[0]._pending_lr)
    for b in bs:
        for t in b.tensors:
            t._dptr(NCX)
    arena = bs[0].arena
    if arena is not None and len(bs) == len(arena.buckets) and all(b.arena is arena for b in bs):
        if ctx.nranks > 1:
            if arena._tail_exchanged:                      # the rest is already in flight: only the first bucket is left
                check(lib.okl_allreduce_async(ctx.h, arena.grad.ptr, arena.buckets[0].total))
                arena._tail_exchanged = False
            else:
                check(lib.okl_allreduce_async(ctx.h, arena.grad.ptr, arena.total))
            check(lib.okl_comm_wait(ctx.h))
        check(lib.okl_accum_update(ctx.h, arena.param.ptr, arena.gacc.ptr, arena.grad.ptr, arena.total, lrf, lrp))
    else:
        if ctx.nranks > 1:
            for b in bs:
                if b.arena is not None:                    # partial arena: exchange bucket by bucket
                    check(lib.okl_allreduce_async(ctx.h, b.grad.ptr, b.total))
            check(lib.okl_comm_wait(ctx.h))
        n = len(bs)
        w = (C.c_void_p * n)(*[b.param.ptr for b in bs])
        ga = (C.c_void_p * n)(*[b.gacc.ptr for b in bs])
        g = (C.c_void_p * n)(*[b.grad.ptr for b in bs])
        cnt = (C.c_size_t * n)(*[b.total for b in bs])
        # BF16 mode: a weight matrix whose bf16 shadow is live (the GEMMs of this step read it) gets the shadow rewritten by the
        # update kernel itself - no separate fp32 -> bf16 pass over the parameters after every step
        shadowed = [b._live_shadow() if ctx.precision == _lib.BF16 else None for b in bs]
        if any(t is not None for t in shadowed):
            w16 = (C.c_void_p * n)(*[t._bf16.ptr if t is not None else None for t in shadowed])
            n16 = (C.c_size_t * n)(*[t.size if t is not None else 0 for t in shadowed])
            check(lib.okl_accum_update_multi16(ctx.h, n, w, ga, g, cnt, w16, n16, lrf, lrp))
        else:
            shadowed = [None] * n
            check(lib.okl_accum_update_multi(ctx.h, n, w, ga, g, cnt, lrf, lrp))
        for b, t16 in zip(bs, shadowed):
            b._fresh_shadow = t16
    for b in bs:
        b._stepped = True
        for t, gv in zip(b.tensors, b.gviews):
            t._touch()
            gv._touch()
            t._grad = gv
        t16 = getattr(b, "_fresh_shadow", None)
        if t16 is not None:
            t16._bf16_ver = (t16._dev_ver, t16._layout)         # written together with the master
        b._fresh_shadow = None
        b._pending_lr = None


# ====================================================================================================== Dense
class DenseLayer(_ParamLayer):
    """okl.py:18-54.  weights (in,out) = randn*0.1 float64, biases (1,out) zeros; Y = X.W + b;
    backward: dW = X^T.G, db = sum_rows G, dX = G.W^T (pre-update W), then the in-layer accumulated update."""

    def __init__(self, input_size, output_size):
        self.weights = Tensor(np.random.randn(input_size, output_size) * 0.1)
        self.biases = Tensor(np.zeros((1, output_size)))
        self._bucket = _ParamBucket([self.weights, self.biases])

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        ctx = get_ctx()
        self._bucket.materialize()
        K, N = self.weights.shape
        if inputs.ndim != 2 or inputs.shape[1] != K:
            raise ValueError(f"shapes {inputs.shape} and {self.weights.shape} not aligned")
        self.inputs = inputs
        M = inputs.shape[0]
        out = Tensor._empty((M, N), np.float64)
        self._head_pre = None
        if self._fused_head:
            # classifier head: the probabilities are produced by the loss's fused kernel (okl_dense_softmax_ce) when a
            # CrossEntropyLoss consumes them next; any other reader triggers the ordinary dense + softmax kernel
            self._bf16_path = False
            layer = self

            def compute(t):
                c = get_ctx()
                check(lib.okl_dense_fwd(c.h, layer.inputs._dptr(NCX), layer._bucket.ptr(0), layer._bucket.ptr(1), t._ptr_raw(),
                                        M, K, N, layer._fused_act))
                t._touch()
            out._lazy_fn = compute
            out._head_of = self
            self.outputs = out
            return out
        self._bf16_path = (ctx.precision == _lib.BF16 and M >= 64 and M % 8 == 0 and K >= 64 and N >= 64 and K % 8 == 0 and N % 8 == 0
                           and self._fused_act <= ACT_TANH)
        if self._bf16_path:                      # operands and result carry bf16 shadows: no per-call conversion
            check(lib.okl_dense_fwd_bf16(ctx.h, inputs._b16(), self._bucket.ptr16(0), self._bucket.ptr(1), out._ptr_raw(),
                                         out._new_b16(), M, K, N, self._fused_act))
        else:
            check(lib.okl_dense_fwd(ctx.h, inputs._dptr(NCX), self._bucket.ptr(0), self._bucket.ptr(1), out._ptr_raw(), M, K, N,
                                    self._fused_act))
        self.outputs = out
        return out

    def backward(self, dL_dout: Tensor, lr: float):
        dL_dout = _as_tensor(dL_dout)
        ctx = get_ctx()
        K, N = self.weights.shape
        M = self.inputs.shape[0]
        if dL_dout.shape != (M, N):
            raise ValueError(f"dL_dout shape {dL_dout.shape} does not match the layer output {(M, N)}")
        need_dx = getattr(self, "_need_dx", True)
        b = self._bucket
        pre, self._head_pre = getattr(self, "_head_pre", None), None
        if pre is not None and pre[0] is dL_dout and len(pre) > 2 and pre[2]:
            b.step(lr, self._defer_update)                   # big fused head: dX, dW and db were all produced with the loss
            return pre[1]
        if pre is not None and pre[0] is dL_dout:            # fused classifier head: dX already produced with the loss
            dX = pre[1]
            aux = bool(self._aux_lane and self._defer_update)
            if aux:
                check(lib.okl_aux_mark(ctx.h))
                check(lib.okl_aux_begin(ctx.h, self._aux_lane))
            try:
                check(lib.okl_dense_bwd(ctx.h, self.inputs._dptr(NCX), b.ptr(0), dL_dout._dptr(NCX), None, b.grad_ptr(0),
                                        b.grad_ptr(1), M, K, N, None))
                b.step(lr, self._defer_update)
            finally:
                if aux:
                    check(lib.okl_aux_end(ctx.h))
            return dX
        dX = Tensor._empty((M, K), np.float64) if need_dx else None
        xp, gp, wp = self.inputs._dptr(NCX), dL_dout._dptr(NCX), b.ptr(0)
        mask = xp if (need_dx and self._dx_relu_mask) else None
        if getattr(self, "_bf16_path", False) and ctx.precision == _lib.BF16:
            x16, g16, w16 = self.inputs._b16(), dL_dout._b16(), b.ptr16(0)
            dxp, dx16 = (dX._ptr_raw(), dX._new_b16()) if need_dx else (None, None)
            aux = self._aux_lane and self._defer_update and need_dx
            if aux:
                check(lib.okl_aux_mark(ctx.h))
            if need_dx:
                check(lib.okl_dense_bwd_bf16(ctx.h, x16, w16, g16, gp, dxp, dx16, None, None, M, K, N, mask))
            if aux:
                check(lib.okl_aux_begin(ctx.h, self._aux_lane))
            try:
                check(lib.okl_dense_bwd_bf16(ctx.h, x16, w16, g16, gp, None, None, b.grad_ptr(0), b.grad_ptr(1), M, K, N, None))
                b.step(lr, True if aux else self._defer_update)
            finally:
                if aux:
                    check(lib.okl_aux_end(ctx.h))
            return dX
        if self._aux_lane and self._defer_update and need_dx:
            check(lib.okl_aux_mark(ctx.h))
            check(lib.okl_dense_bwd(ctx.h, xp, wp, gp, dX._ptr_raw(), None, None, M, K, N, mask))     # critical path first
            check(lib.okl_aux_begin(ctx.h, self._aux_lane))          # dW, db (+ their all-reduce) are leaves: own lane
            try:
                check(lib.okl_dense_bwd(ctx.h, xp, wp, gp, None, b.grad_ptr(0), b.grad_ptr(1), M, K, N, None))
                b.step(lr, True)
            finally:
                check(lib.okl_aux_end(ctx.h))
        else:
            check(lib.okl_dense_bwd(ctx.h, xp, wp, gp, dX._ptr_raw() if need_dx else None, b.grad_ptr(0), b.grad_ptr(1), M, K, N, mask))
            b.step(lr, self._defer_update)
        return dX

    def get_params(self):
        return {'weights': self.weights.data, 'biases': self.biases.data}

    def set_params(self, params):
        self.weights.data = params['weights']
        self.biases.data = params['biases']


# ====================================================================================================== Conv
def _tup(v, nd):
    return tuple(int(a) for a in v) if isinstance(v, (tuple, list)) else (int(v),) * nd


class _ConvBase(_ParamLayer):
    _nd = 2
    _np_dtype = np.float64

    def _init(self, in_channels, out_channels, kernel_size, stride, padding):
        nd = self._nd
        self.in_channels, self.out_channels = in_channels, out_channels
        self._k, self._s, self._p = _tup(kernel_size, nd), _tup(stride, nd), _tup(padding, nd)

    def _desc(self, x_shape, x_layout, y_layout, act=ACT_NONE):
        nd = self._nd
        if len(x_shape) != nd + 2:
            raise ValueError(f"expected a {nd + 2}-D input (N, C, *spatial), got shape {x_shape}")
        if x_shape[1] != self.in_channels:
            raise ValueError(f"input has {x_shape[1]} channels, layer expects {self.in_channels}")
        d = ConvDesc()
        d.nd, d.N, d.C, d.O = nd, x_shape[0], self.in_channels, self.out_channels
        for i in range(nd):
            d.inp[i], d.k[i], d.s[i], d.p[i] = x_shape[2 + i], self._k[i], self._s[i], self._p[i]
        d.act, d.x_layout, d.y_layout = act, x_layout, y_layout
        out = (C.c_int * 3)()
        check(lib.okl_conv_out_extent(C.byref(d), out))
        return d, tuple(out[i] for i in range(nd))

    def _tc_eligible(self):
        """Shapes the tcgen05 implicit-GEMM convolution takes (okl_tc_conv_supported / okl_conv2_supported): 1-D/2-D, unit stride,
        padding <= k-1, channel counts that are multiples of 32 - or, in BF16 mode, 2-D layers with 2..4 filter rows and channel
        counts that are multiples of 8 (TMA zero-fills the partial 64-channel chunks; the bias gradient needs a free chunk slot of
        the weight-gradient kernel or 8 x 2^n output channels).  Those layers run channels-last in TF32/BF16 mode."""
        if not (all(s == 1 for s in self._s) and all(p <= k - 1 for p, k in zip(self._p, self._k))):
            return False
        Cc, Oc = self.in_channels, self.out_channels
        if self._nd == 3:
            # Conv3D on the BF16 kernels (5-D TMA boxes, one more tap loop; the weight gradient one depth tap per CTA group).  Whether
            # the tiles fit the volume is only known with the input shape: _forward asks okl_conv2_supported and falls back.
            if get_ctx().precision != _lib.BF16 or not _CT_3D or Cc % 8 or Oc % 8 or not (2 <= self._k[1] <= 4):
                return False
            free_slot = ((self._k[2] * ((Cc + 63) // 64)) % 2 == 1 and 2 * self._p[1] <= self._k[1] - 1 and 2 * self._p[0] <= self._k[0] - 1)
            o8 = Oc // 8
            return bool(free_slot or (o8 & (o8 - 1)) == 0)
        if Cc % 32 == 0 and Oc % 32 == 0:
            return True
        if get_ctx().precision != _lib.BF16 or self._nd != 2 or Cc % 8 or Oc % 8 or not (2 <= self._k[0] <= 4) or not _ANY_C:
            return False
        free_slot = (self._k[1] * ((Cc + 63) // 64)) % 2 == 1 and 2 * self._p[0] <= self._k[0] - 1
        o8 = Oc // 8
        return bool(free_slot or (o8 & (o8 - 1)) == 0)

    _fused_pool = None             # NeuralNetwork planner: (relu, pool) layers fused into this conv (first block, no dX needed)

    def _forward_fused_pool(self, inputs):
        """Conv -> ReLU -> MaxPool(2,2) in one kernel: returns the POOLED tensor; the un-pooled activation is a lazy placeholder."""
        ctx = get_ctx()
        d, oe = self._desc(inputs.shape, NCX, NCX, ACT_RELU)
        N, O = inputs.shape[0], self.out_channels
        PH, PW = oe[0] // 2, oe[1] // 2
        pooled = Tensor._empty((N, O, PH, PW), np.float64)
        mask = _DevBuf(ctx, max(N * O * PH * PW, 1))
        check(lib.okl_conv_relu_pool_fwd(ctx.h, C.byref(d), inputs._dptr(NCX), self._bucket.ptr(0), self._bucket.ptr(1),
                                         pooled._ptr_raw(), mask.ptr))
        self._pool_mask, self._fwd_layouts = mask, (NCX, NCX)
        full = Tensor._view(None, 0, (N, O) + oe, self._np_dtype, bound=False)      # never written by the fused kernel
        full._dev_valid = False
        layer = self

        def recompute(t):
            dd, _ = layer._desc(layer.inputs.shape, NCX, NCX, ACT_RELU)
            t._buf, t._off = _DevBuf(get_ctx(), t.size * 4), 0
            check(lib.okl_conv_fprop(get_ctx().h, C.byref(dd), layer.inputs._dptr(NCX), layer._bucket.ptr(0), layer._bucket.ptr(1),
                                     t._ptr_raw()))
            t._touch()
        full._lazy_fn = recompute
        relu, pool = self._fused_pool
        relu.inputs = relu.outputs = full
        pool.inputs, pool.output = full, pooled
        pooled._fused_from = self
        return pooled, full

    def _backward_fused_pool(self, pooled_grad, lr):
        ctx = get_ctx()
        x = self.inputs
        d, oe = self._desc(x.shape, NCX, NCX)
        b = self._bucket
        check(lib.okl_conv_relu_pool_wgrad(ctx.h, C.byref(d), x._dptr(NCX), _as_tensor(pooled_grad)._dptr(NCX), self._pool_mask.ptr,
                                           b.grad_ptr(0), b.grad_ptr(1)))
        b.step(lr, self._defer_update)
        return None

    def _forward(self, inputs):
        inputs = _as_tensor(inputs)
        ctx = get_ctx()
        self._bucket.materialize()
        self.inputs = inputs
        if self._fused_pool is not None:
            pooled, full = self._forward_fused_pool(inputs)
            self._fused_out = pooled
            return full
        tc = ctx.precision != _lib.FP32 and self._tc_eligible()
        if tc and self._nd == 3:                 # 3-D: the tile / volume fit is shape dependent
            dd, _ = self._desc(inputs.shape, NXC, NXC, self._fused_act)
            tc = bool(inputs.shape[0] > 0 and self._fused_act in (ACT_NONE, ACT_RELU) and lib.okl_conv2_supported(ctx.h, C.byref(dd)))
        if tc:
            xl = yl = NXC
        else:
            xl = inputs._layout if inputs.ndim >= 3 and inputs._dev_valid else NCX
            if xl == NXC and self.in_channels > 1 and self._nd == 3:
                xl = NCX
            yl = getattr(self, "_out_layout", NCX)
            if (yl == NCX and getattr(self, "_out_nxc_if_first", False) and xl == NCX and ctx.precision == _lib.BF16 and _BF16_ONLY and
                    inputs.shape[0] > 0 and self._fused_act in (ACT_NONE, ACT_RELU) and not getattr(self, "_need_dx", True)):
                # last parametric layer of a network (only activations and the loss follow): if the small-K tcgen05 kernels take the
                # shape, the output is written channels-last bf16 and the loss consumes it in that format
                dd, _ = self._desc(inputs.shape, NCX, NXC, self._fused_act)
                if lib.okl_conv_first_supported(ctx.h, C.byref(dd)):
                    yl = NXC
        d, oe = self._desc(inputs.shape, xl, yl, self._fused_act)
        oshape = (inputs.shape[0], self.out_channels) + oe
        self._first_tc = bool(ctx.precision == _lib.BF16 and _BF16_ONLY and xl == NCX and yl == NXC and inputs.shape[0] > 0 and
                              self._fused_act in (ACT_NONE, ACT_RELU) and lib.okl_conv_first_supported(ctx.h, C.byref(d)))
        if self._first_tc:                       # tiny-C first layer on tcgen05: fp32 images in, channels-last bf16 out
            out = Tensor._empty_b16(oshape, self._np_dtype, NXC)
            xr = inputs._rows_pending()                   # gather-free batch: the kernel reads the dataset rows through the index
            xp, xi = xr if xr is not None else (inputs._dptr(NCX), None)
            self._vexp = None
            if _VEXP and lib.okl_conv_vexp_supported(ctx.h, C.byref(d)):
                # C * KW <= 16: the horizontal taps are expanded into 16 bf16 channels by one memory-bound pass and the layer runs as
                # a channels-last convolution with a KH x 1 filter on the TMA-fed implicit-GEMM kernels (no software im2col)
                de = ConvDesc()
                check(lib.okl_conv_vexp_desc(ctx.h, C.byref(d), C.byref(de)))
                xe = _DevBuf(ctx, max(inputs.shape[0] * de.inp[0] * de.inp[1], 1) * 32)
                check(lib.okl_conv_vexp_expand(ctx.h, C.byref(d), xp, xi, xe.ptr))
                check(lib.okl_conv2_fprop(ctx.h, C.byref(de), xe.ptr, self._vexp_bank(d), self._bucket.ptr(1), None, out._bf16.ptr))
                self._vexp = (de, xe)
            else:
                check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), xp, xi, self._bucket.ptr(0), self._bucket.ptr(1), out._bf16.ptr))
            self._fwd_layouts, self._bf16_path, self._conv2 = (xl, yl), False, False
            return out
        b16_ok = bool(ctx.precision == _lib.BF16 and xl == NXC and yl == NXC and inputs.shape[0] > 0)
        self._conv2 = bool(b16_ok and self._fused_act in (ACT_NONE, ACT_RELU) and lib.okl_conv2_supported(ctx.h, C.byref(d)))
        self._bf16_path = bool(self._conv2 or (b16_ok and lib.okl_conv_bf16_supported(ctx.h, C.byref(d))))
        if self._conv2:                          # second-generation tcgen05 kernel (okl_conv_tc.cu): slab reuse, 256-pixel tiles, TMA stores
            w2, _ = self._filter_banks(d)
            if _BF16_ONLY:                       # the output exists as channels-last bf16 only; fp32 is converted if somebody asks
                out = Tensor._empty_b16(oshape, self._np_dtype, NXC)
                check(lib.okl_conv2_fprop(ctx.h, C.byref(d), inputs._b16(NXC), w2, self._bucket.ptr(1), None, out._bf16.ptr))
            else:
                out = Tensor._empty(oshape, self._np_dtype, layout=yl)
                check(lib.okl_conv2_fprop(ctx.h, C.byref(d), inputs._b16(NXC), w2, self._bucket.ptr(1), out._ptr_raw(), out._new_b16()))
            self._fwd_layouts = (xl, yl)
            return out
        out = Tensor._empty(oshape, self._np_dtype, layout=yl)
        if self._bf16_path:                    # activations carry channels-last bf16 shadows: 2-byte operands, no conversions
            check(lib.okl_conv_fprop_bf16(ctx.h, C.byref(d), inputs._b16(NXC), self._bucket.ptr(0), self._bucket.ptr(1), out._ptr_raw(),
                                          out._new_b16()))
        else:
            check(lib.okl_conv_fprop(ctx.h, C.byref(d), inputs._dptr(xl), self._bucket.ptr(0), self._bucket.ptr(1), out._ptr_raw()))
        self._fwd_layouts = (xl, yl)
        return out

    def _backward(self, dL_dout, lr):
        if self._fused_pool is not None:
            return self._backward_fused_pool(dL_dout, lr)
        dL_dout = _as_tensor(dL_dout)
        ctx = get_ctx()
        x = self.inputs
        xl, yl = self._fwd_layouts
        d, oe = self._desc(x.shape, xl, yl)
        if dL_dout.shape != (x.shape[0], self.out_channels) + oe:
            raise ValueError(f"dL_dout shape {dL_dout.shape} does not match the layer output {(x.shape[0], self.out_channels) + oe}")
        b = self._bucket
        need_dx = getattr(self, "_need_dx", True)
        if getattr(self, "_first_tc", False) and ctx.precision == _lib.BF16 and not need_dx:
            vx = getattr(self, "_vexp", None)
            if vx is not None:                            # filter + bias gradient of the expanded convolution, mapped back to (O, C, KH, KW)
                de, xe = vx
                dwe = _DevBuf(ctx, self.out_channels * 16 * self._k[0] * 4)
                check(lib.okl_conv2_wgrad(ctx.h, C.byref(de), xe.ptr, dL_dout._b16(NXC), dwe.ptr, b.grad_ptr(1)))
                check(lib.okl_conv_vexp_dw(ctx.h, C.byref(d), dwe.ptr, b.grad_ptr(0)))
                self._vexp = None
            else:
                xr = x._rows_pending()
                xp, xi = xr if xr is not None else (x._dptr(NCX), None)
                check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), xp, xi, dL_dout._b16(NXC), b.grad_ptr(0), b.grad_ptr(1)))
            b.step(lr, self._defer_update)
            return None
        aux = self._aux_lane and self._defer_update and need_dx
        b16 = getattr(self, "_bf16_path", False) and ctx.precision == _lib.BF16
        if b16 and getattr(self, "_conv2", False):
            return self._backward_conv2(dL_dout, lr, d, need_dx, aux)
        g, xp, wp = dL_dout._dptr(yl), x._dptr(xl), b.ptr(0)
        if b16:
            x16, g16 = x._b16(NXC), dL_dout._b16(NXC)
        dX = None
        if need_dx:
            if aux:
                check(lib.okl_aux_mark(ctx.h))
            dX = Tensor._empty(x.shape, self._dx_dtype(x), layout=xl)
            mask = xp if self._dx_relu_mask else None
            if b16:
                check(lib.okl_conv_dgrad_bf16(ctx.h, C.byref(d), g16, wp, dX._ptr_raw(), dX._new_b16(), mask))
            else:
                check(lib.okl_conv_dgrad(ctx.h, C.byref(d), g, wp, dX._ptr_raw(), mask))   # critical path first; pre-update filters
        if aux:
            check(lib.okl_aux_begin(ctx.h, self._aux_lane))          # dW, db (+ their all-reduce) are leaves: own lane
        try:
            if b16:
                check(lib.okl_conv_wgrad_bf16(ctx.h, C.byref(d), x16, g16, g, b.grad_ptr(0), b.grad_ptr(1)))
            else:
                check(lib.okl_conv_wgrad(ctx.h, C.byref(d), xp, g, b.grad_ptr(0), b.grad_ptr(1)))
            b.step(lr, True if aux else self._defer_update)
        finally:
            if aux:
                check(lib.okl_aux_end(ctx.h))
        return dX

    def _backward_conv2(self, dL_dout, lr, d, need_dx, aux):
        """Backward of the second-generation BF16 path: every operand and result is channels-last bf16 (no fp32 activation /
        gradient tensors at all); dX first (critical path), dW / db on the auxiliary lane."""
        ctx = get_ctx()
        x, b = self.inputs, self._bucket
        x16, g16 = x._b16(NXC), dL_dout._b16(NXC)
        dX = None
        if need_dx:
            if aux:
                check(lib.okl_aux_mark(ctx.h))
            _, wf = self._filter_banks(d)
            mask16 = x16 if self._dx_relu_mask else None
            if _BF16_ONLY:
                dX = Tensor._empty_b16(x.shape, self._dx_dtype(x), NXC)
                check(lib.okl_conv2_dgrad(ctx.h, C.byref(d), g16, wf, None, dX._bf16.ptr, mask16))
            else:
                dX = Tensor._empty(x.shape, self._dx_dtype(x), layout=NXC)
                check(lib.okl_conv2_dgrad(ctx.h, C.byref(d), g16, wf, dX._ptr_raw(), dX._new_b16(), mask16))
        if aux:
            check(lib.okl_aux_begin(ctx.h, self._aux_lane))
        try:
            db_done = getattr(dL_dout, "_db_done", None) is self          # the producer of dL_dout already reduced it (pool backward)
            rows = _numel(dL_dout.shape) // self.out_channels
            if _lib.HAVE_CONV2_WGRAD and lib.okl_conv2_wgrad_supported(ctx.h, C.byref(d)):
                check(lib.okl_conv2_wgrad(ctx.h, C.byref(d), x16, g16, b.grad_ptr(0), None if db_done else b.grad_ptr(1)))
            else:
                check(lib.okl_conv_wgrad_bf16(ctx.h, C.byref(d), x16, g16, None, b.grad_ptr(0), None))
                if not db_done:
                    check(lib.okl_channel_sum_bf16(ctx.h, g16, b.grad_ptr(1), rows, self.out_channels))
            b.step(lr, True if aux else self._defer_update)
        finally:
            if aux:
                check(lib.okl_aux_end(ctx.h))
        return dX

    def _vexp_bank(self, d):
        """bf16 filter bank (O, KH, 16) of the first layer's expanded form, refreshed when the fp32 masters changed (see _filter_banks)."""
        ctx = get_ctx()
        wt = self._bucket.tensors[0]
        wp = self._bucket.ptr(0)
        ver = ("capture", id(getattr(ctx, "_capture_keep", None)), wt._dev_ver) if ctx.capturing else (wt._dev_ver, ctx.epoch)
        n = self.out_channels * self._k[0] * 16 * 2
        if getattr(self, "_vbank_buf", None) is None or self._vbank_buf.nbytes < n:
            self._vbank_buf, self._vbank_ver = _DevBuf(ctx, n), None
        if self._vbank_ver != ver:
            check(lib.okl_conv_vexp_bank(ctx.h, C.byref(d), wp, self._vbank_buf.ptr))
            self._vbank_ver = ver
        return self._vbank_buf.ptr

    def _filter_banks(self, d):
        """bf16 K-major filter banks of the tcgen05 convolution - [O][tap][C] for fprop, flipped [C][tap][O] for dgrad - refreshed by
        ONE kernel whenever the fp32 master filters changed (once per train step, after the update)."""
        ctx = get_ctx()
        self._bank_desc = d
        wt = self._bucket.tensors[0]
        wp = self._bucket.ptr(0)                                   # flushes a pending host assignment first
        # inside a captured step the refresh is recorded once per capture (forward) and replayed with the graph; outside, a graph
        # replay (ctx.epoch) or an update (_dev_ver) since the last refresh makes the banks stale
        ver = ("capture", id(getattr(ctx, "_capture_keep", None)), wt._dev_ver) if ctx.capturing else (wt._dev_ver, ctx.epoch)
        if getattr(self, "_bank_bufs", None) is None or self._bank_bufs[0].nbytes < wt.size * 2:
            self._bank_bufs = (_DevBuf(ctx, wt.size * 2), _DevBuf(ctx, wt.size * 2))
            self._bank_ver = None
        if self._bank_ver != ver:
            check(lib.okl_conv2_filter_banks(ctx.h, C.byref(d), wp, self._bank_bufs[0].ptr, self._bank_bufs[1].ptr))
            self._bank_ver = ver
        return self._bank_bufs[0].ptr, self._bank_bufs[1].ptr

    def _dx_dtype(self, x):
        return np.float64 if self._nd == 2 else (x._np_dtype if x._np_dtype in (np.float32, np.float64) else np.float64)


class Conv1DLayer(_ConvBase):
    """okl.py:669-764.  weights (O,C,k) float32 * 0.1, biases (O,1) float32; output float32."""
    _nd, _np_dtype = 1, np.float32

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        self._init(in_channels, out_channels, kernel_size, stride, padding)
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weights = Tensor(np.random.randn(out_channels, in_channels, kernel_size).astype(np.float32) * 0.1)
        self.biases = Tensor(np.zeros((out_channels, 1), dtype=np.float32))
        self._bucket = _ParamBucket([self.weights, self.biases])

    def forward(self, inputs: Tensor):
        self.output = self._forward(inputs)
        return self.output

    def backward(self, dL_dout: Tensor, lr: float):
        dX = self._backward(dL_dout, lr)
        if dX is not None:
            _set_input_grad(self.inputs, dX)            # okl.py:750
        return dX

    def get_params(self):
        return {'weights': self.weights.data, 'biases': self.biases.data}

    def set_params(self, params):
        self.weights.data = params['weights']
        self.biases.data = params['biases']


class Conv2DLayer(_ConvBase):
    """okl.py:767-933.  filters (O,C,kh,kw) float64 * 0.1, biases (O,1); int-or-tuple kernel/stride/padding; output float64.
    ``transposed=True`` (okl.py:815-840, 891-926): filters (C_in, C_out, kh, kw); the forward is the data-gradient
    kernel of the mirrored ordinary convolution, the input gradient its forward kernel and the filter gradient its weight-gradient
    kernel with the activations swapped - no kernel of its own.  As shipped, padding > 0 raises (the output is allocated at the
    cropped size and then scattered into, okl.py:821-832); the shipped backward raises for every shape (okl.py:904-912 broadcast,
    924 bias shape) - implemented with the intended sums, cross-checked with torch.autograd like the other conv backward paths."""
    _nd, _np_dtype = 2, np.float64

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, transposed=False):
        self._init(in_channels, out_channels, kernel_size, stride, padding)
        self.kernel_size, self.stride, self.padding = self._k, self._s, self._p
        self.transposed = transposed
        if transposed:
            self.filters = Tensor(np.random.randn(in_channels, out_channels, *self._k) * 0.1)
        else:
            self.filters = Tensor(np.random.randn(out_channels, in_channels, *self._k) * 0.1)
        self.biases = Tensor(np.zeros((out_channels, 1)))
        self._bucket = _ParamBucket([self.filters, self.biases])

    def _tc_eligible(self):
        return (not self.transposed) and super()._tc_eligible()

    def _mirror_desc(self, N, out_hw):
        """descriptor of the ordinary convolution (C = out_channels -> O = in_channels) whose dgrad / fprop / wgrad are this
        transposed layer's forward / input gradient / filter gradient; its input extent is our OUTPUT extent"""
        d = ConvDesc()
        d.nd, d.N, d.C, d.O = 2, N, self.out_channels, self.in_channels
        for i in range(2):
            d.inp[i], d.k[i], d.s[i], d.p[i] = out_hw[i], self._k[i], self._s[i], 0
        d.act, d.x_layout, d.y_layout = ACT_NONE, NCX, NCX
        return d

    def forward_transposed(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 4:
            raise ValueError(f"expected a 4-D input (N, C, H, W), got shape {inputs.shape}")
        if inputs.shape[1] != self.in_channels:
            raise ValueError(f"input has {inputs.shape[1]} channels, layer expects {self.in_channels}")
        if self._p[0] > 0 or self._p[1] > 0:
            raise ValueError("operands could not be broadcast together: forward_transposed with padding > 0 scatters into an output "
                             "allocated at the cropped size (okl.py:821-832 raises the same way)")
        ctx = get_ctx()
        self._bucket.materialize()
        self.inputs = inputs
        N, _, H, W = inputs.shape
        oh, ow = (H - 1) * self._s[0] + self._k[0], (W - 1) * self._s[1] + self._k[1]
        d = self._mirror_desc(N, (oh, ow))
        out = Tensor._empty((N, self.out_channels, oh, ow), np.float64)
        if out.size:
            check(lib.okl_conv_dgrad(ctx.h, C.byref(d), inputs._dptr(NCX), self._bucket.ptr(0), out._ptr_raw(), None))
            check(lib.okl_bias_add(ctx.h, out._ptr_raw(), self._bucket.ptr(1), N, self.out_channels, oh * ow))
            if self._fused_act != ACT_NONE:            # planner fused a following ReLU / sigmoid into "this layer's epilogue"
                check(lib.okl_act_fwd(ctx.h, self._fused_act, out._ptr_raw(), out._ptr_raw(), out.size))
        self.outputs = out
        return out

    def backward_transposed(self, dL_dout: Tensor, lr: float):
        dL_dout = _as_tensor(dL_dout)
        ctx = get_ctx()
        x = self.inputs
        N, _, H, W = x.shape
        oh, ow = (H - 1) * self._s[0] + self._k[0], (W - 1) * self._s[1] + self._k[1]
        if dL_dout.shape != (N, self.out_channels, oh, ow):
            raise ValueError(f"dL_dout shape {dL_dout.shape} does not match the layer output {(N, self.out_channels, oh, ow)}")
        d = self._mirror_desc(N, (oh, ow))
        b = self._bucket
        gp = dL_dout._dptr(NCX)
        need_dx = getattr(self, "_need_dx", True)
        dX = Tensor._empty(x.shape, np.float64) if need_dx else None
        if need_dx and dX.size:
            check(lib.okl_conv_fprop(ctx.h, C.byref(d), gp, b.ptr(0), None, dX._ptr_raw()))
            if self._dx_relu_mask:                      # a ReLU fused into the previous layer: its backward applied here
                check(lib.okl_act_bwd(ctx.h, ACT_RELU, x._dptr(NCX), dX._ptr_raw(), dX._ptr_raw(), dX.size))
        if N:
            check(lib.okl_conv_wgrad(ctx.h, C.byref(d), gp, x._dptr(NCX), b.grad_ptr(0), None))
            check(lib.okl_bias_grad(ctx.h, gp, b.grad_ptr(1), N, self.out_channels, oh * ow))
        b.step(lr, self._defer_update)
        return dX

    def forward(self, inputs: Tensor):
        if self.transposed:
            return self.forward_transposed(inputs)
        self.outputs = self._forward(inputs)
        # fused first block: what flows on to the (pass-through) ReLU and MaxPool layers is already the pooled tensor
        return self._fused_out if self._fused_pool is not None else self.outputs

    def forward_normal(self, inputs: Tensor):
        return self.forward(inputs)

    def backward(self, dL_dout: Tensor, lr: float):
        if self.transposed:
            return self.backward_transposed(dL_dout, lr)
        return self._backward(dL_dout, lr)

    def backward_normal(self, dL_dout: Tensor, lr: float):
        return self._backward(dL_dout, lr)

    def get_params(self):
        return {'filters': self.filters.data, 'biases': self.biases.data}

    def set_params(self, params):
        self.filters.data = params['filters']
        self.biases.data = params['biases']


class Conv3DLayer(_ConvBase):
    """okl.py:936-1049.  filters (O,C,kd,kh,kw) float32 * 0.1, biases (O,1) float32; output float32."""
    _nd, _np_dtype = 3, np.float32

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        self._init(in_channels, out_channels, kernel_size, stride, padding)
        self.kernel_size, self.stride, self.padding = self._k, self._s, self._p
        self.filters = Tensor(np.random.randn(out_channels, in_channels, *self._k).astype(np.float32) * 0.1)
        self.biases = Tensor(np.zeros((out_channels, 1), dtype=np.float32))
        self._bucket = _ParamBucket([self.filters, self.biases])

    def forward(self, inputs: Tensor):
        self.output = self._forward(inputs)
        return self.output

    def backward(self, dL_dout: Tensor, lr: float):
        dX = self._backward(dL_dout, lr)
        if dX is not None:
            _set_input_grad(self.inputs, dX)            # okl.py:1035
        return dX

    def get_params(self):
        return {'filters': self.filters.data, 'biases': self.biases.data}

    def set_params(self, params):
        self.filters.data = params['filters']
        self.biases.data = params['biases']


# ====================================================================================================== activations
def _float_dtype(t):
    return t._np_dtype if t._np_dtype in (np.dtype(np.float32), np.dtype(np.float64)) else np.dtype(np.float64)


class _Activation:
    _act = ACT_NONE
    _bwd_ref = "inputs"          # which stored tensor the derivative is evaluated from
    _fused_into_prev = False     # NeuralNetwork planner: the producing layer's epilogue already applied this activation
    _bwd_fused_into_next = False # NeuralNetwork planner: the following pooling layer's backward already applied relu'

    _in_fused_block = False       # NeuralNetwork planner: part of a fused Conv->ReLU->MaxPool block (pure pass-through)

    def _fwd(self, inputs):
        inputs = _as_tensor(inputs)
        if self._in_fused_block:
            return inputs                      # .inputs / .outputs were set by the conv layer (lazy un-pooled activation)
        self.inputs = inputs
        if self._fused_into_prev:
            return inputs                      # the tensor already holds act(x); relu backward uses y > 0 <=> x > 0
        ctx = get_ctx()
        out = Tensor._empty(inputs.shape, _float_dtype(inputs), layout=inputs._layout if inputs._dev_valid else NCX)
        src = inputs._dptr()
        out._layout = inputs._layout
        check(lib.okl_act_fwd(ctx.h, self._act, src, out._ptr_raw(), inputs.size))
        return out

    def _bwd(self, dL_dout, ref):
        dL_dout = _as_tensor(dL_dout)
        ctx = get_ctx()
        if dL_dout.shape != ref.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dout.shape} {ref.shape}")
        lay = ref._layout if ref._dev_valid else NCX
        rp = ref._dptr()
        lay = ref._layout
        dx = Tensor._empty(ref.shape, np.result_type(_float_dtype(dL_dout), _float_dtype(ref)), layout=lay)
        check(lib.okl_act_bwd(ctx.h, self._act, rp, dL_dout._dptr(lay), dx._ptr_raw(), ref.size))
        return dx

    def get_params(self):
        return None

    def set_params(self, params):
        pass


class ReLUActivationLayer(_Activation):
    """okl.py:94-120: max(0, x); backward g * (x > 0), also accumulated into inputs.grad.
    The shipped forward goes through np.vectorize, which truncates the whole output to int64 when x.flat[0] <= 0
    (SURVEY.md fact 5); this implementation always returns the intended float max(0, x)."""
    _act = ACT_RELU

    def forward(self, inputs: Tensor):
        self.outputs = self._fwd(inputs)
        return self.outputs

    _precomputed = None          # gradient tensor an MSELoss already multiplied by relu' (NeuralNetwork fast path)

    def backward(self, dL_dout: Tensor, lr: float):
        if self._in_fused_block:
            return dL_dout
        pre, self._precomputed = self._precomputed, None
        dx = _as_tensor(dL_dout) if (self._bwd_fused_into_next or (pre is not None and pre is dL_dout)) else self._bwd(dL_dout, self.inputs)
        if not self._fused_into_prev:            # fused: `inputs` IS the post-activation tensor, no separate pre-activation
            _set_input_grad(self.inputs, dx)
        return dx


class SigmoidActivationLayer(_Activation):
    """okl.py:500-522: 1/(1+exp(-x)); backward g * y * (1 - y) from the stored output."""
    _act = ACT_SIGMOID

    def forward(self, inputs: Tensor):
        self.outputs = self._fwd(inputs)
        return self.outputs

    def backward(self, dL_dout: Tensor, lr: float = None):
        return self._bwd(dL_dout, self.outputs)


class TanhActivationLayer(_Activation):
    """okl.py:2070-2108: tanh(x); backward g * (1 - tanh(x)^2) recomputed from the input; sets inputs.grad."""
    _act = ACT_TANH

    def forward(self, inputs: Tensor):
        self.output = self._fwd(inputs)
        return self.output

    def backward(self, dL_dout: Tensor, lr: float):
        if self._fused_into_prev:
            raise RuntimeError("tanh cannot be fused: its backward needs the pre-activation input")
        dx = self._bwd(dL_dout, self.inputs)
        _set_input_grad(self.inputs, dx)
        return dx


# ---- the reference's other elementwise activations: one device kernel pair, per-layer formulas in the kernel
class _Act2:
    """y = f(x); backward g * f'(x) evaluated from the stored INPUT.  Sub-classes fix the kernel id and which side effects the
    reference has on ``inputs.grad``."""
    _act = 0
    _sets_input_grad = "accumulate"       # "accumulate" (okl.py:81, 151, 187, 219), "assign" (GELU, okl.py:304), None (Softsign)

    def _params(self):
        return 0.0, 1.0

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        self.inputs = inputs
        ctx = get_ctx()
        out = Tensor._empty(inputs.shape, _float_dtype(inputs), layout=inputs._layout if inputs._dev_valid else NCX)
        src = inputs._dptr()
        out._layout = inputs._layout
        a, sc = self._params()
        check(lib.okl_act2_fwd(ctx.h, self._act, src, out._ptr_raw(), inputs.size, float(a), float(sc)))
        self.outputs = out
        return out

    def _dx(self, dL_dout):
        dL_dout = _as_tensor(dL_dout)
        ref = self.inputs
        if dL_dout.shape != ref.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dout.shape} {ref.shape}")
        ctx = get_ctx()
        rp = ref._dptr()
        dx = Tensor._empty(ref.shape, np.result_type(_float_dtype(dL_dout), _float_dtype(ref)), layout=ref._layout)
        a, sc = self._params()
        check(lib.okl_act2_bwd(ctx.h, self._act, rp, dL_dout._dptr(ref._layout), dx._ptr_raw(), ref.size, float(a), float(sc)))
        return dx, dL_dout

    def backward(self, dL_dout: Tensor, lr: float = None):
        dx, _ = self._dx(dL_dout)
        if self._sets_input_grad == "accumulate":
            _set_input_grad(self.inputs, dx)
        elif self._sets_input_grad == "assign" and self.inputs.requires_grad:
            self.inputs._grad = dx
        return dx

    def get_params(self):
        return None

    def set_params(self, params):
        pass


class ELUActivationLayer(_Act2):
    """okl.py:57-91: x if x > 0 else alpha (exp(x) - 1)."""
    _act = _lib.ACT_ELU

    def __init__(self, alpha=1.0):
        self.alpha = alpha

    def _params(self):
        return self.alpha, 1.0


class SELUActivationLayer(_Act2):
    """okl.py:168-196: scale * ELU_alpha(x)."""
    _act = _lib.ACT_SELU

    def __init__(self, alpha=1.6732632423543772848170429916717, scale=1.0507009873554804934193349852946):
        self.alpha, self.scale = alpha, scale

    def _params(self):
        return self.alpha, self.scale


class LeakyReLUActivationLayer(_Act2):
    """okl.py:199-227: x if x > 0 else alpha x.  (The shipped backward goes through np.vectorize, which truncates the derivative to
    integers when x.flat[0] > 0 - the same dtype-inference quirk as ReLU's forward; this implements the float derivative.)"""
    _act = _lib.ACT_LEAKY

    def __init__(self, alpha: float = 0.01):
        self.alpha = alpha

    def _params(self):
        return self.alpha, 1.0


class PReLUActivationLayer(LeakyReLUActivationLayer):
    """okl.py:230-268: LeakyReLU whose slope is learned: d_alpha = sum(g * x * [x <= 0]); alpha -= lr * d_alpha inside backward."""

    def __init__(self, alpha: float = 0.01):
        self.alpha = alpha
        self.d_alpha = 0

    def backward(self, dL_dout: Tensor, lr: float):
        dx, g = self._dx(dL_dout)                      # uses the PRE-update alpha (okl.py:249 before 253)
        ctx = get_ctx()
        buf = _DevBuf(ctx, 4)
        check(lib.okl_prelu_alpha_grad(ctx.h, self.inputs._dptr(), g._dptr(self.inputs._layout), self.inputs.size, buf.ptr))
        v = C.c_float()
        check(lib.okl_read_scalar(ctx.h, buf.ptr, C.byref(v)))
        self.d_alpha = np.float64(v.value)
        self.alpha -= lr * self.d_alpha
        _set_input_grad(self.inputs, dx)
        return dx

    def get_params(self):
        return {'alpha': self.alpha}

    def set_params(self, params):
        if params and 'alpha' in params:
            self.alpha = params['alpha']


class SwishActivationLayer(_Act2):
    """okl.py:123-165: x * sigmoid(clip(x, -88, 88)); derivative swish + sigmoid (1 - swish), clipped to +-1e9."""
    _act = _lib.ACT_SWISH

    def forward(self, inputs: Tensor):
        out = super().forward(inputs)
        self.sigmoid_values = None                     # the reference keeps sigmoid(x) here; recomputed in the backward kernel
        return out


class GELUActivationLayer(_Act2):
    """okl.py:271-312: tanh form with the argument clipped to +-15; the derivative is the reference's expression as written."""
    _act = _lib.ACT_GELU
    _sets_input_grad = "assign"

    def __init__(self):
        self.sqrt_2_over_pi = np.sqrt(2 / np.pi)
        self.tanh_clip_value = 15


class SoftsignActivationLayer(_Act2):
    """okl.py:315-338: x / (1 + |x|)."""
    _act = _lib.ACT_SOFTSIGN
    _sets_input_grad = None


class HardTanhActivationLayer(_Act2):
    """okl.py:2111-2153: clip(x, -1, 1); gradient passes where -1 <= x <= 1."""
    _act = _lib.ACT_HARDTANH

    def __init__(self):
        self.inputs = self.output = self.dl_dinputs = self.dl_doutputs = None

    def forward(self, inputs: Tensor):
        self.output = super().forward(inputs)
        return self.output


class LinearActivationLayer:
    """okl.py:1989-2008: identity; backward passes the gradient through and accumulates it into inputs.grad."""

    def forward(self, inputs: Tensor):
        self.inputs = _as_tensor(inputs)
        return inputs

    def backward(self, dL_dout: Tensor, lr: float):
        dx = _as_tensor(dL_dout)
        _set_input_grad(self.inputs, dx)
        return dx

    def get_params(self):
        return None

    def set_params(self, params):
        pass


class SoftminActivationLayer:
    """okl.py:402-442: exp(-(x + rowmax)) normalised (+ nan_to_num) exactly as written; the Jacobian has the softmax form."""

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 2:
            raise ValueError(f"softmin expects a 2-D (batch, classes) input, got shape {inputs.shape}")
        ctx = get_ctx()
        out = Tensor._empty(inputs.shape, _float_dtype(inputs))
        check(lib.okl_rowwise(ctx.h, 0, inputs._dptr(NCX), None, out._ptr_raw(), inputs.shape[0], inputs.shape[1]))
        self.outputs = out
        return out

    def backward(self, dL_dout: Tensor, lr: float = None):
        dL_dout = _as_tensor(dL_dout)
        p = self.outputs
        if dL_dout.shape != p.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dout.shape} {p.shape}")
        ctx = get_ctx()
        dx = Tensor._empty(p.shape, np.float64)
        check(lib.okl_softmax_bwd(ctx.h, p._dptr(NCX), dL_dout._dptr(NCX), dx._ptr_raw(), p.shape[0], p.shape[1]))
        return dx

    def get_params(self):
        return None

    def set_params(self, params: dict):
        pass


class LogSoftmaxActivationLayer:
    """okl.py:445-497: (x - max) - log(sum(exp(clip(x - max, -500, 500)))) in float64; backward = the reference's Jacobian product
    g_j - sum_k exp(ls_k) g_k."""

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 2:
            raise ValueError(f"log-softmax expects a 2-D (batch, classes) input, got shape {inputs.shape}")
        ctx = get_ctx()
        out = Tensor._empty(inputs.shape, np.float64)
        check(lib.okl_rowwise(ctx.h, 1, inputs._dptr(NCX), None, out._ptr_raw(), inputs.shape[0], inputs.shape[1]))
        self.outputs = out
        return out

    def backward(self, dL_dout: Tensor, lr: float = None):
        dL_dout = _as_tensor(dL_dout)
        ls = self.outputs
        if dL_dout.shape != ls.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dout.shape} {ls.shape}")
        ctx = get_ctx()
        dx = Tensor._empty(ls.shape, np.float64)
        check(lib.okl_rowwise(ctx.h, 2, ls._dptr(NCX), dL_dout._dptr(NCX), dx._ptr_raw(), ls.shape[0], ls.shape[1]))
        return dx

    def get_params(self):
        return None

    def set_params(self, params: dict):
        pass


class SoftmaxActivationLayer:
    """okl.py:355-399: stable softmax over axis 1 (+ nan_to_num); backward = Jacobian product, computed in closed form
    p * (g - sum_k g_k p_k) instead of the reference's B x C x C Python triple loop."""

    _fused_into_prev = False   # NeuralNetwork planner: the preceding (small) DenseLayer's kernel already applied the softmax

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        ctx = get_ctx()
        if inputs.ndim != 2:
            raise ValueError(f"softmax expects a 2-D (batch, classes) input, got shape {inputs.shape}")
        self.inputs = inputs
        if self._fused_into_prev:
            self.outputs = inputs
            return inputs
        out = Tensor._empty(inputs.shape, _float_dtype(inputs))
        check(lib.okl_softmax_fwd(ctx.h, inputs._dptr(NCX), out._ptr_raw(), inputs.shape[0], inputs.shape[1]))
        self.outputs = out
        return out

    _precomputed = None        # (loss gradient, dz) filled by CrossEntropyLoss.forward in NeuralNetwork's fast path

    def backward(self, dL_dout: Tensor, lr: float = None):
        pre, self._precomputed = self._precomputed, None
        if pre is not None and dL_dout is pre[0]:
            return pre[1]
        dL_dout = _as_tensor(dL_dout)
        ctx = get_ctx()
        p = self.outputs
        if dL_dout.shape != p.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dout.shape} {p.shape}")
        dx = Tensor._empty(p.shape, np.float64)
        check(lib.okl_softmax_bwd(ctx.h, p._dptr(NCX), dL_dout._dptr(NCX), dx._ptr_raw(), p.shape[0], p.shape[1]))
        return dx

    def get_params(self):
        return None

    def set_params(self, params: dict):
        pass


# ====================================================================================================== recurrent cell
class RNNLayer:
    """okl.py:1646-1697: one step of an Elman cell, hidden = tanh(x Wx + h_prev Wh + b).  The whole cell runs on
    the Dense kernels: two ``okl_dense_fwd`` and two ``okl_dense_bwd`` (each gives a weight gradient AND an input gradient).  As
    shipped, backward only ACCUMULATES into ``.grad`` - it never updates the weights - which is the bucket update with lr = 0;
    it returns (dL/dx, dL/dh_prev) and adds them to ``inputs.grad`` / ``prev_hidden.grad``."""

    def __init__(self, input_size, hidden_size):
        self.input_size, self.hidden_size = input_size, hidden_size
        self.Wh = Tensor(np.random.randn(hidden_size, hidden_size) * 0.01)
        self.Wx = Tensor(np.random.randn(input_size, hidden_size) * 0.01)
        self.b = Tensor(np.zeros((1, hidden_size)))
        self._bucket = _ParamBucket([self.Wh, self.Wx, self.b])

    def forward(self, inputs: Tensor, prev_hidden=None):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 2 or inputs.shape[1] != self.input_size:
            raise ValueError(f"shapes {inputs.shape} and {(self.input_size, self.hidden_size)} not aligned")
        B, H = inputs.shape[0], self.hidden_size
        if prev_hidden is None:
            prev_hidden = Tensor(np.zeros((B, H)))
        prev_hidden = _as_tensor(prev_hidden)
        if prev_hidden.shape != (B, H):
            raise ValueError(f"shapes {prev_hidden.shape} and {(H, H)} not aligned")
        ctx = get_ctx()
        b = self._bucket
        b.materialize()
        self.inputs, self.prev_hidden = inputs, prev_hidden
        zx = Tensor._empty((B, H), np.float64)
        zh = Tensor._empty((B, H), np.float64)
        out = Tensor._empty((B, H), np.float64)
        if B:
            check(lib.okl_dense_fwd(ctx.h, inputs._dptr(NCX), b.ptr(1), b.ptr(2), zx._ptr_raw(), B, self.input_size, H, ACT_NONE))
            check(lib.okl_dense_fwd(ctx.h, prev_hidden._dptr(NCX), b.ptr(0), None, zh._ptr_raw(), B, H, H, ACT_NONE))
            check(lib.okl_binary(ctx.h, 0, zx._ptr_raw(), zh._ptr_raw(), zx._ptr_raw(), zx.size, zh.size))
            check(lib.okl_act_fwd(ctx.h, ACT_TANH, zx._ptr_raw(), out._ptr_raw(), out.size))
        self.hidden = out
        return out

    def backward(self, dL_dh: Tensor, lr: float):
        dL_dh = _as_tensor(dL_dh)
        B, H = self.hidden.shape
        if dL_dh.shape != (B, H):
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dh.shape} {(B, H)}")
        ctx = get_ctx()
        b = self._bucket
        dz = Tensor._empty((B, H), np.float64)
        dx = Tensor._empty((B, self.input_size), np.float64)
        dprev = Tensor._empty((B, H), np.float64)
        if B:
            check(lib.okl_act2_bwd(ctx.h, _lib.ACT_TANH_OUT, self.hidden._dptr(NCX), dL_dh._dptr(NCX), dz._ptr_raw(), dz.size, 0.0, 1.0))
            check(lib.okl_dense_bwd(ctx.h, self.prev_hidden._dptr(NCX), b.ptr(0), dz._ptr_raw(), dprev._ptr_raw(), b.grad_ptr(0), None,
                                    B, H, H, None))
            check(lib.okl_dense_bwd(ctx.h, self.inputs._dptr(NCX), b.ptr(1), dz._ptr_raw(), dx._ptr_raw(), b.grad_ptr(1), b.grad_ptr(2),
                                    B, self.input_size, H, None))
            b.step(0.0, False)                       # accumulate only (okl.py:1669-1674 never touch the weights)
        _set_input_grad(self.prev_hidden, dprev)
        _set_input_grad(self.inputs, dx)
        return dx, dprev

    def get_params(self):
        return {'Wh': self.Wh.data, 'Wx': self.Wx.data, 'b': self.b.data}

    def set_params(self, params):
        self.Wh.data = params['Wh']
        self.Wx.data = params['Wx']
        self.b.data = params['b']


class LSTMLayer:
    """okl.py:1412-1534: one LSTM step.  Gates f, i, o = sigmoid, c~ = tanh of [x | h_prev] W_g + b_g - four
    ``okl_dense_fwd`` with the activation fused; cell / hidden by one pointwise kernel.  backward(dL_dh, dL_dc, lr): the gate
    gradients by one pointwise kernel, then four ``okl_dense_bwd`` (weight, bias and [x | h] gradients); as shipped it ACCUMULATES
    the eight parameter gradients without ever applying them (bucket update with lr = 0) and returns (dL/dh_prev, dL/dc_prev)
    only, adding them to ``prev_hidden.grad`` / ``prev_cell.grad``."""

    def __init__(self, input_size, hidden_size):
        self.input_size, self.hidden_size = input_size, hidden_size
        shape = (input_size + hidden_size, hidden_size)
        self.Wf, self.Wi, self.Wc, self.Wo = (Tensor(np.random.randn(*shape) * 0.01) for _ in range(4))
        self.bf, self.bi, self.bc, self.bo = (Tensor(np.zeros((1, hidden_size))) for _ in range(4))
        self._bucket = _ParamBucket([self.Wf, self.Wi, self.Wc, self.Wo, self.bf, self.bi, self.bc, self.bo])

    def sigmoid(self, x: Tensor):
        x = _as_tensor(x)
        out = Tensor._empty(x.shape, _float_dtype(x))
        check(lib.okl_act_fwd(get_ctx().h, ACT_SIGMOID, x._dptr(NCX), out._ptr_raw(), x.size))
        return out

    def forward(self, inputs: Tensor, prev_hidden=None, prev_cell=None):
        inputs = _as_tensor(inputs)
        I, H = self.input_size, self.hidden_size
        if inputs.ndim != 2 or inputs.shape[1] != I:
            raise ValueError(f"shapes {(inputs.shape[0], inputs.shape[-1] + H)} and {(I + H, H)} not aligned")
        B = inputs.shape[0]
        prev_hidden = Tensor(np.zeros((B, H))) if prev_hidden is None else _as_tensor(prev_hidden)
        prev_cell = Tensor(np.zeros((B, H))) if prev_cell is None else _as_tensor(prev_cell)
        if prev_hidden.shape != (B, H) or prev_cell.shape != (B, H):
            raise ValueError(f"operands could not be broadcast together with shapes {prev_hidden.shape} {prev_cell.shape} {(B, H)}")
        ctx = get_ctx()
        b = self._bucket
        b.materialize()
        self.inputs, self.prev_hidden, self.prev_cell = inputs, prev_hidden, prev_cell
        comb = Tensor._empty((B, I + H), np.float64)
        gates = [Tensor._empty((B, H), np.float64) for _ in range(4)]
        cell, hidden = Tensor._empty((B, H), np.float64), Tensor._empty((B, H), np.float64)
        if B:
            check(lib.okl_concat_cols(ctx.h, inputs._dptr(NCX), prev_hidden._dptr(NCX), comb._ptr_raw(), B, I, H))
            for gi, act in enumerate((ACT_SIGMOID, ACT_SIGMOID, ACT_TANH, ACT_SIGMOID)):            # f, i, c~, o
                check(lib.okl_dense_fwd(ctx.h, comb._ptr_raw(), b.ptr(gi), b.ptr(4 + gi), gates[gi]._ptr_raw(), B, I + H, H, act))
            check(lib.okl_lstm_pointwise_fwd(ctx.h, gates[0]._ptr_raw(), gates[1]._ptr_raw(), gates[2]._ptr_raw(), gates[3]._ptr_raw(),
                                             prev_cell._dptr(NCX), cell._ptr_raw(), hidden._ptr_raw(), B * H))
        self._comb = comb
        self.f, self.i, self.c_candidate, self.o = gates
        self.cell, self.hidden = cell, hidden
        return hidden, cell

    def backward(self, dL_dh: Tensor, dL_dc: Tensor, lr: float):
        dL_dh, dL_dc = _as_tensor(dL_dh), _as_tensor(dL_dc)
        I, H = self.input_size, self.hidden_size
        B = self.hidden.shape[0]
        if dL_dh.shape != (B, H) or dL_dc.shape != (B, H):
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dh.shape} {dL_dc.shape} {(B, H)}")
        ctx = get_ctx()
        b = self._bucket
        dz = [Tensor._empty((B, H), np.float64) for _ in range(4)]
        dcomb = [Tensor._empty((B, I + H), np.float64) for _ in range(4)]
        dprev_c, dprev_h = Tensor._empty((B, H), np.float64), Tensor._empty((B, H), np.float64)
        if B:
            check(lib.okl_lstm_pointwise_bwd(ctx.h, self.f._ptr_raw(), self.i._ptr_raw(), self.c_candidate._ptr_raw(), self.o._ptr_raw(),
                                             self.prev_cell._dptr(NCX), self.cell._ptr_raw(), dL_dh._dptr(NCX), dL_dc._dptr(NCX),
                                             dz[0]._ptr_raw(), dz[1]._ptr_raw(), dz[2]._ptr_raw(), dz[3]._ptr_raw(), dprev_c._ptr_raw(), B * H))
            for gi in range(4):
                check(lib.okl_dense_bwd(ctx.h, self._comb._ptr_raw(), b.ptr(gi), dz[gi]._ptr_raw(), dcomb[gi]._ptr_raw(), b.grad_ptr(gi),
                                        b.grad_ptr(4 + gi), B, I + H, H, None))
            check(lib.okl_sum4_cols(ctx.h, dcomb[0]._ptr_raw(), dcomb[1]._ptr_raw(), dcomb[2]._ptr_raw(), dcomb[3]._ptr_raw(),
                                    dprev_h._ptr_raw(), B, I + H, I, H))
            b.step(0.0, False)                       # accumulate only: okl.py:1493-1504 never touch the weights
        _set_input_grad(self.prev_hidden, dprev_h)
        _set_input_grad(self.prev_cell, dprev_c)
        return dprev_h, dprev_c

    def get_params(self):
        return {'Wf': self.Wf.data, 'Wi': self.Wi.data, 'Wc': self.Wc.data, 'Wo': self.Wo.data,
                'bf': self.bf.data, 'bi': self.bi.data, 'bc': self.bc.data, 'bo': self.bo.data}

    def set_params(self, params):
        for k in ('Wf', 'Wi', 'Wc', 'Wo', 'bf', 'bi', 'bc', 'bo'):
            getattr(self, k).data = params[k]


class GRULayer:
    """okl.py:1537-1643: one GRU step.  z, r = sigmoid([x | h_prev] W + b), h~ = tanh([x | r h_prev] Wh + bh),
    hidden = (1 - z) h_prev + z h~.  Three ``okl_dense_fwd`` (activations fused) and three ``okl_dense_bwd`` plus pointwise
    kernels; backward accumulates the six parameter gradients without applying them (as shipped) and returns (dL/dx, dL/dh_prev)."""

    def __init__(self, input_size, hidden_size):
        self.input_size, self.hidden_size = input_size, hidden_size
        shape = (input_size + hidden_size, hidden_size)
        self.Wz, self.Wr, self.Wh = (Tensor(np.random.randn(*shape) * 0.01) for _ in range(3))
        self.bz, self.br, self.bh = (Tensor(np.zeros((1, hidden_size))) for _ in range(3))
        self._bucket = _ParamBucket([self.Wz, self.Wr, self.Wh, self.bz, self.br, self.bh])

    sigmoid = LSTMLayer.sigmoid

    def forward(self, inputs: Tensor, prev_hidden=None):
        inputs = _as_tensor(inputs)
        I, H = self.input_size, self.hidden_size
        if inputs.ndim != 2 or inputs.shape[1] != I:
            raise ValueError(f"shapes {(inputs.shape[0], inputs.shape[-1] + H)} and {(I + H, H)} not aligned")
        B = inputs.shape[0]
        prev_hidden = Tensor(np.zeros((B, H))) if prev_hidden is None else _as_tensor(prev_hidden)
        if prev_hidden.shape != (B, H):
            raise ValueError(f"operands could not be broadcast together with shapes {prev_hidden.shape} {(B, H)}")
        ctx = get_ctx()
        b = self._bucket
        b.materialize()
        self.inputs, self.prev_hidden = inputs, prev_hidden
        comb, comb2 = Tensor._empty((B, I + H), np.float64), Tensor._empty((B, I + H), np.float64)
        z, r, rh, hc, hidden = (Tensor._empty((B, H), np.float64) for _ in range(5))
        if B:
            xp, hp = inputs._dptr(NCX), prev_hidden._dptr(NCX)
            check(lib.okl_concat_cols(ctx.h, xp, hp, comb._ptr_raw(), B, I, H))
            check(lib.okl_dense_fwd(ctx.h, comb._ptr_raw(), b.ptr(0), b.ptr(3), z._ptr_raw(), B, I + H, H, ACT_SIGMOID))
            check(lib.okl_dense_fwd(ctx.h, comb._ptr_raw(), b.ptr(1), b.ptr(4), r._ptr_raw(), B, I + H, H, ACT_SIGMOID))
            check(lib.okl_binary(ctx.h, 1, r._ptr_raw(), hp, rh._ptr_raw(), B * H, B * H))
            check(lib.okl_concat_cols(ctx.h, xp, rh._ptr_raw(), comb2._ptr_raw(), B, I, H))
            check(lib.okl_dense_fwd(ctx.h, comb2._ptr_raw(), b.ptr(2), b.ptr(5), hc._ptr_raw(), B, I + H, H, ACT_TANH))
            check(lib.okl_gru_pointwise(ctx.h, 0, z._ptr_raw(), hc._ptr_raw(), hp, None, hidden._ptr_raw(), None, B * H))
        self._comb, self._comb2 = comb, comb2
        self.z, self.r, self.h_candidate, self.hidden = z, r, hc, hidden
        return hidden

    def backward(self, dL_dh: Tensor, lr: float):
        dL_dh = _as_tensor(dL_dh)
        I, H = self.input_size, self.hidden_size
        B = self.hidden.shape[0]
        if dL_dh.shape != (B, H):
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dh.shape} {(B, H)}")
        ctx = get_ctx()
        b = self._bucket
        dzh, dzz, dzr, dprev = (Tensor._empty((B, H), np.float64) for _ in range(4))
        dcz, dcr, dch = (Tensor._empty((B, I + H), np.float64) for _ in range(3))
        dx = Tensor._empty((B, I), np.float64)
        if B:
            hp, dhp = self.prev_hidden._dptr(NCX), dL_dh._dptr(NCX)
            check(lib.okl_gru_pointwise(ctx.h, 1, self.z._ptr_raw(), self.h_candidate._ptr_raw(), hp, dhp, dzh._ptr_raw(), dzz._ptr_raw(), B * H))
            check(lib.okl_dense_bwd(ctx.h, self._comb2._ptr_raw(), b.ptr(2), dzh._ptr_raw(), dch._ptr_raw(), b.grad_ptr(2), b.grad_ptr(5),
                                    B, I + H, H, None))
            check(lib.okl_gru_dr(ctx.h, dch._ptr_raw(), hp, self.r._ptr_raw(), dzr._ptr_raw(), B, I, H))
            check(lib.okl_dense_bwd(ctx.h, self._comb._ptr_raw(), b.ptr(0), dzz._ptr_raw(), dcz._ptr_raw(), b.grad_ptr(0), b.grad_ptr(3),
                                    B, I + H, H, None))
            check(lib.okl_dense_bwd(ctx.h, self._comb._ptr_raw(), b.ptr(1), dzr._ptr_raw(), dcr._ptr_raw(), b.grad_ptr(1), b.grad_ptr(4),
                                    B, I + H, H, None))
            check(lib.okl_gru_dinputs(ctx.h, dcz._ptr_raw(), dcr._ptr_raw(), dch._ptr_raw(), self.r._ptr_raw(), self.z._ptr_raw(), dhp,
                                      dx._ptr_raw(), dprev._ptr_raw(), B, I, H))
            b.step(0.0, False)                       # accumulate only (okl.py:1590-1595 never touch the weights)
        _set_input_grad(self.prev_hidden, dprev)
        _set_input_grad(self.inputs, dx)
        return dx, dprev

    def get_params(self):
        return {'Wz': self.Wz.data, 'Wr': self.Wr.data, 'Wh': self.Wh.data, 'bz': self.bz.data, 'br': self.br.data, 'bh': self.bh.data}

    def set_params(self, params):
        for k in ('Wz', 'Wr', 'Wh', 'bz', 'br', 'bh'):
            getattr(self, k).data = params[k]


# ====================================================================================================== batch normalisation
class BatchNormLayer:
    """okl.py:1145-1222: (B, F) inputs; training mode normalises with the batch mean / population variance and
    updates the running statistics, eval mode uses them.  As shipped: ``var = max(var, eps)`` before ``+ eps``; the backward's
    chain rule uses the RUNNING variance while x_norm uses the batch one; gamma / beta gradients accumulate forever and the update
    ``p -= lr * p.grad`` happens inside backward (same rule as the other layers, so the same bucket machinery - one all-reduce per
    step under data parallelism; statistics stay per replica like the reference's)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.gamma = Tensor(np.ones((1, num_features)))
        self.beta = Tensor(np.zeros((1, num_features)))
        self._bucket = _ParamBucket([self.gamma, self.beta])
        self._running = None                          # device [mean | var], created on first use
        self._running_host = (np.zeros((1, num_features)), np.ones((1, num_features)))
        self.training = True

    # running statistics live on the device; the attributes mirror the reference's ndarrays
    def _running_buf(self):
        if self._running is None:
            ctx = get_ctx()
            F = self.num_features
            self._running = _DevBuf(ctx, 2 * F * 4)
            h = np.ascontiguousarray(np.concatenate([np.asarray(self._running_host[0], dtype=np.float64).reshape(-1),
                                                     np.asarray(self._running_host[1], dtype=np.float64).reshape(-1)]))
            if h.size != 2 * F:
                raise ValueError("running statistics have the wrong number of features")
            check(lib.okl_h2d(ctx.h, self._running.ptr, h.ctypes.data, h.size, _lib.F64))
        return self._running

    def _read_running(self):
        if self._running is None:
            return self._running_host
        ctx = get_ctx()
        F = self.num_features
        out = np.empty(2 * F, dtype=np.float64)
        check(lib.okl_d2h(ctx.h, out.ctypes.data, self._running.ptr, out.size, _lib.F64))
        return out[:F].reshape(1, F), out[F:].reshape(1, F)

    @property
    def running_mean(self):
        return self._read_running()[0]

    @running_mean.setter
    def running_mean(self, v):
        self._running_host = (np.asarray(v, dtype=np.float64), self._read_running()[1])
        self._running = None

    @property
    def running_var(self):
        return self._read_running()[1]

    @running_var.setter
    def running_var(self, v):
        self._running_host = (self._read_running()[0], np.asarray(v, dtype=np.float64))
        self._running = None

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 2:
            raise ValueError(f"too many values to unpack (expected 2)" if inputs.ndim > 2 else "not enough values to unpack (expected 2, got 1)")
        B, F = inputs.shape
        if F != self.num_features:
            raise ValueError(f"operands could not be broadcast together with shapes {inputs.shape} (1,{self.num_features})")
        ctx = get_ctx()
        self._bucket.materialize()
        self.inputs, self.batch_size = inputs, B
        self._mv = _DevBuf(ctx, 2 * F * 4)
        out = Tensor._empty((B, F), np.float64)
        if B:
            check(lib.okl_bn_fwd(ctx.h, inputs._dptr(NCX), self._bucket.ptr(0), self._bucket.ptr(1), self._running_buf().ptr, self._mv.ptr,
                                 out._ptr_raw(), B, F, float(self.eps), float(self.momentum), 1 if self.training else 0))
        self.outputs = out
        return out

    def backward(self, dL_dout: Tensor, lr: float):
        dL_dout = _as_tensor(dL_dout)
        B, F = self.inputs.shape
        if dL_dout.shape != (B, F):
            raise ValueError(f"operands could not be broadcast together with shapes {dL_dout.shape} {(B, F)}")
        ctx = get_ctx()
        b = self._bucket
        dx = Tensor._empty((B, F), np.float64)
        check(lib.okl_bn_bwd(ctx.h, self.inputs._dptr(NCX), dL_dout._dptr(NCX), b.ptr(0), self._running_buf().ptr, self._mv.ptr,
                             dx._ptr_raw(), b.grad_ptr(0), b.grad_ptr(1), B, F, float(self.eps)))
        b.step(lr, False)
        return dx

    def get_params(self):
        return {'gamma': self.gamma.data, 'beta': self.beta.data, 'running_mean': self.running_mean, 'running_var': self.running_var}

    def set_params(self, params):
        self.gamma.data = params['gamma']
        self.beta.data = params['beta']
        self.running_mean = params['running_mean']
        self.running_var = params['running_var']


class _RegularizationLayer:
    """okl.py:1052-1142: wraps a layer that has ``gamma`` / ``beta`` (BatchNormLayer).  After the wrapped layer's backward the
    penalty gradient is added to the accumulated gradients AND applied to the parameters once more (as shipped).  The vectors are
    a few hundred floats: the arithmetic runs on the lazy host mirrors, exactly the reference's NumPy expressions."""

    def __init__(self, layer, lambda_):
        self.layer, self.lambda_ = layer, lambda_

    def forward(self, inputs):
        return self.layer.forward(inputs)

    def _penalty_grad(self, p):
        raise NotImplementedError

    def backward(self, dL_dout, lr):
        dL_dx = self.layer.backward(dL_dout, lr)
        for t in (self.layer.gamma, self.layer.beta):
            pg = self._penalty_grad(np.asarray(t.data))
            t.grad = np.asarray(t.grad) + pg.reshape(np.asarray(t.grad).shape)
            t.data = np.asarray(t.data) - lr * pg.reshape(np.asarray(t.data).shape)
        return dL_dx

    def get_params(self):
        return self.layer.get_params()

    def set_params(self, params):
        self.layer.set_params(params)


class L1RegularizationLayer(_RegularizationLayer):
    def _penalty_grad(self, p):
        return self.lambda_ * np.sign(p)                    # okl.py:1070-1071


class L2RegularizationLayer(_RegularizationLayer):
    def _penalty_grad(self, p):
        return 2 * self.lambda_ * p                         # okl.py:1099-1100


class L3RegularizationLayer(_RegularizationLayer):
    def _penalty_grad(self, p):
        return 3 * self.lambda_ * p ** 2                    # okl.py:1126-1127


class DropoutLayer:
    """okl.py:1336-1377.  The mask is drawn on the HOST with the reference's own call (np.random.binomial on the global NumPy
    stream, scaled by 1 / (1 - rate)), so a seeded run drops exactly the units the reference drops; it is uploaded once and the
    products run on the device."""

    def __init__(self, rate=0.5):
        self.rate = rate

    def forward(self, inputs: Tensor, training=True):
        inputs = _as_tensor(inputs)
        self.inputs = inputs
        if not training:
            return inputs
        self.mask = np.random.binomial(1, 1 - self.rate, size=inputs.shape) / (1 - self.rate)
        self._mask_dev = Tensor(self.mask, requires_grad=False)
        self.outputs = inputs * self._mask_dev
        return self.outputs

    def backward(self, dL_dout: Tensor, lr: float):
        dx = _as_tensor(dL_dout) * self._mask_dev
        _set_input_grad(self.inputs, dx)
        return dx

    def get_params(self):
        return None

    def set_params(self, params):
        pass


# ====================================================================================================== pooling / flatten
class _Pool:
    _kind = POOL_MAX
    _fused_relu_bwd = False      # NeuralNetwork planner: the ReLU right before this layer gets its backward applied here

    def __init__(self, pool_size=2, stride=2):
        self.pool_size = pool_size
        self.stride = stride

    _in_fused_block = False      # NeuralNetwork planner: computed inside the preceding conv's fused kernel (pure pass-through)
    _db_sink = None              # NeuralNetwork planner: conv layer whose bias gradient = channel sums of this layer's input gradient

    def forward(self, inputs):
        if self._in_fused_block:
            return inputs                      # already pooled; .inputs / .output were set by the conv layer
        inputs = _as_tensor(inputs)
        ctx = get_ctx()
        if inputs.ndim != 4:
            raise ValueError(f"not enough values to unpack (expected 4, got {inputs.ndim})")      # okl.py:1732
        N, Cc, H, W = inputs.shape
        ps, st = int(self.pool_size), int(self.stride)
        if H < ps or W < ps:
            raise ValueError("pooling window larger than the input")
        OH, OW = (H - ps) // st + 1, (W - ps) // st + 1
        self._b16_path = bool(self._kind == POOL_MAX and ps == 2 and st == 2 and inputs._b16_primary and inputs._layout == NXC and
                              lib.okl_pool2_bf16_supported(Cc, H, W))
        if self._b16_path:                     # channels-last bf16 in, channels-last bf16 out (BF16 tensor-core pipeline)
            out = Tensor._empty_b16((N, Cc, OH, OW), np.float64, NXC)
            check(lib.okl_pool2_bf16_fwd(ctx.h, inputs._b16(NXC), out._bf16.ptr, N, Cc, H, W))
            self.inputs, self.output = inputs, out
            return out
        src = inputs._dptr()
        lay = inputs._layout
        out = Tensor._empty((N, Cc, OH, OW), np.float64, layout=lay)
        check(lib.okl_pool_fwd(ctx.h, self._kind, src, out._ptr_raw(), N, Cc, H, W, ps, st, lay))
        self.inputs = inputs
        self.output = out
        return out

    def backward(self, dL_dout: Tensor, lr: float = None):
        if self._in_fused_block:
            return dL_dout                     # the conv layer consumes the pooled gradient directly
        dL_dout = _as_tensor(dL_dout)
        ctx = get_ctx()
        x = self.inputs
        N, Cc, H, W = x.shape
        if dL_dout.shape != self.output.shape:
            raise ValueError(f"dL_dout shape {dL_dout.shape} does not match the layer output {self.output.shape}")
        if getattr(self, "_b16_path", False) and x._b16_primary:
            dx = Tensor._empty_b16(x.shape, np.float64, NXC)
            sink = self._db_sink if (self._fused_relu_bwd and self._db_sink is not None and
                                     getattr(self._db_sink, "_conv2", False) and self._db_sink.out_channels == Cc) else None
            check(lib.okl_pool2_bf16_bwd(ctx.h, x._b16(NXC), dL_dout._b16(NXC), dx._bf16.ptr, N, Cc, H, W, 1 if self._fused_relu_bwd else 0,
                                         sink._bucket.grad_ptr(1) if sink is not None else None,
                                         int(sink._aux_lane) if (sink is not None and sink._aux_lane and sink._defer_update) else 0))
            if sink is not None:
                dx._db_done = sink                 # the convolution in front already has its bias gradient: this kernel reduced it
            _set_input_grad(x, dx)
            return dx
        xp = x._dptr()
        lay = x._layout
        dx = Tensor._empty(x.shape, np.float64, layout=lay)
        check(lib.okl_pool_bwd(ctx.h, self._kind, xp, dL_dout._dptr(lay), dx._ptr_raw(), N, Cc, H, W, int(self.pool_size),
                               int(self.stride), lay, 1 if self._fused_relu_bwd else 0))
        _set_input_grad(x, dx)                         # okl.py:1768 / 1821
        return dx

    def get_params(self):
        return None

    def set_params(self, params):
        pass


class MaxPoolingLayer(_Pool):
    """okl.py:1715-1778.  Window max on NCHW, no padding, float64 output.  Backward (R4): every position equal to its
    window's max receives that window's gradient (ties -> all maxima; overlapping windows overwrite in raster order)."""
    _kind = POOL_MAX


class AveragePoolingLayer(_Pool):
    """okl.py:1781-1831.  Window mean; backward (R5): dX[window] += g / pool^2."""
    _kind = POOL_AVG


class _Unpool:
    """okl.py:2619-2705: every input pixel stamps a pool x pool block at (i * stride, j * stride) by plain assignment (overlaps:
    the last stamp in (i, j) order wins; gaps stay 0); backward sums the block.  Output float64."""
    _kind = 0

    def __init__(self, pool_size=2, stride=2):
        self.pool_size, self.stride = pool_size, stride

    def forward(self, inputs):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 4:
            raise ValueError(f"not enough values to unpack (expected 4, got {inputs.ndim})")
        N, Cc, H, W = inputs.shape
        ps, st = int(self.pool_size), int(self.stride)
        ctx = get_ctx()
        out = Tensor._empty((N, Cc, (H - 1) * st + ps, (W - 1) * st + ps), np.float64)
        if out.size:
            check(lib.okl_unpool_fwd(ctx.h, self._kind, inputs._dptr(NCX), out._ptr_raw(), N * Cc, H, W, ps, st))
        self.inputs, self.output = inputs, out
        return out

    def backward(self, dL_dout: Tensor, lr: float = None):
        dL_dout = _as_tensor(dL_dout)
        N, Cc, H, W = self.inputs.shape
        ps, st = int(self.pool_size), int(self.stride)
        if dL_dout.shape != (N, Cc, (H - 1) * st + ps, (W - 1) * st + ps):
            raise ValueError(f"could not broadcast input array from shape {dL_dout.shape} into the unpooled extent")
        ctx = get_ctx()
        dx = Tensor._empty(self.inputs.shape, _float_dtype(self.inputs))
        if dx.size:
            check(lib.okl_unpool_bwd(ctx.h, self._kind, dL_dout._dptr(NCX), dx._ptr_raw(), N * Cc, H, W, ps, st))
        _set_input_grad(self.inputs, dx)
        return dx

    def get_params(self):
        return None

    def set_params(self, params):
        pass


class MaxUnpoolingLayer(_Unpool):
    _kind = 0


class AverageUnpoolingLayer(_Unpool):
    _kind = 1


class PaddingLayer:
    """okl.py:1920-1943: int, (pad_h, pad_w) or (top, bottom, left, right)."""

    def __init__(self, padding):
        if isinstance(padding, int):
            self.padding = ((padding, padding), (padding, padding))
        elif len(padding) == 2:
            self.padding = ((padding[0], padding[0]), (padding[1], padding[1]))
        elif len(padding) == 4:
            self.padding = ((padding[0], padding[1]), (padding[2], padding[3]))
        else:
            raise ValueError("Invalid padding format. Use int, (pad_h, pad_w) or (pad_top, pad_bottom, pad_left, pad_right)")

    def forward(self, inputs: Tensor) -> Tensor:
        raise NotImplementedError

    def backward(self, dL_dout: Tensor, lr: float = None) -> Tensor:
        raise NotImplementedError

    def get_params(self):
        return None

    def set_params(self, params):
        pass

    _reflect = 0

    def _pad(self, inputs):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 4:
            raise ValueError("operands could not be broadcast together with remapped shapes [original->remapped]")   # np.pad pad_width
        (pt, pb), (pl, pr) = self.padding
        N, Cc, H, W = inputs.shape
        if self._reflect and ((max(pt, pb) > 0 and H < 2) or (max(pl, pr) > 0 and W < 2)):
            raise ValueError("can't extend empty axis / single element using modes other than 'constant'")
        ctx = get_ctx()
        out = Tensor._empty((N, Cc, H + pt + pb, W + pl + pr), _float_dtype(inputs))
        if out.size:
            check(lib.okl_pad2d(ctx.h, inputs._dptr(NCX), out._ptr_raw(), N * Cc, H, W, pt, pb, pl, pr, self._reflect))
        self.inputs = inputs
        return out

    def _crop(self, dL_dout):
        dL_dout = _as_tensor(dL_dout)
        (pt, pb), (pl, pr) = self.padding
        N, Cc, GH, GW = dL_dout.shape
        # okl.py:1955 slices pad_top:-pad_bottom - an empty slice when the bottom / right pad is 0 (as shipped)
        h = max(GH - pb - pt, 0) if pb > 0 else 0
        w = max(GW - pr - pl, 0) if pr > 0 else 0
        ctx = get_ctx()
        dx = Tensor._empty((N, Cc, h, w), _float_dtype(dL_dout))
        if dx.size:
            check(lib.okl_crop2d(ctx.h, dL_dout._dptr(NCX), dx._ptr_raw(), N * Cc, GH, GW, pt, pl, h, w))
        if dx.shape == self.inputs.shape:
            _set_input_grad(self.inputs, dx)
        elif self.inputs.requires_grad and self.inputs._grad is not None:
            raise ValueError(f"operands could not be broadcast together with shapes {self.inputs.shape} {dx.shape}")
        elif self.inputs.requires_grad:
            self.inputs._grad = dx
        return dx


class ZeroPaddingLayer(PaddingLayer):
    """okl.py:1946-1962."""

    def forward(self, inputs: Tensor) -> Tensor:
        return self._pad(inputs)

    def backward(self, dL_dout: Tensor, lr: float = None) -> Tensor:
        return self._crop(dL_dout)


class ReflectionPaddingLayer(PaddingLayer):
    """okl.py:1965-1981: np.pad mode 'reflect'; the backward is the plain crop (as shipped, not the adjoint of the reflection)."""
    _reflect = 1

    def forward(self, inputs: Tensor) -> Tensor:
        return self._pad(inputs)

    def backward(self, dL_dout: Tensor, lr: float = None) -> Tensor:
        return self._crop(dL_dout)


class FlattenLayer:
    """okl.py:639-651: (N, ...) -> (N, -1) and back; a zero-copy view of the device block."""

    _lazy_for_head = False       # NeuralNetwork planner: the fused classifier head reads the channels-last bf16 source directly

    def __init__(self):
        self.inputs_shape = None

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        self.inputs_shape = inputs.shape
        if self._lazy_for_head and inputs._b16_primary and inputs._layout == NXC and inputs.ndim >= 3:
            # nothing is moved: the flat (N, C * S) tensor is materialised (bf16 -> fp32, channels-last -> reference order) only if
            # somebody other than the fused head asks for it
            out = Tensor._view(None, 0, (inputs.shape[0], _numel(inputs.shape[1:])), inputs._np_dtype, bound=False)
            out._dev_valid = False
            out._flat_src = inputs

            def materialize(t, src=inputs):
                flat = src.reshape((src.shape[0], -1))
                t._buf, t._off, t._dev_ver, t._dev_valid = flat._buf, flat._off, flat._dev_ver, True
            out._lazy_fn = materialize
            return out
        return inputs.reshape((inputs.shape[0], -1))

    def backward(self, dL_dout: Tensor, lr: float):
        dL_dout = _as_tensor(dL_dout)
        if dL_dout.shape == tuple(self.inputs_shape):          # the fused head already produced the gradient in the source's shape
            return dL_dout
        return dL_dout.reshape(self.inputs_shape)

    def get_params(self):           # the reference lacks this (NeuralNetwork.save breaks on CNNs, okl.py:3022); harmless to add
        return None

    def set_params(self, params):
        pass


class Unfold:
    """okl.py:568-636: im2col as a standalone layer.  forward: (N, C, H, W) -> (N, C*k*k, OH*OW) float64, row c*k*k + u*k + v,
    column i*OW + j; with padding > 0 a border window is clipped and copied to the top-left of a zero patch (as shipped, not true
    zero padding).  backward: col2im scatter-add; with padding > 0 the reference's clipped slice cannot take the k x k block and
    NumPy raises ValueError - so does this."""

    def __init__(self, kernel_size, stride=1, padding=0):
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.input_shape = None

    def _extent(self, H, W):
        k, s, p = int(self.kernel_size), int(self.stride), int(self.padding)
        return k, s, p, (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 4:
            raise ValueError(f"not enough values to unpack (expected 4, got {inputs.ndim})")          # okl.py:577
        self.input_shape = inputs.shape
        N, Cc, H, W = inputs.shape
        k, s, p, OH, OW = self._extent(H, W)
        if OH < 0 or OW < 0:
            raise ValueError("negative dimensions are not allowed")
        ctx = get_ctx()
        out = Tensor._empty((N, Cc * k * k, max(OH, 0) * max(OW, 0)), np.float64)
        if out.size:
            check(lib.okl_unfold_fwd(ctx.h, inputs._dptr(NCX), out._ptr_raw(), N, Cc, H, W, k, s, p))
        return out

    def backward(self, dL_dout: Tensor):
        dL_dout = _as_tensor(dL_dout)
        N, Cc, H, W = self.input_shape
        k, s, p, OH, OW = self._extent(H, W)
        if dL_dout.ndim != 3:
            raise ValueError(f"not enough values to unpack (expected 3, got {dL_dout.ndim})")         # okl.py:615
        if dL_dout.size != N * Cc * k * k * OH * OW:
            raise ValueError(f"cannot reshape array of size {dL_dout.size} into shape {(N, Cc, k, k, OH, OW)}")
        if p > 0 and OH * OW > 0:
            # okl.py:633: the clipped border slice is smaller than the (N, C, k, k) block added to it
            raise ValueError("operands could not be broadcast together: Unfold.backward with padding > 0 adds a k x k block to a "
                             "clipped border slice (okl.py:633 raises the same way)")
        ctx = get_ctx()
        dx = Tensor._empty(self.input_shape, np.float64)
        if dx.size:
            check(lib.okl_unfold_bwd(ctx.h, dL_dout._dptr(NCX), dx._ptr_raw(), N, Cc, H, W, k, s))
        return dx


class Fold:
    """okl.py:525-565: NOT col2im - a pure reshape / transposition (N, C*k*k, OH*OW) -> (N, C, k, k, OH, OW) -> transpose
    (0, 1, 4, 2, 5, 3) -> (N, -1, H, W), which equals col2im only for stride == kernel, padding == 0; backward is its inverse."""

    def __init__(self, output_size, kernel_size, stride=1, padding=0):
        self.output_size, self.kernel_size, self.stride, self.padding = output_size, kernel_size, stride, padding
        self.input_shape = None

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        if inputs.ndim != 3:
            raise ValueError(f"not enough values to unpack (expected 3, got {inputs.ndim})")
        self.input_shape = inputs.shape
        N, nch, L = inputs.shape
        H, W = self.output_size
        k, s, p = int(self.kernel_size), int(self.stride), int(self.padding)
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        Cc = nch // (k * k)
        if Cc * k * k * OH * OW != nch * L or OH < 0 or OW < 0:
            raise ValueError(f"cannot reshape array of size {inputs.size} into shape {(N, Cc, k, k, OH, OW)}")
        if H * W == 0 or (Cc * k * k * OH * OW) % (H * W) != 0:
            raise ValueError(f"cannot reshape array of size {inputs.size} into shape ({N},-1,{H},{W})")
        ctx = get_ctx()
        out = Tensor._empty((N, Cc * k * k * OH * OW // (H * W), H, W), _float_dtype(inputs))
        if out.size:
            check(lib.okl_fold_permute(ctx.h, inputs._dptr(NCX), out._ptr_raw(), N * Cc, k, OH, OW, 0))
        return out

    def backward(self, dL_dout: Tensor):
        dL_dout = _as_tensor(dL_dout)
        if dL_dout.ndim != 4:
            raise ValueError(f"not enough values to unpack (expected 4, got {dL_dout.ndim})")
        N, _, H, W = dL_dout.shape
        k, s, p = int(self.kernel_size), int(self.stride), int(self.padding)
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        blk = OH * k * OW * k
        if blk <= 0 or (dL_dout.size // max(N, 1)) % blk != 0 or dL_dout.size != _numel(self.input_shape):
            raise ValueError(f"cannot reshape array of size {dL_dout.size} into shape {self.input_shape}")
        ctx = get_ctx()
        out = Tensor._empty(self.input_shape, _float_dtype(dL_dout))
        if out.size:
            check(lib.okl_fold_permute(ctx.h, dL_dout._dptr(NCX), out._ptr_raw(), dL_dout.size // blk, k, OH, OW, 1))
        return out


class UnflattenLayer:
    """okl.py:654-666."""

    def __init__(self, output_shape):
        self.output_shape = tuple(output_shape)

    def forward(self, inputs: Tensor):
        inputs = _as_tensor(inputs)
        return inputs.reshape((inputs.shape[0],) + self.output_shape)

    def backward(self, dL_dout: Tensor, lr: float):
        dL_dout = _as_tensor(dL_dout)
        return dL_dout.reshape((dL_dout.shape[0], -1))

    def get_params(self):
        return None

    def set_params(self, params):
        pass
