"""CrossEntropyLoss / MSELoss with the reference's call protocol (okl.py:1380-1409, 1833-1839): ``forward(outputs, targets)``
returns the loss value, ``backward(outputs, targets)`` returns dL/doutputs as a Tensor.  The value is computed on the device
together with the gradient (one kernel) and is returned as a ``LossValue`` - a float-like object that reads the device scalar
only when something actually looks at it, so a training loop does not synchronise with the host every batch.

Data parallel: the gradient is divided by the GLOBAL batch (local rows x ranks) so that the sum of the ranks' parameter
gradients equals the single-GPU gradient."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, get_ctx, NCX
from .tensor import Tensor, _DevBuf


class LossValue:
    """Float-like view of a device fp32 scalar (mean loss of one batch)."""
    __slots__ = ("_buf", "_val", "_epoch")

    def __init__(self, buf):
        self._buf, self._val, self._epoch = buf, None, -1

    def _get(self):
        ctx = get_ctx()
        if self._val is None or self._epoch != ctx.epoch:
            v = C.c_float()
            check(lib.okl_read_scalar(ctx.h, self._buf.ptr, C.byref(v)))
            self._val, self._epoch = np.float64(v.value), ctx.epoch
        return self._val

    @property
    def data(self):                       # okl.py:2973 does float(np.sum(loss.data))
        return np.array(self._get())

    def item(self):
        return float(self._get())

    def __float__(self):
        return float(self._get())

    def __repr__(self):
        return repr(self._get())

    def __format__(self, spec):
        return format(float(self._get()), spec)

    def __add__(self, o):
        return float(self) + o

    __radd__ = __add__

    def __sub__(self, o):
        return float(self) - o

    def __rsub__(self, o):
        return o - float(self)

    def __mul__(self, o):
        return float(self) * o

    __rmul__ = __mul__

    def __truediv__(self, o):
        return float(self) / o

    def __lt__(self, o):
        return float(self) < o

    def __le__(self, o):
        return float(self) <= o

    def __gt__(self, o):
        return float(self) > o

    def __ge__(self, o):
        return float(self) >= o

    def __eq__(self, o):
        return float(self) == o

    def __hash__(self):
        return hash(float(self))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._get(), dtype=dtype)


def _targets_i32(targets: Tensor, rows: int):
    """Class indices as an int32 device vector (okl.py:1392 does targets.data.astype(int))."""
    ctx = get_ctx()
    cached = getattr(targets, "_i32", None)
    if cached is not None and cached[1] == (targets._dev_ver, id(targets._host)):
        return cached[0]
    if targets.size != rows:
        raise ValueError(f"targets has {targets.size} entries for {rows} samples")
    buf = _DevBuf(ctx, rows * 4)
    host = np.ascontiguousarray(np.asarray(targets.data).reshape(-1).astype(np.int64))
    check(lib.okl_h2d_i32(ctx.h, buf.ptr, host.ctypes.data, rows, _lib.I64))
    targets._i32 = (buf, (targets._dev_ver, id(targets._host)))
    return buf


class CrossEntropyLoss:
    """okl.py:1380-1409: expects probabilities (after softmax) and integer class targets.
    forward: mean(-log(clip(p, 1e-12, 1-1e-12)[i, t_i]));  backward: (clip(p) - onehot) / samples.
    As in the reference, ``backward`` ignores its arguments and uses what ``forward`` stashed (okl.py:1399-1403)."""
    _fuse_softmax = None       # set by NeuralNetwork's fast path
    _loss_accum = None         # device scalar accumulating the epoch loss (set by NeuralNetwork.train)

    def forward(self, outputs: Tensor, targets: Tensor):
        ctx = get_ctx()
        if outputs.ndim != 2:
            raise ValueError(f"CrossEntropyLoss expects (batch, classes) probabilities, got shape {outputs.shape}")
        rows, cols = outputs.shape
        tbuf = targets._i32_dev if hasattr(targets, "_i32_dev") else _targets_i32(targets, rows)
        self._tbuf = tbuf
        loss = _DevBuf(ctx, 4)
        grad = Tensor._empty((rows, cols), np.float64)
        tptr = tbuf if isinstance(tbuf, int) else tbuf.ptr
        sm = self._fuse_softmax          # NeuralNetwork fast path: the softmax layer whose backward is computed right here
        dz = Tensor._empty((rows, cols), np.float64) if (sm is not None and sm.outputs is outputs) else None
        head = getattr(outputs, "_head_of", None)
        if dz is not None and head is not None and getattr(outputs, "_lazy_fn", None) is not None and head.outputs is outputs:
            # the probabilities have not been computed yet: Dense -> Softmax -> CE forward and backward down to the dense
            # layer's dX in one kernel
            x = head.inputs
            src = getattr(x, "_flat_src", None)
            if (getattr(head, "_big_head", False) and src is not None and getattr(x, "_lazy_fn", None) is not None and src._b16_primary and
                    src._layout == _lib.NXC and lib.okl_head_big_supported(rows, src.shape[1], x.shape[1] // src.shape[1], cols)):
                # Flatten -> Dense -> Softmax -> CE on a channels-last bf16 activation: forward, loss, every gradient (dX as bf16 in
                # the activation's own layout, dW, db) in four small kernels; nothing is permuted or widened
                outputs._lazy_fn = None
                Cc, S_ = src.shape[1], x.shape[1] // src.shape[1]
                need_dx = getattr(head, "_need_dx", True)
                dx = Tensor._empty_b16(src.shape, np.float64, _lib.NXC) if need_dx else None
                bk = head._bucket
                if getattr(head, "_wt_scratch", None) is None or head._wt_scratch.nbytes < cols * x.shape[1] * 2:
                    head._wt_scratch = _DevBuf(ctx, cols * x.shape[1] * 2)
                outputs._buf, outputs._off = _DevBuf(ctx, rows * cols * 4), 0
                check(lib.okl_head_big(ctx.h, src._b16(_lib.NXC), bk.ptr(0), bk.ptr(1), tptr, outputs._ptr_raw(), loss.ptr, grad._ptr_raw(),
                                       dz._ptr_raw(), dx._bf16.ptr if dx is not None else None, bk.grad_ptr(0), bk.grad_ptr(1),
                                       head._wt_scratch.ptr, rows, Cc, S_, cols, float(rows * ctx.nranks), self._loss_accum,
                                       int(head._aux_lane) if (head._aux_lane and head._defer_update) else 0))
                outputs._touch()
                head._head_pre = (dz, dx, True)
                sm._precomputed = (grad, dz)
                self.outputs, self.targets = outputs, targets
                self._grad = grad
                self._loss = LossValue(loss)
                return self._loss
            outputs._lazy_fn = None
            K = x.shape[1]
            dx = Tensor._empty((rows, K), np.float64) if getattr(head, "_need_dx", True) else None
            xp = x._dptr(NCX)
            check(lib.okl_dense_softmax_ce(ctx.h, xp, head._bucket.ptr(0), head._bucket.ptr(1), tptr, outputs._ptr_raw(), loss.ptr,
                                           grad._ptr_raw(), dz._ptr_raw(), dx._ptr_raw() if dx is not None else None,
                                           xp if (dx is not None and head._dx_relu_mask) else None, rows, K, cols,
                                           float(rows * ctx.nranks), self._loss_accum))
            outputs._touch()
            head._head_pre = (dz, dx)
            sm._precomputed = (grad, dz)
            self.outputs, self.targets = outputs, targets
            self._grad = grad
            self._loss = LossValue(loss)
            return self._loss
        check(lib.okl_ce_fwd_bwd(ctx.h, outputs._dptr(NCX), tptr, loss.ptr, grad._ptr_raw(), rows, cols, float(rows * ctx.nranks),
                                 self._loss_accum, dz._ptr_raw() if dz is not None else None))
        if dz is not None:
            sm._precomputed = (grad, dz)
        self.outputs, self.targets = outputs, targets
        self._grad = grad
        self._loss = LossValue(loss)
        return self._loss

    def backward(self, dL_dloss=None, lr=None):
        g = self._grad
        if self.outputs.requires_grad:                      # okl.py:1405
            self.outputs._grad = g if self.outputs._grad is None else (
                (self.outputs._grad if isinstance(self.outputs._grad, Tensor) else Tensor(self.outputs._grad)) + g)
        return g


class MSELoss:
    """okl.py:1833-1839: mean((o - t)^2); gradient 2 (o - t) / t.size."""
    _loss_accum = None
    _fuse_relu = None          # NeuralNetwork fast path: trailing ReLU layer whose backward is applied in the loss kernel

    def forward(self, outputs: Tensor, targets: Tensor):
        ctx = get_ctx()
        if outputs.shape != targets.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {outputs.shape} {targets.shape}")
        loss = _DevBuf(ctx, 4)
        grad = Tensor._empty(outputs.shape, np.float64)
        n = outputs.size
        relu = self._fuse_relu
        fuse = relu is not None and getattr(relu, "outputs", None) is outputs
        if (outputs.ndim >= 3 and outputs._dev_valid and outputs._layout == _lib.NXC and outputs._b16_primary and outputs.shape[1] % 8 == 0 and
                0 < outputs.shape[0] <= 65535 and n // (outputs.shape[0] * outputs.shape[1]) > 1 and
                outputs._bf16_ver == (outputs._dev_ver, _lib.NXC) and getattr(outputs, "_lazy_fn", None) is None):
            # bf16 channels-last outputs of a BF16 tensor-core convolution: loss and gradient straight from / to that format against
            # the caller's fp32 (N, C, ...) targets - no fp32 copy of the outputs, no fp32 gradient, nothing permuted
            N_, C_ = outputs.shape[0], outputs.shape[1]
            S_ = n // (N_ * C_)
            grad = Tensor._empty_b16(outputs.shape, np.float64, _lib.NXC)
            tr = targets._rows_pending() if isinstance(targets, Tensor) else None     # gather-free batch: dataset rows through the index
            tp, ti = tr if tr is not None else (targets._dptr(NCX), None)
            check(lib.okl_mse_fwd_bwd_nxc_bf16(ctx.h, outputs._bf16.ptr, tp, ti, loss.ptr, grad._bf16.ptr, N_, C_, S_,
                                               float(n * ctx.nranks), self._loss_accum, 1 if fuse else 0))
        elif (outputs.ndim >= 3 and outputs._dev_valid and outputs._layout == _lib.NXC and outputs.shape[1] > 1 and
                outputs.shape[0] <= 65535 and getattr(outputs, "_lazy_fn", None) is None):
            # channels-last outputs of a tensor-core convolution: loss and gradient are computed in that layout against the
            # caller's (N, C, ...) targets - neither tensor is permuted in memory
            N_, C_ = outputs.shape[0], outputs.shape[1]
            S_ = n // (N_ * C_)
            grad._layout = _lib.NXC
            check(lib.okl_mse_fwd_bwd_nxc(ctx.h, outputs._dptr(), targets._dptr(NCX), loss.ptr, grad._ptr_raw(), N_, C_, S_,
                                          float(n * ctx.nranks), self._loss_accum, 1 if fuse else 0))
        else:
            check(lib.okl_mse_fwd_bwd(ctx.h, outputs._dptr(NCX), targets._dptr(NCX), loss.ptr, grad._ptr_raw(), n, float(n * ctx.nranks),
                                      self._loss_accum, 1 if fuse else 0))
        if fuse:
            relu._precomputed = grad             # already g * (y > 0): the layer's backward passes it through
        self._grad = grad
        self._key = (id(outputs), id(targets))
        self._loss = LossValue(loss)
        return self._loss

    def backward(self, outputs: Tensor, targets: Tensor):
        if getattr(self, "_key", None) != (id(outputs), id(targets)):
            self.forward(outputs, targets)
        return self._grad
