"""CPU oracle for the okrolearn Conv/Dense train-step hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference's algorithm for the path named in
BASELINE.json ``north_star``.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.  Nothing under ``okrolearn_b200/`` imports it, and the product path has no
CPU fallback.

Pinning status: **pinned against the reference itself run in the build container**
(``tests/golden/make_golden.py`` imports ``/root/reference/src`` and asserts this module
reproduces its outputs, then commits the vectors under ``tests/golden/``).  Forward of every
layer, Dense backward + update, Softmax backward, CE/MSE fwd+bwd and the schedulers are pinned
by the shipped reference code.  Conv1D/2D/3D and Max/AvgPool *backward* raise ``ValueError``
in the shipped reference for every shape; for those this oracle follows the
reference loops with the five minimal repairs R1-R5 listed there, and is pinned against
``torch.autograd`` (CPU, fp64) instead.  The reference's own conv known-answer tests
(tests/conv2d.py:43-48, tests/conv1d.py:30-31) hold wrong expected arrays (the reference code
itself disagrees with them and matches torch), so they are recorded, not enforced.
The SURVEY 8(f) rows added later (Unfold / Fold, the other activations, BatchNorm, the regularisation wrappers, Dropout, un-pooling,
padding, RNN / LSTM / GRU cells, transposed-conv forward) are pinned the same way on the running reference; the transposed-conv
backward - which raises as shipped - on ``torch.autograd``.

Every function cites the reference file:line it follows.  "okl.py" = src/okrolearn/okrolearn.py.

Two flavours are provided for the conv/pool loops:
  * ``*_loops``  - faithful restatement (one Python iteration per output position, the same
                   broadcast-multiply + ``np.sum`` per position as the reference).  This is the
                   oracle proper and the timed CPU baseline ("port").
  * ``*_fast``   - the same contraction vectorised with ``sliding_window_view`` + ``einsum`` in the
                   same dtype; validated against ``*_loops`` in ``tests/test_oracle.py``; used
                   where the loops would take minutes.
"""
from __future__ import annotations

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

__all__ = [
    "dense_forward", "dense_backward",
    "conv_forward_loops", "conv_backward_loops", "conv_forward_fast", "conv_backward_fast",
    "relu_forward", "relu_forward_as_shipped", "relu_backward", "sigmoid_forward", "sigmoid_backward",
    "tanh_forward", "tanh_backward", "softmax_forward", "softmax_backward", "softmax_backward_jacobian",
    "maxpool_forward", "maxpool_backward", "avgpool_forward", "avgpool_backward",
    "cross_entropy_forward", "cross_entropy_backward", "mse_forward", "mse_backward",
    "accum_update", "sgd_update", "AdamState", "adam_update", "unfold_forward", "fold_forward",
    "OracleDense", "OracleConv", "OracleReLU", "OracleSigmoid", "OracleTanh", "OracleSoftmax",
    "OracleMaxPool", "OracleAvgPool", "OracleFlatten", "OracleCrossEntropy", "OracleMSE", "OracleNetwork",
    "lr_constant", "lr_step", "lr_time_based", "lr_exponential", "lr_linear",
]


def _tup(v, nd):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * nd


# --------------------------------------------------------------------------------------
# Dense  (okl.py:18-54, tensor.py:57-67 / 12-22 / 117-125)
# --------------------------------------------------------------------------------------
def dense_forward(x, W, b):
    """Y = X.W + b   (okl.py:31; Tensor.dot tensor.py:58 = np.dot; __add__ tensor.py:13)."""
    return np.dot(x, W) + b


def dense_backward(x, W, g):
    """okl.py:35-37: dW = X^T.G ; db = sum_rows(G) keepdims -> (1,out) ; dX = G.W^T (pre-update W)."""
    dW = np.dot(x.T, g)
    db = np.sum(g, axis=0, keepdims=True)
    dX = np.dot(g, W.T)
    return dW, db, dX


# --------------------------------------------------------------------------------------
# Conv1D / Conv2D / Conv3D  (okl.py:669-764, 767-883, 936-1049)
# --------------------------------------------------------------------------------------
def _conv_out_dtype(nd):
    # okl.py:799 (Conv2D: np.zeros default -> float64), okl.py:702 / 964 (Conv1D / Conv3D: float32)
    return np.float64 if nd == 2 else np.float32


def _pad(x, pad):
    nd = len(pad)
    if any(p > 0 for p in pad):
        return np.pad(x, ((0, 0), (0, 0)) + tuple((p, p) for p in pad), mode="constant")
    return x


def _out_extent(x_shape, k, s, p):
    # okl.py:695 / 792-793 / 952-954: floor division
    return tuple((x_shape[2 + d] + 2 * p[d] - k[d]) // s[d] + 1 for d in range(len(k)))


def conv_forward_loops(x, W, b, stride=1, padding=0):
    """Cross-correlation with zero padding, one Python iteration per output position.

    Follows okl.py:691-715 (1-D), 790-813 (2-D), 948-985 (3-D):
        y[n,o,pos] = sum_{c,tap} x_pad[n,c,pos*s+tap] * W[o,c,tap] + b[o]
    Output dtype: float32 for nd in (1,3), float64 for nd == 2 (reference allocation dtype).
    """
    nd = W.ndim - 2
    k = W.shape[2:]
    s, p = _tup(stride, nd), _tup(padding, nd)
    xp = _pad(x, p)
    out_ext = _out_extent(x.shape, k, s, p)
    N, O = x.shape[0], W.shape[0]
    y = np.zeros((N, O) + out_ext, dtype=_conv_out_dtype(nd))
    red_axes = tuple(range(2, 3 + nd))
    for pos in np.ndindex(*out_ext):
        win = tuple(slice(pos[d] * s[d], pos[d] * s[d] + k[d]) for d in range(nd))
        rf = xp[(slice(None), slice(None)) + win]                       # (N,C,*k)
        y[(slice(None), slice(None)) + pos] = np.sum(rf[:, np.newaxis] * W[np.newaxis], axis=red_axes)
    y += b.reshape((1, -1) + (1,) * nd)                                   # okl.py:711 / 811 / 982
    return y


def conv_backward_loops(x, W, g, stride=1, padding=0):
    """Gradients of conv_forward w.r.t. filters, biases and inputs, following the reference loops
    okl.py:717-754 (1-D), 848-883 (2-D), 987-1039 (3-D) with the repairs of SURVEY.md section 8(c):

      R1  the dL_dout slice gets a trailing axis per spatial dim so it broadcasts as (N,O,1,1[,1,1])
          against (N,1,C,*k) / (1,O,C,*k)                       (okl.py:865-868, 1013-1020)
      R2  bias gradient is reshaped to biases.shape == (O,1)     (okl.py:721, 870, 991)
      R3  dL_dinputs is accumulated at the *padded* shape and un-padded with explicit bounds
          (okl.py:722, 992 allocate unpadded; okl.py:881 uses a ``-0`` slice)

    Returns (dW, db (O,1), dX) in the dtype the reference would produce (zeros_like of W / x_pad).
    """
    nd = W.ndim - 2
    k = W.shape[2:]
    s, p = _tup(stride, nd), _tup(padding, nd)
    xp = _pad(x, p)
    out_ext = g.shape[2:]
    dW = np.zeros_like(W)
    dXp = np.zeros_like(xp, dtype=np.result_type(xp.dtype, W.dtype) if nd == 2 else x.dtype)
    tail = (np.newaxis,) * nd
    for pos in np.ndindex(*out_ext):
        win = tuple(slice(pos[d] * s[d], pos[d] * s[d] + k[d]) for d in range(nd))
        rf = xp[(slice(None), slice(None)) + win]                                   # (N,C,*k)
        gpos = g[(slice(None), slice(None)) + pos]                                   # (N,O)
        gb = gpos[(slice(None), slice(None), np.newaxis) + tail]                     # (N,O,1,*1)   [R1]
        dW += np.sum(rf[:, np.newaxis] * gb, axis=0)                                 # (O,C,*k)
        dXp[(slice(None), slice(None)) + win] += np.sum(W[np.newaxis] * gb, axis=1)  # (N,C,*k)
    db = np.sum(g, axis=(0,) + tuple(range(2, 2 + nd))).reshape(W.shape[0], 1)       # [R2]
    unpad = tuple(slice(p[d], xp.shape[2 + d] - p[d]) for d in range(nd))            # [R3]
    dX = dXp[(slice(None), slice(None)) + unpad]
    return dW, db, dX


def _windows(xp, k, s):
    """(N,C,*out,*k) strided view of the padded input."""
    nd = len(k)
    v = sliding_window_view(xp, k, axis=tuple(range(2, 2 + nd)))
    sl = (slice(None), slice(None)) + tuple(slice(None, None, s[d]) for d in range(nd))
    return v[sl]


_EINS = {
    1: ("nclt,oct->nol", "nclt,nol->oct"),
    2: ("nchwuv,ocuv->nohw", "nchwuv,nohw->ocuv"),
    3: ("ncdhwtuv,octuv->nodhw", "ncdhwtuv,nodhw->octuv"),
}


def conv_forward_fast(x, W, b, stride=1, padding=0, dtype=None):
    """Vectorised twin of conv_forward_loops (same contraction, same output dtype rule)."""
    nd = W.ndim - 2
    k = W.shape[2:]
    s, p = _tup(stride, nd), _tup(padding, nd)
    dt = dtype or _conv_out_dtype(nd)
    xp = _pad(x, p)
    win = _windows(xp, k, s)
    y = np.einsum(_EINS[nd][0], win, W, optimize=True).astype(dt, copy=False)
    y = y + b.reshape((1, -1) + (1,) * nd).astype(dt)
    return y


def conv_backward_fast(x, W, g, stride=1, padding=0):
    """Vectorised twin of conv_backward_loops (R1-R3 semantics)."""
    nd = W.ndim - 2
    k = W.shape[2:]
    s, p = _tup(stride, nd), _tup(padding, nd)
    xp = _pad(x, p)
    win = _windows(xp, k, s)
    dW = np.einsum(_EINS[nd][1], win, g, optimize=True).astype(W.dtype, copy=False)
    db = np.sum(g, axis=(0,) + tuple(range(2, 2 + nd))).reshape(W.shape[0], 1)
    dXp = np.zeros(xp.shape, dtype=np.result_type(xp.dtype, W.dtype) if nd == 2 else x.dtype)
    out_ext = g.shape[2:]
    # scatter one filter tap at a time (k^nd iterations instead of prod(out) iterations)
    for tap in np.ndindex(*k):
        sl = tuple(slice(tap[d], tap[d] + s[d] * (out_ext[d] - 1) + 1, s[d]) for d in range(nd))
        wt = W[(slice(None), slice(None)) + tap]                       # (O,C)
        contrib = np.tensordot(g, wt, axes=([1], [0]))                 # (N,*out,C)
        dXp[(slice(None), slice(None)) + sl] += np.moveaxis(contrib, -1, 1)
    unpad = tuple(slice(p[d], xp.shape[2 + d] - p[d]) for d in range(nd))
    return dW, db, dXp[(slice(None), slice(None)) + unpad]


# --------------------------------------------------------------------------------------
# Unfold / Fold  (okl.py:568-636 / 525-565) - define the im2col K-ordering (c, u, v)
# --------------------------------------------------------------------------------------
def unfold_forward(x, kernel_size, stride=1):
    """im2col for padding == 0 (okl.py:575-612): out (N, C*k*k, OH*OW), row = c*k*k + u*k + v,
    column = i*OW + j.  (The reference's padding>0 branch misplaces border patches, SURVEY 8(a) IM;
    it is not restated.)"""
    N, C, H, Wd = x.shape
    k = kernel_size
    win = _windows(x, (k, k), (stride, stride))                          # (N,C,OH,OW,k,k)
    OH, OW = win.shape[2], win.shape[3]
    return win.transpose(0, 1, 4, 5, 2, 3).reshape(N, C * k * k, OH * OW)


def unfold_forward_loops(x, kernel_size, stride=1, padding=0):
    """okl.py:575-612 as written, including the padding > 0 behaviour: a border patch is the CLIPPED window copied into the
    top-left corner of a zero k x k patch (not true zero padding, SURVEY 8(a) IM).  Output float64 (np.zeros default)."""
    N, C, H, Wd = x.shape
    k = kernel_size
    OH = (H + 2 * padding - k) // stride + 1
    OW = (Wd + 2 * padding - k) // stride + 1
    out = np.zeros((N, C * k * k, OH * OW))
    for i in range(OH):
        for j in range(OW):
            hs, ws = i * stride - padding, j * stride - padding
            patch = x[:, :, max(0, hs):min(H, hs + k), max(0, ws):min(Wd, ws + k)]
            if patch.shape[2] < k or patch.shape[3] < k:
                pp = np.zeros((N, C, k, k))
                pp[:, :, :patch.shape[2], :patch.shape[3]] = patch
                patch = pp
            out[:, :, i * OW + j] = patch.reshape(N, -1)
    return out


def unfold_backward(g, input_shape, kernel_size, stride=1, padding=0):
    """okl.py:614-636: col2im scatter-add.  With padding > 0 the clipped border slice cannot take the (N, C, k, k) block and
    NumPy raises ValueError - exactly what the reference does."""
    N, C, H, Wd = input_shape
    k = kernel_size
    OH = (H + 2 * padding - k) // stride + 1
    OW = (Wd + 2 * padding - k) // stride + 1
    dx = np.zeros(input_shape)
    g6 = np.asarray(g).reshape(N, C, k, k, OH, OW)
    for i in range(OH):
        for j in range(OW):
            hs, ws = i * stride - padding, j * stride - padding
            dx[:, :, max(0, hs):min(H, hs + k), max(0, ws):min(Wd, ws + k)] += g6[:, :, :, :, i, j]
    return dx


def fold_backward(g, input_shape, output_size, kernel_size, stride=1, padding=0):
    """okl.py:552-565: the inverse reshape / transposition of Fold.forward."""
    N = g.shape[0]
    k = kernel_size
    H, Wd = g.shape[2], g.shape[3]
    OH = (H + 2 * padding - k) // stride + 1
    OW = (Wd + 2 * padding - k) // stride + 1
    return g.reshape(N, -1, OH, k, OW, k).transpose(0, 1, 3, 5, 2, 4).reshape(input_shape)


def fold_forward_ref(cols, output_size, kernel_size, stride=1, padding=0):
    """okl.py:533-550 as written (output extents from stride / padding as the reference computes them)."""
    N, nch, L = cols.shape
    k = kernel_size
    H, Wd = output_size
    OH = (H + 2 * padding - k) // stride + 1
    OW = (Wd + 2 * padding - k) // stride + 1
    return cols.reshape(N, nch // (k * k), k, k, OH, OW).transpose(0, 1, 4, 2, 5, 3).reshape(N, -1, H, Wd)


def fold_forward(cols, output_size, kernel_size):
    """okl.py:533-550: pure reshape/transposition, equals col2im only when stride == kernel, p == 0."""
    N = cols.shape[0]
    k = kernel_size
    H, Wd = output_size
    C = cols.shape[1] // (k * k)
    OH, OW = H // k, Wd // k
    return cols.reshape(N, C, k, k, OH, OW).transpose(0, 1, 4, 2, 5, 3).reshape(N, C, H, Wd)


# --------------------------------------------------------------------------------------
# Activations  (okl.py:94-120, 500-522, 2070-2108, 355-399)
# --------------------------------------------------------------------------------------
def relu_forward(x):
    """Intended semantics of okl.py:103 ``max(0, x)`` elementwise, keeping x's float dtype."""
    return np.maximum(x, 0)


def relu_forward_as_shipped(x):
    """okl.py:103 through tensor.py:44-46: ``np.vectorize(lambda v: max(0, v))``.  np.vectorize infers
    the output dtype from the first element, so when x.flat[0] <= 0 the whole output is int64
    (SURVEY.md fact 5).  Parity fixtures force x.flat[0] > 0, where this equals relu_forward."""
    return np.vectorize(lambda v: max(0, v))(x)


def relu_backward(x, g):
    """okl.py:107-108: g * (1 if x > 0 else 0)."""
    return g * (x > 0)


def unpool_forward(x, pool_size=2, stride=2, average=False):
    """okl.py:2624-2637 / 2666-2679 as written (assignment: later stamps overwrite earlier ones)."""
    N, C, H, Wd = x.shape
    out = np.zeros((N, C, (H - 1) * stride + pool_size, (Wd - 1) * stride + pool_size))
    dv = pool_size * pool_size if average else 1
    for i in range(H):
        for j in range(Wd):
            out[:, :, i * stride:i * stride + pool_size, j * stride:j * stride + pool_size] = np.expand_dims(x[:, :, i, j], axis=(2, 3)) / dv
    return out


def unpool_backward(g, in_shape, pool_size=2, stride=2, average=False):
    """okl.py:2640-2649 / 2682-2691."""
    N, C, H, Wd = in_shape
    dx = np.zeros(in_shape)
    dv = pool_size * pool_size if average else 1
    for i in range(H):
        for j in range(Wd):
            dx[:, :, i, j] = np.sum(g[:, :, i * stride:i * stride + pool_size, j * stride:j * stride + pool_size], axis=(2, 3)) / dv
    return dx


def pad_forward(x, padding, reflect=False):
    """okl.py:1948-1949 / 1968: np.pad over the spatial axes; padding = ((top, bottom), (left, right))."""
    return np.pad(x, ((0, 0), (0, 0)) + tuple(padding), mode="reflect") if reflect else np.pad(x, ((0, 0), (0, 0)) + tuple(padding), mode="constant", constant_values=0)


def pad_backward(g, padding):
    """okl.py:1953-1955 as written (a zero bottom / right pad makes the slice empty)."""
    (pt, pb), (pl, pr) = padding
    return g[:, :, pt:-pb, pl:-pr]


def conv2d_transposed_forward(x, F, b, stride, padding):
    """okl.py:815-840 as written: F is (in_channels, out_channels, kh, kw); every input pixel scatters its k x k stamp, the
    result is cropped by the padding, then the bias is added."""
    s, p = _tup(stride, 2), _tup(padding, 2)
    N, Ci, H, Wd = x.shape
    Co, kh, kw = F.shape[1], F.shape[2], F.shape[3]
    OH, OW = (H - 1) * s[0] - 2 * p[0] + kh, (Wd - 1) * s[1] - 2 * p[1] + kw
    out = np.zeros((N, Co, OH + 2 * p[0], OW + 2 * p[1]))
    for i in range(H):
        for j in range(Wd):
            out[:, :, i * s[0]:i * s[0] + kh, j * s[1]:j * s[1] + kw] += np.sum(x[:, :, i, j][:, :, None, None, None] * F, axis=1)
    if p[0] > 0 or p[1] > 0:
        out = out[:, :, p[0]:out.shape[2] - p[0], p[1]:out.shape[3] - p[1]]
    return out + np.asarray(b).reshape(1, -1, 1, 1)


def conv2d_transposed_backward(x, F, g, stride, padding):
    """okl.py:891-917 (the loops as written; the bias gradient is returned as (O, 1) - the shipped update at okl.py:924 raises on
    its (1, O, 1, 1) shape exactly like the ordinary layer's, repair R2).  Returns dF, db, dX."""
    s, p = _tup(stride, 2), _tup(padding, 2)
    N, Ci, H, Wd = x.shape
    kh, kw = F.shape[2], F.shape[3]
    gp = np.pad(g, ((0, 0), (0, 0), (p[0], p[0]), (p[1], p[1])), mode="constant")
    dF, dX = np.zeros_like(F, dtype=np.float64), np.zeros((N, Ci, H, Wd))
    for i in range(H):
        for j in range(Wd):
            win = gp[:, None, :, i * s[0]:i * s[0] + kh, j * s[1]:j * s[1] + kw]
            dF += np.sum(x[:, :, i:i + 1, j:j + 1][:, :, None, :, :] * win, axis=0)
            dX[:, :, i, j] = np.sum(F * win, axis=(2, 3, 4))
    return dF, np.sum(g, axis=(0, 2, 3)).reshape(-1, 1), dX


class OracleRNN:
    """okl.py:1646-1697: hidden = tanh(x Wx + h_prev Wh + b); backward accumulates the three gradients (no update) and returns
    (dL/dx, dL/dh_prev)."""

    def __init__(self, Wh, Wx, b):
        self.Wh, self.Wx, self.b = np.array(Wh, dtype=np.float64), np.array(Wx, dtype=np.float64), np.array(b, dtype=np.float64)
        self.gWh = self.gWx = self.gb = None

    def forward(self, x, h_prev=None):
        if h_prev is None:
            h_prev = np.zeros((x.shape[0], self.Wh.shape[0]))
        self.x, self.h_prev = x, h_prev
        self.h = np.tanh(x @ self.Wx + h_prev @ self.Wh + self.b)
        return self.h

    def backward(self, g):
        dz = g * (1 - self.h ** 2)
        dWh, dWx, db = self.h_prev.T @ dz, self.x.T @ dz, np.sum(dz, axis=0, keepdims=True)
        self.gWh = dWh if self.gWh is None else self.gWh + dWh
        self.gWx = dWx if self.gWx is None else self.gWx + dWx
        self.gb = db if self.gb is None else self.gb + db
        return dz @ self.Wx.T, dz @ self.Wh.T


class OracleLSTM:
    """okl.py:1412-1534 line by line: gates on [x | h_prev]; backward accumulates the eight gradients (no update) and returns
    (dL/dh_prev, dL/dc_prev)."""

    def __init__(self, W, b, input_size):
        self.W = [np.array(w, dtype=np.float64) for w in W]          # Wf, Wi, Wc, Wo
        self.b = [np.array(v, dtype=np.float64) for v in b]          # bf, bi, bc, bo
        self.I = input_size
        self.gW, self.gb = [None] * 4, [None] * 4

    def forward(self, x, h_prev=None, c_prev=None):
        H = self.W[0].shape[1]
        h_prev = np.zeros((x.shape[0], H)) if h_prev is None else h_prev
        c_prev = np.zeros((x.shape[0], H)) if c_prev is None else c_prev
        sg = lambda z: 1 / (1 + np.exp(-z))                            # noqa: E731
        self.comb, self.c_prev = np.concatenate((x, h_prev), axis=1), c_prev
        self.f, self.i = sg(self.comb @ self.W[0] + self.b[0]), sg(self.comb @ self.W[1] + self.b[1])
        self.cc = np.tanh(self.comb @ self.W[2] + self.b[2])
        self.cell = self.f * c_prev + self.i * self.cc
        self.o = sg(self.comb @ self.W[3] + self.b[3])
        self.h = self.o * np.tanh(self.cell)
        return self.h, self.cell

    def backward(self, dh, dc):
        d_o = dh * np.tanh(self.cell)
        dcell = dc + dh * self.o * (1 - np.tanh(self.cell) ** 2)
        dz = [dcell * self.c_prev * self.f * (1 - self.f), dcell * self.cc * self.i * (1 - self.i), dcell * self.i * (1 - self.cc ** 2),
              d_o * self.o * (1 - self.o)]
        for k in range(4):
            dW, db = self.comb.T @ dz[k], np.sum(dz[k], axis=0, keepdims=True)
            self.gW[k] = dW if self.gW[k] is None else self.gW[k] + dW
            self.gb[k] = db if self.gb[k] is None else self.gb[k] + db
        dprev_h = sum(dz[k] @ self.W[k][self.I:].T for k in range(4))
        return dprev_h, dcell * self.f


class OracleGRU:
    """okl.py:1537-1643 line by line; backward accumulates the six gradients (no update), returns (dL/dx, dL/dh_prev)."""

    def __init__(self, W, b, input_size):
        self.W = [np.array(w, dtype=np.float64) for w in W]          # Wz, Wr, Wh
        self.b = [np.array(v, dtype=np.float64) for v in b]          # bz, br, bh
        self.I = input_size
        self.gW, self.gb = [None] * 3, [None] * 3

    def forward(self, x, h_prev=None):
        H = self.W[0].shape[1]
        h_prev = np.zeros((x.shape[0], H)) if h_prev is None else h_prev
        sg = lambda v: 1 / (1 + np.exp(-v))                            # noqa: E731
        self.x, self.h_prev = x, h_prev
        comb = np.concatenate((x, h_prev), axis=1)
        self.z, self.r = sg(comb @ self.W[0] + self.b[0]), sg(comb @ self.W[1] + self.b[1])
        self.hc = np.tanh(np.concatenate((x, self.r * h_prev), axis=1) @ self.W[2] + self.b[2])
        self.h = (1 - self.z) * h_prev + self.z * self.hc
        return self.h

    def backward(self, dh):
        I = self.I
        dz, dhc = dh * (self.hc - self.h_prev), dh * self.z
        dzh = dhc * (1 - self.hc ** 2)
        dr = (dzh @ self.W[2][I:].T) * self.h_prev
        dzz, dzr = dz * self.z * (1 - self.z), dr * self.r * (1 - self.r)
        comb, comb2 = np.concatenate((self.x, self.h_prev), axis=1), np.concatenate((self.x, self.r * self.h_prev), axis=1)
        for k, (c, d) in enumerate(((comb, dzz), (comb, dzr), (comb2, dzh))):
            dW, db = c.T @ d, np.sum(d, axis=0, keepdims=True)
            self.gW[k] = dW if self.gW[k] is None else self.gW[k] + dW
            self.gb[k] = db if self.gb[k] is None else self.gb[k] + db
        dprev = dzz @ self.W[0][I:].T + dzr @ self.W[1][I:].T + (dzh @ self.W[2][I:].T) * self.r + dh * (1 - self.z)
        dx = dzz @ self.W[0][:I].T + dzr @ self.W[1][:I].T + dzh @ self.W[2][:I].T
        return dx, dprev


class OracleBatchNorm:
    """okl.py:1145-1222 restated line by line (2-D inputs; backward uses the running variance in the chain rule as shipped;
    gamma / beta gradients accumulate forever and the update is applied inside backward)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        self.eps, self.momentum = eps, momentum
        self.gamma, self.beta = np.ones((1, num_features)), np.zeros((1, num_features))
        self.running_mean, self.running_var = np.zeros((1, num_features)), np.ones((1, num_features))
        self.ggrad = self.bgrad = None
        self.training = True

    def forward(self, x):
        self.B = x.shape[0]
        if self.training:
            mean, var = np.mean(x, axis=0), np.var(x, axis=0)
            self.running_mean = self.momentum * mean + (1 - self.momentum) * self.running_mean
            self.running_var = self.momentum * var + (1 - self.momentum) * self.running_var
        else:
            mean, var = self.running_mean, self.running_var
        var = np.maximum(var, self.eps)
        self.xc = x - mean
        self.xn = self.xc / np.sqrt(var + self.eps)
        return np.nan_to_num(self.gamma * self.xn + self.beta, nan=0.0, posinf=1.0, neginf=0.0)

    def backward(self, g, lr):
        dgamma, dbeta = np.sum(g * self.xn, axis=0), np.sum(g, axis=0)
        dxn = g * self.gamma
        rv = self.running_var + self.eps
        dvar = np.sum(dxn * self.xc * -0.5 * np.power(rv, -1.5), axis=0)
        dmean = np.sum(dxn * -1 / np.sqrt(rv), axis=0) + dvar * np.mean(-2 * self.xc, axis=0)
        dx = dxn / np.sqrt(rv) + dvar * 2 * self.xc / self.B + dmean / self.B
        self.ggrad = dgamma if self.ggrad is None else self.ggrad + dgamma
        self.bgrad = dbeta if self.bgrad is None else self.bgrad + dbeta
        self.gamma = self.gamma - lr * self.ggrad
        self.beta = self.beta - lr * self.bgrad
        return dx


# ---- the other elementwise activations, each as the reference computes it
_S2PI = np.sqrt(2 / np.pi)


def elu(x, alpha=1.0):                      # okl.py:70
    return np.where(x > 0, x, alpha * (np.exp(x) - 1))


def elu_d(x, alpha=1.0):                    # okl.py:74-78
    return np.where(x > 0, 1.0, alpha * np.exp(x))


def selu(x, alpha=1.6732632423543772848170429916717, scale=1.0507009873554804934193349852946):        # okl.py:179
    return scale * np.where(x > 0, x, alpha * (np.exp(x) - 1))


def selu_d(x, alpha=1.6732632423543772848170429916717, scale=1.0507009873554804934193349852946):      # okl.py:183-184
    return np.where(x > 0, scale, scale * alpha * np.exp(x))


def leaky(x, alpha=0.01):                   # okl.py:212, 245
    return np.where(x > 0, x, alpha * x)


def leaky_d(x, alpha=0.01):                 # okl.py:216, 249 (intended float semantics; np.vectorize truncates when x.flat[0] > 0)
    return np.where(x > 0, 1.0, alpha)


def prelu_alpha_grad(x, g):                 # okl.py:250-251
    return float(np.sum(g * np.where(x <= 0, x, 0.0)))


def swish(x):                               # okl.py:131-137
    return x * (1 / (1 + np.exp(-np.clip(x, -88, 88))))


def swish_d(x):                             # okl.py:141-149
    sg = 1 / (1 + np.exp(-np.clip(x, -88, 88)))
    sw = x * sg
    return np.clip(sw + sg * (1 - sw), -1e9, 1e9)


def gelu(x):                                # okl.py:277-285
    return 0.5 * x * (1 + np.tanh(np.clip(_S2PI * (x + 0.044715 * np.power(x, 3)), -15, 15)))


def gelu_d(x):                              # okl.py:291-301 AS WRITTEN (0.5 * tanh, not 0.5 * (1 + tanh))
    th = np.tanh(np.clip(_S2PI * (x + 0.044715 * np.power(x, 3)), -15, 15))
    return 0.5 * th + 0.5 * x * (1 - th ** 2) * _S2PI * (1 + 3 * 0.044715 * np.power(x, 2))


def softsign(x):                            # okl.py:322
    return x / (1 + np.abs(x))


def softsign_d(x):                          # okl.py:330
    return 1 / ((1 + np.abs(x)) ** 2)


def hardtanh(x):                            # okl.py:2135
    return np.clip(x, -1, 1)


def hardtanh_d(x):                          # okl.py:2139
    return np.logical_and(x >= -1, x <= 1).astype(float)


def softmin_forward(x):                     # okl.py:410-416 as written: exp(-(x + max))
    inv = np.exp(-(x + np.max(x, axis=1, keepdims=True)))
    return np.nan_to_num(inv / np.sum(inv, axis=1, keepdims=True), nan=0.0, posinf=1.0, neginf=0.0)


def logsoftmax_forward(x):                  # okl.py:455-468
    x = x.astype(np.float64)
    mx = np.max(x, axis=1, keepdims=True)
    return (x - mx) - np.log(np.sum(np.exp(np.clip(x - mx, -500, 500)), axis=1, keepdims=True))


def logsoftmax_backward(ls, g):             # okl.py:470-488: J[j,k] = 1 - exp(ls_j) if j == k else -exp(ls_k);  out = J . g
    return g - np.sum(np.exp(ls) * g, axis=1, keepdims=True)


def sigmoid_forward(x):
    """okl.py:509: 1 / (1 + exp(-x))."""
    return 1 / (1 + np.exp(-x))


def sigmoid_backward(y, g):
    """okl.py:514-515: g * y * (1 - y) using the stored output."""
    return g * (y * (1 - y))


def tanh_forward(x):
    """okl.py:2088."""
    return np.tanh(x)


def tanh_backward(x, g):
    """okl.py:2092: g * (1 - tanh(x)^2), recomputed from the input."""
    return g * (1 - np.tanh(x) ** 2)


def softmax_forward(x):
    """okl.py:364-373: stable softmax over axis 1 + nan_to_num."""
    sh = x - np.max(x, axis=1, keepdims=True)
    e = np.exp(sh)
    p = e / np.sum(e, axis=1, keepdims=True)
    return np.nan_to_num(p, nan=0.0, posinf=1.0, neginf=0.0)


def softmax_backward_jacobian(p, g):
    """okl.py:376-393 verbatim structure: J[i,j,k] = p_ij (delta_jk - p_ik); out = einsum('ijk,ik->ij')."""
    B, C = p.shape
    J = np.zeros((B, C, C))
    for i in range(B):
        for j in range(C):
            for k in range(C):
                if j == k:
                    J[i, j, k] = p[i, j] * (1 - p[i, k])
                else:
                    J[i, j, k] = -p[i, j] * p[i, k]
    return np.einsum("ijk,ik->ij", J, g)


def softmax_backward(p, g):
    """Closed form of softmax_backward_jacobian: p * (g - sum_k g_k p_k)  (SURVEY 8(a) S1)."""
    return p * (g - np.sum(g * p, axis=1, keepdims=True))


# --------------------------------------------------------------------------------------
# Pooling  (okl.py:1715-1831)
# --------------------------------------------------------------------------------------
def _pool_ext(H, Wd, ps, st):
    return (H - ps) // st + 1, (Wd - ps) // st + 1            # okl.py:1736-1737


def maxpool_forward(x, pool_size=2, stride=2):
    """okl.py:1731-1751: window max on NCHW, no padding, float64 output."""
    N, C, H, Wd = x.shape
    OH, OW = _pool_ext(H, Wd, pool_size, stride)
    y = np.zeros((N, C, OH, OW))
    for i in range(OH):
        for j in range(OW):
            hs, ws = i * stride, j * stride
            y[:, :, i, j] = np.max(x[:, :, hs:hs + pool_size, ws:ws + pool_size], axis=(2, 3))
    return y


def maxpool_backward(x, g, pool_size=2, stride=2):
    """okl.py:1753-1772 applied on NCHW [R4]: every position equal to its window's max receives
    (by assignment, later windows overwrite) the upstream gradient of that window."""
    dX = np.zeros_like(x, dtype=np.result_type(x.dtype, g.dtype))
    OH, OW = g.shape[2], g.shape[3]
    for i in range(OH):
        for j in range(OW):
            hs, ws = i * stride, j * stride
            win = x[:, :, hs:hs + pool_size, ws:ws + pool_size]
            mask = win == np.max(win, axis=(2, 3), keepdims=True)
            tgt = dX[:, :, hs:hs + pool_size, ws:ws + pool_size]
            np.copyto(tgt, np.broadcast_to(g[:, :, i:i + 1, j:j + 1], tgt.shape), where=mask)
    return dX


def avgpool_forward(x, pool_size=2, stride=2):
    """okl.py:1787-1807."""
    N, C, H, Wd = x.shape
    OH, OW = _pool_ext(H, Wd, pool_size, stride)
    y = np.zeros((N, C, OH, OW))
    for i in range(OH):
        for j in range(OW):
            hs, ws = i * stride, j * stride
            y[:, :, i, j] = np.mean(x[:, :, hs:hs + pool_size, ws:ws + pool_size], axis=(2, 3))
    return y


def avgpool_backward(x, g, pool_size=2, stride=2):
    """okl.py:1809-1825 applied on NCHW [R5]: dX[window] += g / ps^2."""
    dX = np.zeros_like(x, dtype=np.result_type(x.dtype, g.dtype))
    OH, OW = g.shape[2], g.shape[3]
    for i in range(OH):
        for j in range(OW):
            hs, ws = i * stride, j * stride
            dX[:, :, hs:hs + pool_size, ws:ws + pool_size] += g[:, :, i:i + 1, j:j + 1] / (pool_size * pool_size)
    return dX


# --------------------------------------------------------------------------------------
# Losses  (okl.py:1380-1409, 1833-1839)
# --------------------------------------------------------------------------------------
def cross_entropy_forward(p, t):
    """okl.py:1389-1397: mean(-log(clip(p,1e-12,1-1e-12)[range, t]))."""
    n = len(p)
    c = np.clip(p, 1e-12, 1 - 1e-12)
    return np.mean(-np.log(c[range(n), np.asarray(t).astype(int)]))


def cross_entropy_backward(p, t, denom=None):
    """okl.py:1399-1403: (clip(p) - onehot) / samples.  ``denom`` overrides samples for
    data-parallel shards (global batch)."""
    n = len(p)
    c = np.clip(p, 1e-12, 1 - 1e-12).copy()
    c[range(n), np.asarray(t).astype(int)] -= 1
    return c / (denom or n)


def mse_forward(o, t):
    """okl.py:1835."""
    return np.mean((o - t) ** 2)


def mse_backward(o, t, denom=None):
    """okl.py:1838: 2 (o - t) / t.size."""
    return 2 * (o - t) / (denom or t.size)


# --------------------------------------------------------------------------------------
# Parameter updates
# --------------------------------------------------------------------------------------
def accum_update(w, gacc, g, lr):
    """The reference's effective rule (okl.py:39-45, 744-748, 872-878, 1029-1033):
    gacc = g if gacc is None else gacc + g   (never reset);   w -= lr * gacc.
    Returns (w, gacc)."""
    gacc = g if gacc is None else gacc + g
    w = w - lr * gacc
    return w, gacc


def sgd_update(w, v, g, lr, momentum):
    """optimizers.py:140-144: v = mu v - lr g ; w += v (the shipped method then raises AttributeError
    while touching ndarray.grad, after the update was applied)."""
    v = momentum * (np.zeros_like(g) if v is None else v) - lr * g
    return w + v, v


class AdamState:
    """optimizers.py:20-26."""

    def __init__(self):
        self.m, self.v, self.t = {}, {}, 0


def adam_update(w, g, key, st, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
    """optimizers.py:28-42; ``t`` increments per call (per parameter tensor)."""
    st.t += 1
    if key not in st.m:
        st.m[key] = np.zeros_like(g)
        st.v[key] = np.zeros_like(g)
    st.m[key] = beta1 * st.m[key] + (1 - beta1) * g
    st.v[key] = beta2 * st.v[key] + (1 - beta2) * (g ** 2)
    m_hat = st.m[key] / (1 - beta1 ** st.t)
    v_hat = st.v[key] / (1 - beta2 ** st.t)
    return w - lr * m_hat / (np.sqrt(v_hat) + eps)


# --------------------------------------------------------------------------------------
# LR schedulers  (okl.py:2708-2761)
# --------------------------------------------------------------------------------------
def lr_constant(lr0, epoch):
    return lr0


def lr_step(lr0, step_size, gamma, epoch):
    return lr0 * (gamma ** (epoch // step_size))


def lr_time_based(lr0, decay, epoch):
    return lr0 / (1 + decay * epoch)


def lr_exponential(lr0, decay, epoch):
    return lr0 * np.exp(-decay * epoch)


def lr_linear(lr0, final_lr, total_epochs, epoch):
    return max(final_lr, lr0 - (lr0 - final_lr) / total_epochs * epoch)


# --------------------------------------------------------------------------------------
# Layer-shaped wrappers (ndarray in / ndarray out) so a train step reads like the reference's
# NeuralNetwork._forward/_backward (okl.py:2884-2914) with the in-layer update.
# --------------------------------------------------------------------------------------
class OracleDense:
    def __init__(self, W, b):
        self.W, self.b = np.array(W, dtype=np.float64), np.array(b, dtype=np.float64).reshape(1, -1)
        self.gW = self.gb = None

    def forward(self, x):
        self.x = x
        return dense_forward(x, self.W, self.b)

    def backward(self, g, lr):
        dW, db, dX = dense_backward(self.x, self.W, g)
        self.dW, self.db = dW, db
        self.W, self.gW = accum_update(self.W, self.gW, dW, lr)
        self.b, self.gb = accum_update(self.b, self.gb, db, lr)
        return dX


class OracleConv:
    def __init__(self, W, b, stride=1, padding=0, fast=False):
        nd = np.ndim(W) - 2
        dt = np.float64 if nd == 2 else np.float32        # okl.py:780 vs 686 / 945
        self.W, self.b = np.array(W, dtype=dt), np.array(b, dtype=dt).reshape(-1, 1)
        self.stride, self.padding, self.fast = stride, padding, fast
        self.gW = self.gb = None

    def forward(self, x):
        self.x = x
        f = conv_forward_fast if self.fast else conv_forward_loops
        return f(x, self.W, self.b, self.stride, self.padding)

    def backward(self, g, lr):
        f = conv_backward_fast if self.fast else conv_backward_loops
        dW, db, dX = f(self.x, self.W, g, self.stride, self.padding)
        self.dW, self.db = dW, db
        self.W, self.gW = accum_update(self.W, self.gW, dW, lr)
        self.b, self.gb = accum_update(self.b, self.gb, db.astype(self.b.dtype), lr)
        return dX


class OracleReLU:
    def __init__(self, as_shipped=False):
        self.as_shipped = as_shipped

    def forward(self, x):
        self.x = x
        return relu_forward_as_shipped(x) if self.as_shipped else relu_forward(x)

    def backward(self, g, lr):
        return relu_backward(self.x, g)


class OracleSigmoid:
    def forward(self, x):
        self.y = sigmoid_forward(x)
        return self.y

    def backward(self, g, lr):
        return sigmoid_backward(self.y, g)


class OracleTanh:
    def forward(self, x):
        self.x = x
        return tanh_forward(x)

    def backward(self, g, lr):
        return tanh_backward(self.x, g)


class OracleSoftmax:
    def forward(self, x):
        self.p = softmax_forward(x)
        return self.p

    def backward(self, g, lr):
        return softmax_backward(self.p, g)


class OracleMaxPool:
    def __init__(self, pool_size=2, stride=2):
        self.ps, self.st = pool_size, stride

    def forward(self, x):
        self.x = x
        return maxpool_forward(x, self.ps, self.st)

    def backward(self, g, lr):
        return maxpool_backward(self.x, g, self.ps, self.st)


class OracleAvgPool(OracleMaxPool):
    def forward(self, x):
        self.x = x
        return avgpool_forward(x, self.ps, self.st)

    def backward(self, g, lr):
        return avgpool_backward(self.x, g, self.ps, self.st)


class OracleFlatten:
    def forward(self, x):
        self.shape = x.shape
        return x.reshape(x.shape[0], -1)            # okl.py:643-648

    def backward(self, g, lr):
        return g.reshape(self.shape)                # okl.py:650-651


class OracleCrossEntropy:
    def forward(self, p, t):
        self.p, self.t = p, t
        return cross_entropy_forward(p, t)

    def backward(self, p, t, denom=None):
        return cross_entropy_backward(self.p, self.t, denom)


class OracleMSE:
    def forward(self, o, t):
        return mse_forward(o, t)

    def backward(self, o, t, denom=None):
        return mse_backward(o, t, denom)


class OracleNetwork:
    """okl.py:2884-2914: sequential forward (final output divided by temperature), reverse
    sequential backward with the update applied inside each parametric layer."""

    def __init__(self, layers, temperature=1.0):
        self.layers, self.temperature = list(layers), temperature

    def forward(self, x):
        self.acts = []
        for layer in self.layers:
            x = layer.forward(x)
            self.acts.append(x)
        x = x / self.temperature
        if self.layers and hasattr(self.layers[-1], "p"):
            self.layers[-1].p = x                   # okl.py:2890 rescales layer.outputs in place
        return x

    def backward(self, g, lr):
        self.grads = []
        for layer in reversed(self.layers):
            g = layer.backward(g, lr)
            self.grads.append(g)
        return g

    def train_step(self, x, t, loss, lr, denom=None):
        out = self.forward(x)
        val = loss.forward(out, t)
        g = loss.backward(out, t, denom) if denom is not None else loss.backward(out, t)
        self.backward(g, lr)
        return float(val)
