"""Device-resident ``Tensor`` with the reference's public surface (reference: src/okrolearn/tensor.py:5-67, 117-140,
1055-1084).

``.data`` / ``.grad`` keep their NumPy meaning for user code, but the storage of record lives in HBM: the host ndarray is a
lazily materialised mirror (device -> host on read, host -> device on the next device use after a write), so that
``forward`` / ``backward`` / ``train`` never round-trip through the host between ops (BASELINE.json north_star).

Semantics that differ from a plain ndarray attribute, by construction:
  * ``t.data`` returns a host *snapshot*.  ``t.data = arr`` and ``t.data -= x`` (getter + setter) write through.  Element
    writes into a previously returned snapshot (``t.data[0] = 1``) are detected on the next device use by a sampled
    checksum (best effort) and re-uploaded.
  * device arithmetic is fp32 (or TF32/BF16 inside the tensor-core kernels); the mirror is cast to the dtype the reference
    would have produced (float64 for Dense/Conv2D/pooling outputs, float32 for Conv1D/Conv3D).
"""
from __future__ import annotations

import ctypes as C
import zlib

import numpy as np

from . import _lib
from ._lib import lib, check, get_ctx, NCX, NXC


class _DevBuf:
    """Owning handle of one okl_malloc block."""
    __slots__ = ("ptr", "nbytes", "ctx", "__weakref__")

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.ptr = ctx.malloc(self.nbytes)
        keep = getattr(ctx, "_capture_keep", None)
        if keep is not None:          # buffers a captured CUDA graph writes must outlive the graph, never return to the pool
            keep.append(self)

    def __del__(self):
        try:
            if self.ptr:
                self.ctx.free(self.ptr)
                self.ptr = 0
        except Exception:      # interpreter shutdown
            pass


_PINNED = {}        # base address -> (nbytes, owner) of page-locked host arrays handed out by pinned_empty()


class _PinnedBlock:
    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib.okl_pinned_alloc(max(int(nbytes), 1), C.byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)
        _PINNED[self.ptr] = self

    def __del__(self):
        try:
            _PINNED.pop(self.ptr, None)
            lib.okl_pinned_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32):
    """NumPy array in page-locked (DMA-able, GPU-mapped) host memory: the source format of ``stream_inputs`` training."""
    dtype = np.dtype(dtype)
    n = _numel(shape) * dtype.itemsize
    blk = _PinnedBlock(n)
    raw = (C.c_ubyte * max(n, 1)).from_address(blk.ptr)
    arr = np.frombuffer(raw, dtype=np.uint8, count=n).view(dtype).reshape(shape)
    _KEEP[id(raw)] = blk
    import weakref
    weakref.finalize(raw, _KEEP.pop, id(raw), None)
    return arr


_KEEP = {}


def pinned_base(arr):
    """Address of ``arr``'s data if it lies inside a pinned_empty() block and is C-contiguous, else None."""
    if not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous:
        return None
    a = arr.ctypes.data
    for base, blk in _PINNED.items():
        if base <= a and a + arr.nbytes <= base + blk.nbytes:
            return a
    return None


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class Tensor:
    def __init__(self, data, requires_grad=True):
        # tensor.py:7 copies (np.array(data)); the dtype of the input is kept (ints stay ints)
        if isinstance(data, Tensor):
            data = data.data
        self._host = np.array(data)
        self._shape = self._host.shape
        self._np_dtype = self._host.dtype
        self._buf = None            # _DevBuf (owner) or None
        self._off = 0               # byte offset into _buf (parameter views)
        self._bound = False         # True for views that must keep their device location (layer parameters)
        self._dev_valid = False
        self._dev_ver = 0
        self._host_stamp = None
        self._layout = NCX
        self._exposed_sig = None
        self._grad = None
        self.requires_grad = requires_grad
        self.backward_fn = None

    # ------------------------------------------------------------------ construction on the device
    @classmethod
    def _empty(cls, shape, np_dtype=np.float64, layout=NCX, requires_grad=True):
        """Uninitialised device tensor (no host mirror yet)."""
        t = cls.__new__(cls)
        ctx = get_ctx()
        t._host = None
        t._shape = tuple(int(s) for s in shape)
        t._np_dtype = np.dtype(np_dtype)
        t._buf = _DevBuf(ctx, _numel(t._shape) * 4)
        t._off = 0
        t._bound = False
        t._dev_valid = True
        t._dev_ver = 0
        t._host_stamp = None
        t._layout = layout
        t._exposed_sig = None
        t._grad = None
        t.requires_grad = requires_grad
        t.backward_fn = None
        return t

    @classmethod
    def _view(cls, buf, off, shape, np_dtype, bound=True):
        """Tensor aliasing ``buf`` at byte offset ``off`` (layer parameters inside one bucket)."""
        t = cls.__new__(cls)
        t._host = None
        t._shape = tuple(int(s) for s in shape)
        t._np_dtype = np.dtype(np_dtype)
        t._buf, t._off, t._bound = buf, int(off), bound
        t._dev_valid = True
        t._dev_ver = 0
        t._host_stamp = None
        t._layout = NCX
        t._exposed_sig = None
        t._grad = None
        t.requires_grad = True
        t.backward_fn = None
        return t

    # ------------------------------------------------------------------ indexed row views (gather-free batches)
    _rows = None                    # (dataset pointer, int32 row-index pointer): this tensor IS rows idx[0..N) of that dataset

    @classmethod
    def _row_view(cls, src_ptr, idx_ptr, shape, np_dtype, buf):
        """The batch of a shuffled epoch WITHOUT a gathered copy: consumers that address dataset rows through the index themselves
        (first-layer convolution, fp32 -> bf16 conversion, MSE: the ``*_rows_i32`` arguments of the C ABI) read ``_rows``; any other
        consumer asks for ``_dptr()`` / ``.data`` and the rows are gathered into ``buf`` then (once)."""
        t = cls._view(buf, 0, shape, np_dtype, bound=False)
        t._rows = (int(src_ptr), int(idx_ptr))

        def gather(tt):
            rows = tt._shape[0]
            if rows:
                check(lib.okl_gather_rows(get_ctx().h, tt._buf.ptr + tt._off, tt._rows[0], tt._rows[1], rows, tt.size // rows))
            tt._touch()
        t._lazy_fn = gather
        return t

    def _rows_pending(self):
        """(dataset pointer, index pointer) while the rows have not been gathered, else None."""
        return self._rows if (self._rows is not None and getattr(self, "_lazy_fn", None) is not None) else None

    # ------------------------------------------------------------------ bf16-primary tensors
    _b16_primary = False            # True: the channels-last bf16 buffer is the storage of record, fp32 is converted on demand

    @classmethod
    def _empty_b16(cls, shape, np_dtype=np.float64, layout=NXC):
        """Uninitialised device tensor stored as bf16 ONLY (activations / gradients between tensor-core layers in BF16 mode).  An
        fp32 view (`_ptr_raw` / `_dptr` / `.data`) is converted from it on demand and cached per device version."""
        t = cls._view(None, 0, shape, np_dtype, bound=False)
        t._layout = layout
        t._b16_primary = True
        t._bf16 = _DevBuf(get_ctx(), max(t.size, 1) * 2)
        t._bf16_ver = (t._dev_ver, layout)
        t._f32_stamp = None
        return t

    def _f32_from_b16(self):
        ctx = get_ctx()
        stamp = (self._dev_ver, ctx.epoch)
        if self._buf is None or self._f32_stamp != stamp:
            if self._buf is None:
                self._buf, self._off = _DevBuf(ctx, max(self.size, 1) * 4), 0
            check(lib.okl_from_bf16(ctx.h, self._bf16.ptr, self._buf.ptr, self.size))
            self._f32_stamp = stamp

    # ------------------------------------------------------------------ shape without touching the host
    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        return _numel(self._shape)

    # ------------------------------------------------------------------ host mirror
    _SIG_FULL_BYTES = 32 << 20

    def _sig(self, a):
        """Signature of a host snapshot handed out by ``.data`` (checked on the next device use to catch in-place edits such as
        ``layer.weights.data[i, j] = 0``).  Up to 32 MiB: CRC-32 of every byte - any edit is seen.  Larger arrays (datasets): the
        full float64 sum plus a strided sample - an edit that preserves both goes unnoticed; assign through ``t.data = arr`` there."""
        if a.size == 0:
            return 0
        if a.nbytes <= self._SIG_FULL_BYTES and a.dtype != object:
            return zlib.crc32(memoryview(np.ascontiguousarray(a)).cast("B"))
        flat = a.reshape(-1)
        step = max(1, flat.size // 4096)
        with np.errstate(all="ignore"):
            return (float(np.sum(flat, dtype=np.float64)), float(np.sum(flat[::step].astype(np.float64)) + float(flat[-1])))

    def _fetch(self):
        ctx = get_ctx()
        n = self.size
        src = self._ptr_raw()
        tmp = None
        if self._layout == NXC and len(self._shape) >= 3 and self._shape[1] > 1 and _numel(self._shape[2:]) > 1:
            tmp = _DevBuf(ctx, n * 4)          # channels-last on the device -> reference layout for the host
            check(lib.okl_permute(ctx.h, tmp.ptr, src, self._shape[0], self._shape[1], _numel(self._shape[2:]), NXC))
            src = tmp.ptr
        out_dt = self._np_dtype if self._np_dtype in (np.dtype(np.float32), np.dtype(np.float64)) else np.dtype(np.float64)
        host = np.empty(self._shape, dtype=out_dt)
        if n:
            check(lib.okl_d2h(ctx.h, host.ctypes.data, src, n, _lib.np_to_okl_dtype(out_dt)))
        if out_dt != self._np_dtype:
            host = host.astype(self._np_dtype)
        self._host = host
        self._host_stamp = (self._dev_ver, ctx.epoch)
        del tmp

    @classmethod
    def wrap(cls, array, requires_grad=True):
        """Tensor over ``array`` WITHOUT the defensive copy of ``Tensor(array)`` (e.g. a pinned_empty() dataset)."""
        t = cls(np.zeros((0,)), requires_grad=requires_grad)
        t.data = array
        return t

    def _materialize(self):
        """A tensor a fused kernel did not write (e.g. the un-pooled activation of a fused Conv->ReLU->MaxPool block) is
        recomputed on first access."""
        fn = getattr(self, "_lazy_fn", None)
        if fn is not None:
            self._lazy_fn = None
            fn(self)

    @property
    def data(self):
        self._materialize()
        self._apply_lazy_perm()
        if self._dev_valid:
            ctx = get_ctx()
            if self._host is None or self._host_stamp != (self._dev_ver, ctx.epoch):
                self._fetch()
            self._exposed_sig = self._sig(self._host)
        return self._host

    @data.setter
    def data(self, value):
        if isinstance(value, Tensor):
            value = value.data
        arr = np.asarray(value)
        if self._bound and _numel(arr.shape) != _numel(self._shape):
            raise ValueError(f"cannot assign array of shape {arr.shape} to a parameter of shape {self._shape}")
        self._host = arr
        self._shape = arr.shape
        self._np_dtype = arr.dtype
        self._dev_valid = False
        self._exposed_sig = None
        self._lazy_perm = None               # a pending reorder belonged to the data that was just replaced

    def _apply_lazy_perm(self):
        """train() reorders the caller's dataset every epoch (okl.py:2959-2960).  Here the reorder is recorded as a pending row
        permutation and applied on the first use of the tensor by ANY consumer - host (``.data``) or device (``_dptr``: a row
        gather on the device, no host round trip) - so both always see the same, reference-ordered rows."""
        perm = getattr(self, "_lazy_perm", None)
        if perm is None:
            return
        self._lazy_perm = None
        rows = self._shape[0] if self._shape else 0
        if perm.shape[0] != rows or rows == 0:
            return
        if self._dev_valid and self._buf is not None and not self._bound:
            ctx = get_ctx()
            idx = _DevBuf(ctx, rows * 4)
            p64 = np.ascontiguousarray(perm, dtype=np.int64)
            check(lib.okl_h2d_i32(ctx.h, idx.ptr, p64.ctypes.data, rows, _lib.I64))
            nb = _DevBuf(ctx, max(self.size, 1) * 4)
            check(lib.okl_gather_rows(ctx.h, nb.ptr, self._ptr_raw(), idx.ptr, rows, self.size // rows))
            self._buf, self._off = nb, 0
            self._touch()                    # host mirror (if any) is stale now; re-fetched on the next .data read
            del idx
        else:
            host = self._host if self._host is not None else self.data
            self._host = np.asarray(host)[perm]
            self._dev_valid = False
            self._exposed_sig = None

    def to_numpy(self):
        return self.data

    @property
    def grad(self):
        g = self._grad
        return g.data if isinstance(g, Tensor) else g

    @grad.setter
    def grad(self, value):
        self._grad = value

    # ------------------------------------------------------------------ device side
    def _ptr_raw(self):
        if self._b16_primary:
            self._f32_from_b16()
        return self._buf.ptr + self._off

    def _touch(self):
        """Call after a kernel wrote this tensor's device memory."""
        self._dev_ver += 1
        self._dev_valid = True
        self._exposed_sig = None

    def _upload(self):
        ctx = get_ctx()
        host = np.ascontiguousarray(self._host)
        n = host.size
        if self._buf is None or (not self._bound and self._buf.nbytes < n * 4):
            self._buf, self._off = _DevBuf(ctx, n * 4), 0
        dt = host.dtype
        if dt == np.bool_:
            host, dt = host.astype(np.uint8), np.dtype(np.uint8)
        if dt not in _lib._NP2OKL:
            if not (np.issubdtype(dt, np.integer) or np.issubdtype(dt, np.floating)):
                raise ValueError(f"Tensor data of dtype {dt} cannot be used on the device")
            host = host.astype(np.float64)
            dt = host.dtype
        if n:
            check(lib.okl_h2d(ctx.h, self._ptr_raw(), host.ctypes.data, n, _lib._NP2OKL[dt]))
        self._layout = NCX
        self._dev_valid = True
        self._dev_ver += 1
        self._host_stamp = (self._dev_ver, ctx.epoch)
        self._exposed_sig = None

    def _dptr(self, layout=None):
        """Device pointer of an up-to-date fp32 copy, in ``layout`` (NCX/NXC) when given."""
        self._materialize()
        self._apply_lazy_perm()
        if self._dev_valid and self._exposed_sig is not None and self._host is not None:
            if self._sig(self._host) != self._exposed_sig:      # the snapshot handed out by .data was mutated in place
                self._dev_valid = False
        if not self._dev_valid:
            self._upload()
        if layout is not None and layout != self._layout and len(self._shape) >= 3:
            self._relayout(layout)
        return self._ptr_raw()

    def _b16(self, layout=NCX):
        """Device pointer of an up-to-date bf16 copy of this tensor in ``layout``; converted on demand, or supplied by the kernel
        that produced the tensor (``_new_b16``)."""
        ctx = get_ctx()
        if (self._b16_primary and self._dev_valid and self._bf16_ver == (self._dev_ver, self._layout) and
                (layout == self._layout or len(self._shape) < 3)):
            return self._bf16.ptr                       # storage of record: no fp32 detour
        if (layout == NXC and len(self._shape) >= 3 and not self._b16_primary and self._shape[0] > 0 and self._shape[1] > 1 and
                self._shape[1] % 8 == 0 and _numel(self._shape[2:]) > 1 and self._shape[0] <= 65535):
            # fp32 in the caller's (N, C, S) layout -> bf16 (N, S, C) in ONE pass; the fp32 copy keeps its layout
            src = self._dptr(None)
            if self._layout == NCX:
                if getattr(self, "_bf16", None) is None or self._bf16.nbytes < self.size * 2:
                    self._bf16, self._bf16_ver = _DevBuf(ctx, max(self.size, 1) * 2), None
                if self._bf16_ver != (self._dev_ver, "nxc-of-ncx"):
                    check(lib.okl_to_bf16_nxc(ctx.h, src, None, self._bf16.ptr, self._shape[0], self._shape[1], _numel(self._shape[2:])))
                    self._bf16_ver = (self._dev_ver, "nxc-of-ncx")
                return self._bf16.ptr
        src = self._dptr(layout)
        if (getattr(self, "_bf16", None) is None or self._bf16_ver != (self._dev_ver, self._layout) or
                self._bf16.nbytes < self.size * 2):
            if getattr(self, "_bf16", None) is None or self._bf16.nbytes < self.size * 2:
                self._bf16 = _DevBuf(ctx, max(self.size, 1) * 2)
            check(lib.okl_to_bf16(ctx.h, src, self._bf16.ptr, self.size))
            self._bf16_ver = (self._dev_ver, self._layout)
        return self._bf16.ptr

    def _new_b16(self):
        """Fresh bf16 shadow buffer that the producing kernel fills together with the fp32 data."""
        self._bf16 = _DevBuf(get_ctx(), max(self.size, 1) * 2)
        self._bf16_ver = (self._dev_ver, self._layout)
        return self._bf16.ptr

    def _relayout(self, layout):
        ctx = get_ctx()
        if self._b16_primary:
            self._f32_from_b16()
        N, Cc, S = self._shape[0], self._shape[1], _numel(self._shape[2:])
        if Cc > 1 and S > 1 and N > 0:
            nb = _DevBuf(ctx, self.size * 4)
            check(lib.okl_permute(ctx.h, nb.ptr, self._ptr_raw(), N, Cc, S, self._layout))
            if self._bound:
                raise RuntimeError("bound parameter views cannot change layout")
            self._buf, self._off = nb, 0
        self._b16_primary = False                       # the re-laid-out fp32 copy is the storage of record from here on
        self._layout = layout

    # ------------------------------------------------------------------ reference ops (tensor.py:12-67, 117-140)
    def _binary(self, other, op):
        ctx = get_ctx()
        if not isinstance(other, Tensor):
            if op == 0:
                raise AttributeError("'%s' object has no attribute 'data'" % type(other).__name__)   # tensor.py:13
            other = Tensor(np.asarray(other, dtype=np.float64))
        a, b = self, other
        if a.size < b.size:
            a, b = b, a
        out_shape = np.broadcast_shapes(self._shape, other._shape)
        ok = out_shape == a._shape and (b.size == a.size or b.size == 1 or
                                        (b.size == a._shape[-1] and all(s == 1 for s in b._shape[:-1])))
        if not ok:
            raise NotImplementedError(f"device broadcast of {self._shape} with {other._shape} is not supported")
        out = Tensor._empty(out_shape, np.result_type(self._np_dtype, other._np_dtype, np.float32))
        check(lib.okl_binary(ctx.h, op, a._dptr(NCX), b._dptr(NCX), out._ptr_raw(), a.size, b.size))
        return out

    def __add__(self, other):
        out = self._binary(other, 0)
        a, b = self, other

        def backward_fn(grad):
            if a.requires_grad:
                a.grad = grad if a.grad is None else a.grad + grad
            if b.requires_grad:
                b.grad = grad if b.grad is None else b.grad + grad

        out.backward_fn = backward_fn
        return out

    def __mul__(self, other):
        out = self._binary(other, 1)
        a, b = self, other

        def backward_fn(grad):
            if isinstance(b, Tensor):
                if a.requires_grad:
                    a.grad = grad * b.data if a.grad is None else a.grad + grad * b.data
                if b.requires_grad:
                    b.grad = grad * a.data if b.grad is None else b.grad + grad * a.data
            elif a.requires_grad:
                a.grad = grad * b if a.grad is None else a.grad + grad * b

        out.backward_fn = backward_fn
        return out

    def __rmul__(self, other):
        return self.__mul__(other)

    def dot(self, other):
        """2-D matrix product on the device (tensor.py:57-67)."""
        ctx = get_ctx()
        if self.ndim != 2 or other.ndim != 2 or self._shape[1] != other._shape[0]:
            raise ValueError(f"shapes {self._shape} and {other._shape} not aligned")
        M, K = self._shape
        N = other._shape[1]
        out = Tensor._empty((M, N), np.result_type(self._np_dtype, other._np_dtype, np.float32))
        check(lib.okl_dense_fwd(ctx.h, self._dptr(NCX), other._dptr(NCX), None, out._ptr_raw(), M, K, N, 0))
        a, b = self, other

        def backward_fn(grad):
            # dA = G . B^T, dB = A^T . G (tensor.py:60-65) as device GEMMs: nothing of a backward pass runs on the host
            g = grad if isinstance(grad, Tensor) else Tensor(np.asarray(grad, dtype=np.float64), requires_grad=False)
            ctx_ = get_ctx()
            if a.requires_grad:
                ga = Tensor._empty((M, K), np.float64)
                check(lib.okl_dense_bwd(ctx_.h, a._dptr(NCX), b._dptr(NCX), g._dptr(NCX), ga._ptr_raw(), None, None, M, K, N, None))
                a._grad = ga if a._grad is None else (a._grad if isinstance(a._grad, Tensor) else Tensor(a._grad)) + ga
            if b.requires_grad:
                gb = Tensor._empty((K, N), np.float64)
                check(lib.okl_dense_bwd(ctx_.h, a._dptr(NCX), b._dptr(NCX), g._dptr(NCX), None, gb._ptr_raw(), None, M, K, N, None))
                b._grad = gb if b._grad is None else (b._grad if isinstance(b._grad, Tensor) else Tensor(b._grad)) + gb

        out.backward_fn = backward_fn
        return out

    def transpose(self):
        ctx = get_ctx()
        if self.ndim != 2:
            out = Tensor(self.data.T)
        else:
            R, Cc = self._shape
            out = Tensor._empty((Cc, R), self._np_dtype)
            if R * Cc:
                check(lib.okl_permute(ctx.h, out._ptr_raw(), self._dptr(NCX), 1, R, Cc, NCX))
        a = self

        def backward_fn(grad):
            if a.requires_grad:
                a.grad = grad.T if a.grad is None else a.grad + grad.T

        out.backward_fn = backward_fn
        return out

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], tuple):
            shape = shape[0]
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = -_numel(shape)
            if known == 0 or self.size % known:
                raise ValueError(f"cannot reshape array of size {self.size} into shape {shape}")
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        if _numel(shape) != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {shape}")
        # zero-copy view of the same device block, in the reference (channels-first) element order
        self._dptr(NCX)
        out = Tensor._view(self._buf, self._off, shape, self._np_dtype, bound=False)
        out._dev_ver = self._dev_ver
        a, old = self, self._shape

        def backward_fn(grad):
            if a.requires_grad:
                a.grad = grad.reshape(old) if a.grad is None else a.grad + grad.reshape(old)

        out.backward_fn = backward_fn
        return out

    def apply(self, func):
        """tensor.py:44-55: ``np.vectorize(func)`` - an arbitrary Python callable per element can only run on the host.
        Not used by any layer of this package (activations are device kernels)."""
        vectorized_func = np.vectorize(func)
        out = Tensor(vectorized_func(self.data))
        a = self

        def backward_fn(grad):
            if a.requires_grad:
                d = grad * vectorized_func(a.data, derivative=True)
                a.grad = d if a.grad is None else a.grad + d

        out.backward_fn = backward_fn
        return out

    def copy(self, requires_grad=None):
        """tensor.py:1055-1076."""
        if requires_grad is None:
            requires_grad = self.requires_grad
        new = Tensor(np.copy(self.data), requires_grad=requires_grad)
        if self.grad is not None:
            new.grad = np.copy(self.grad)
        return new

    def __repr__(self):
        return f"Tensor({self.data})"

    def backward(self):
        if self.backward_fn is None:
            raise RuntimeError("No backward function defined for this tensor.")
        self.backward_fn(np.ones_like(self.data))
