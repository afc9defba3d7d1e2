"""SGDOptimizer / AdamOptimizer with the reference's ``update(param, grad, key)`` protocol (optimizers.py:4-47, 126-147).

``param`` may be a device ``Tensor`` (updated in place by a CUDA kernel; optimizer state lives in HBM) or - as in the
reference - a NumPy array, which is uploaded, updated on the device and written back in place.  ``NeuralNetwork.train`` only
reaches an optimizer for layers that define ``parameters`` / ``gradients`` dicts (okl.py:2984-2988); none of the shipped layers
does, so the effective rule of the hot path is the in-layer accumulated update (layers._ParamBucket)."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import lib, check, get_ctx, NCX
from .tensor import Tensor, _DevBuf


def _zeros(ctx, n):
    b = _DevBuf(ctx, n * 4)
    check(lib.okl_memset(ctx.h, b.ptr, 0, n * 4))
    return b


class _DeviceOptimizer:
    def _prep(self, layer, grad):
        wt = layer if isinstance(layer, Tensor) else Tensor(layer)
        gt = grad if isinstance(grad, Tensor) else Tensor(grad)
        if wt.size != gt.size:
            raise ValueError(f"operands could not be broadcast together with shapes {wt.shape} {gt.shape}")
        return wt, gt

    def _finish(self, layer, wt):
        wt._touch()
        if not isinstance(layer, Tensor):
            layer[...] = wt.data.reshape(layer.shape)            # in-place on the caller's ndarray (``layer -= update``)


class AdamOptimizer(_DeviceOptimizer):
    """optimizers.py:4-47.  ``t`` increments per ``update`` call (per parameter tensor), as in the reference."""

    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.beta1, self.beta2, self.epsilon = lr, beta1, beta2, epsilon
        self.m, self.v, self.t = {}, {}, 0

    def update(self, layer, grad, key):
        ctx = get_ctx()
        self.t += 1
        wt, gt = self._prep(layer, grad)
        if key not in self.m:
            self.m[key], self.v[key] = _zeros(ctx, wt.size), _zeros(ctx, wt.size)
        check(lib.okl_adam_update(ctx.h, wt._dptr(NCX), self.m[key].ptr, self.v[key].ptr, gt._dptr(NCX), wt.size, float(self.lr),
                                  float(self.beta1), float(self.beta2), float(self.epsilon), int(self.t)))
        self._finish(layer, wt)

    def reset(self):
        self.m, self.v, self.t = {}, {}, 0


class SGDOptimizer(_DeviceOptimizer):
    """optimizers.py:126-147: v = momentum v - lr g ; w += v.  (The shipped method then raises AttributeError while setting
    ``.grad`` on the ndarray, after the update was applied; that accident is not reproduced.)"""

    def __init__(self, lr=0.01, momentum=0.9):
        self.lr, self.momentum = lr, momentum
        self.velocity = {}

    def update(self, layer, grad, key):
        ctx = get_ctx()
        wt, gt = self._prep(layer, grad)
        if key not in self.velocity:
            self.velocity[key] = _zeros(ctx, wt.size)
        check(lib.okl_sgd_update(ctx.h, wt._dptr(NCX), self.velocity[key].ptr, gt._dptr(NCX), wt.size, float(self.lr),
                                 float(self.momentum)))
        self._finish(layer, wt)
        if isinstance(layer, Tensor):
            layer.grad = None
            layer.backward_fn = None
