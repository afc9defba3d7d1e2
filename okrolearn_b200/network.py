"""``NeuralNetwork``: the container and train loop of the reference (okl.py:2764-3041) driving device-resident layers.

Same public methods and hook names.  What is different underneath (BASELINE.json north_star: "the training-step loop and
the optimizer update" are rebuilt):
  * the dataset is uploaded to HBM once; the per-epoch shuffle (okl.py:2957-2960, same host RNG call so the permutation is
    identical) is a device gather; batches are zero-copy views.  ``stream_inputs = True`` keeps the dataset in pinned host
    memory instead and copies each batch host->device inside the step (the end-to-end measurement mode).
  * with no hooks / debug mode / breakpoints / profiler attached (nothing can observe intermediate Python objects), a full
    batch's forward + loss + backward + update is captured once into a CUDA graph and replayed: one launch per step, the
    learning rate read from a device scalar.  ``use_graph = False`` (or OKL_GRAPH=0) disables this.
  * in that fast path a parametric layer directly followed by ReLU/Sigmoid applies the activation in its GEMM/conv epilogue
    (``fuse_activations``), and the gradient w.r.t. the *dataset* of the first layer is not computed (nothing can read it).
  * per-batch losses are accumulated on the device and read back once per epoch.
  * data parallel (one process per GPU): each layer's gradient bucket is all-reduced on the comm stream as soon as its
    backward has been enqueued, and all updates are applied after the last layer's backward (overlap with backward).
"""
from __future__ import annotations

import cProfile
import ctypes as C
import io
import logging
import os
import pdb
import pickle
import pstats
import time
from typing import Callable

import numpy as np

from . import _lib
from . import layers as _layers_mod
from ._lib import lib, check, get_ctx, NCX, ACT_NONE
from .tensor import Tensor, _DevBuf, _numel, pinned_empty, pinned_base
from .layers import (apply_pending_updates, ParamArena, _ParamLayer, _Activation, _Pool, ReLUActivationLayer, SigmoidActivationLayer, SoftmaxActivationLayer,
                     MaxPoolingLayer, DenseLayer, Conv2DLayer, FlattenLayer,
                     _ConvBase)
from .losses import CrossEntropyLoss, MSELoss, LossValue


class _Slot:
    """One of the two input slots of the train loop: static device buffers a captured step reads its batch from + that graph."""

    def __init__(self, ctx, batch_size, sample_shape, t_row):
        self.xin = Tensor._empty((batch_size,) + tuple(sample_shape), np.float32)
        self.xbuf = self.xin._buf                  # the slot's own storage (layers only ever see views of it)
        self.tin = _DevBuf(ctx, max(batch_size * t_row, 1) * 4)
        self.idx = _DevBuf(ctx, max(batch_size, 1) * 4)   # the batch's slice of the epoch permutation (gather-free consumers)
        self.graph = None


class _StepState:
    """Per (batch shape, loss, layer list) state of train(): two input slots (step s+1's batch is gathered into the other slot
    while step s computes), the epoch-loss accumulator both graphs add to, and whether an eager warm-up step has run."""

    def __init__(self, ctx, batch_size, sample_shape, t_row):
        self.slots = [_Slot(ctx, batch_size, sample_shape, t_row), _Slot(ctx, batch_size, sample_shape, t_row)]
        self.eloss = _DevBuf(ctx, 4)
        self.warm = False


class _StepGraph:
    def __init__(self, handle, keep, launches):
        self.handle, self.keep, self.launches = handle, keep, launches

    def __del__(self):
        try:
            if self.handle:
                lib.okl_graph_destroy(self.handle)
        except Exception:
            pass


class NeuralNetwork:
    def __init__(self, temperature=1.0):
        self.layers = []
        self.hooks = {'pre_forward': [], 'post_forward': [], 'pre_backward': [], 'post_backward': [], 'pre_epoch': [],
                      'post_epoch': []}
        self.temperature = temperature
        self.profiler = cProfile.Profile()
        self.is_profiling = False
        self.debug_mode = False
        self.layer_outputs = []
        self.gradients = []
        self.parameter_history = []
        self.logger = self._setup_logger()
        self.breakpoints = {}
        self.error_history = []
        # B200 path knobs
        self.use_graph = os.environ.get("OKL_GRAPH", "1") != "0"
        self.fuse_activations = os.environ.get("OKL_FUSE", "1") != "0"
        self.skip_input_grad = os.environ.get("OKL_SKIP_INPUT_GRAD", "1") != "0"
        self.fuse_conv_pool = os.environ.get("OKL_FUSE_CONV_POOL", "1") != "0"
        self.fuse_head = os.environ.get("OKL_FUSE_HEAD", "1") != "0"
        self.prefetch_batches = os.environ.get("OKL_PREFETCH", "1") != "0"
        self.early_exchange = os.environ.get("OKL_EARLY_EXCHANGE", "1") != "0"
        self.use_aux_lanes = os.environ.get("OKL_AUX_LANES", "1") != "0"
        self.arena_bytes = int(os.environ.get("OKL_ARENA_BYTES", str(64 << 20)))   # networks up to this size share one arena
        self.stream_inputs = False          # keep the dataset on the host, copy each batch inside the step
        self.sync_loss_every_step = False   # read the loss back to the host after every step (end-to-end mode)
        self.last_step_losses = []
        self._graphs = {}
        self._plan_key = None

    # ------------------------------------------------------------------ logging / debug (okl.py:2787-2869)
    def _setup_logger(self):
        logger = logging.getLogger('NeuralNetwork')
        logger.setLevel(logging.DEBUG)
        if not logger.handlers:              # the reference adds one handler per instance (duplicated lines); one is enough
            handler = logging.StreamHandler()
            handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
            logger.addHandler(handler)
        return logger

    def set_debug_mode(self, mode: bool):
        self.debug_mode = mode
        self.logger.info(f"Debug mode set to: {mode}")

    def set_breakpoint(self, method_name: str, condition: Callable = None):
        self.breakpoints[method_name] = condition
        self.logger.info(f"Breakpoint set for method: {method_name}")

    def remove_breakpoint(self, method_name: str):
        if method_name in self.breakpoints:
            del self.breakpoints[method_name]
            self.logger.info(f"Breakpoint removed for method: {method_name}")
        else:
            self.logger.warning(f"No breakpoint found for method: {method_name}")

    def _check_breakpoint(self, method_name: str, *args, **kwargs):
        if method_name in self.breakpoints:
            condition = self.breakpoints[method_name]
            if condition is None or condition(*args, **kwargs):
                self.logger.info(f"Breakpoint hit in method: {method_name}")
                pdb.set_trace()

    def add(self, layer):
        self._check_breakpoint('add', layer)
        self.layers.append(layer)
        self._graphs.clear()

    def remove(self, index):
        self._check_breakpoint('remove', index)
        del self.layers[index]
        self._graphs.clear()

    def add_hook(self, hook_type: str, hook_fn: Callable):
        if hook_type in self.hooks:
            self.hooks[hook_type].append(hook_fn)
        else:
            raise ValueError(f"Invalid hook type: {hook_type}")

    def remove_hook(self, hook_type: str, hook_fn: Callable):
        if hook_type in self.hooks and hook_fn in self.hooks[hook_type]:
            self.hooks[hook_type].remove(hook_fn)

    def _run_hooks(self, hook_type: str, *args, **kwargs):
        for hook in self.hooks[hook_type]:
            hook(*args, **kwargs)

    def start_profiling(self):
        self.profiler.enable()
        self.is_profiling = True

    def stop_profiling(self):
        self.profiler.disable()
        self.is_profiling = False

    def print_profile_stats(self, sort_by='cumulative', lines=20):
        if not self.is_profiling:
            print("Profiling was not started. Use start_profiling() first.")
            return
        s = io.StringIO()
        pstats.Stats(self.profiler, stream=s).sort_stats(sort_by).print_stats(lines)
        print(s.getvalue())

    def _log_layer_output(self, layer_index: int, output: 'Tensor'):
        if self.debug_mode:
            self.layer_outputs.append((layer_index, output.copy()))

    def _log_gradient(self, layer_index: int, gradient: 'Tensor'):
        if self.debug_mode and gradient is not None:
            self.gradients.append((layer_index, gradient.copy()))

    def _log_parameters(self):
        if self.debug_mode:
            self.parameter_history.append([layer.get_params() for layer in self.layers])

    def get_layer_outputs(self):
        return self.layer_outputs

    def get_gradients(self):
        return self.gradients

    def parameters(self):
        return []                                        # okl.py:3018-3019

    def get_parameter_history(self):
        return [param.data.copy() for param in self.parameters()]

    # ------------------------------------------------------------------ planning (fast path only)
    def _observed(self):
        """True when user code can observe per-layer Python objects during a step."""
        return bool(self.debug_mode or self.breakpoints or self.is_profiling or any(self.hooks[k] for k in
                    ('pre_forward', 'post_forward', 'pre_backward', 'post_backward')))

    def _plan(self, fast: bool):
        key = (fast, self.fuse_activations, self.skip_input_grad, self.fuse_conv_pool, self.fuse_head, self.use_aux_lanes, tuple(id(l) for l in self.layers), get_ctx().nranks,
               get_ctx().precision)
        if key == self._plan_key:
            return
        self._plan_key = key
        ls = self.layers
        # small networks: one arena for every parameter bucket (single all-reduce, single update kernel)
        pbs = [l._bucket for l in ls if isinstance(l, _ParamLayer)]
        if len(pbs) > 1:
            for b in pbs:
                b.materialize()
            if sum(b.total for b in pbs) * 4 <= self.arena_bytes and not (pbs[0].arena is not None and pbs[0].arena.buckets == pbs):
                self._arena = ParamArena(pbs)
        for l in ls:
            if isinstance(l, _ParamLayer):
                l._fused_act, l._need_dx, l._dx_relu_mask, l._aux_lane = ACT_NONE, True, False, 0
                l._fused_head = l._big_head = False
                l._defer_update = fast or get_ctx().nranks > 1
            if isinstance(l, _ConvBase):
                l._fused_pool = None
            if isinstance(l, _Activation):
                l._fused_into_prev = False
                l._bwd_fused_into_next = False
                l._in_fused_block = False
            if isinstance(l, _Pool):
                l._fused_relu_bwd = False
                l._in_fused_block = False
                l._db_sink = None
            if isinstance(l, SoftmaxActivationLayer):
                l._fused_into_prev = False
            if isinstance(l, FlattenLayer):
                l._lazy_for_head = False
        # layout planning: a convolution that feeds (through activations / pooling) a tensor-core-eligible convolution writes
        # its output channels-last, so no permutation is needed between them
        tc_mode = get_ctx().precision != _lib.FP32
        for i, l in enumerate(ls):
            if isinstance(l, _ConvBase):
                l._out_layout = NCX
                # a small-K convolution that ends the network (e.g. the Conv3D video block): candidate for the first-layer tcgen05 kernels
                l._out_nxc_if_first = bool(tc_mode and all(isinstance(n, _Activation) for n in ls[i + 1:]) and not l._tc_eligible())
                if tc_mode:
                    for nxt in ls[i + 1:]:
                        if isinstance(nxt, (_Activation, _Pool)):
                            continue
                        if isinstance(nxt, _ConvBase) and nxt._tc_eligible():
                            l._out_layout = _lib.NXC
                        break
        if not fast:
            return
        if self.use_aux_lanes:                            # weight-gradient leaves on alternating auxiliary streams
            lane = 0
            for l in ls:
                if isinstance(l, _ParamLayer):
                    l._aux_lane = 1 + lane % 2
                    lane += 1
        if self.fuse_activations:
            for a, b in zip(ls, ls[1:]):
                if isinstance(a, _ParamLayer) and isinstance(b, (ReLUActivationLayer, SigmoidActivationLayer)):
                    a._fused_act, b._fused_into_prev = b._act, True
            for a, b in zip(ls, ls[1:]):                 # small Dense -> Softmax: softmax in the dense kernel's epilogue
                if (isinstance(a, DenseLayer) and isinstance(b, SoftmaxActivationLayer) and a.weights.shape[1] <= 16 and
                        a.weights.shape[0] * a.weights.shape[1] <= 12000):
                    a._fused_act, b._fused_into_prev = _lib.ACT_SOFTMAX, True
                    # ... and when that pair ends the network, a CrossEntropyLoss computes the whole head in its own kernel
                    a._fused_head = bool(self.fuse_head and b is ls[-1] and a.weights.shape[0] <= 256)
            # ... Flatten -> Dense(K up to a few thousand, N <= 16) -> Softmax at the end of a BF16 conv stack: the big fused head
            if (self.fuse_head and tc_mode and get_ctx().precision == _lib.BF16 and len(ls) >= 3 and isinstance(ls[-1], SoftmaxActivationLayer) and
                    isinstance(ls[-2], DenseLayer) and isinstance(ls[-3], FlattenLayer) and ls[-2].weights.shape[1] <= 16 and
                    256 < ls[-2].weights.shape[0] and ls[-2].weights.shape[0] % 256 == 0 and
                    ls[-2].weights.shape[0] * ls[-2].weights.shape[1] * 2 <= 160 * 1024):
                ls[-2]._fused_act, ls[-1]._fused_into_prev = _lib.ACT_SOFTMAX, True
                ls[-2]._fused_head = ls[-2]._big_head = True
                ls[-3]._lazy_for_head = True
            for a, r, b in zip(ls, ls[1:], ls[2:]):      # Param(+ReLU fused) -> Param: relu' applied in the second layer's dX epilogue
                if (isinstance(a, _ParamLayer) and isinstance(r, ReLUActivationLayer) and r._fused_into_prev and
                        isinstance(b, _ParamLayer)):
                    r._bwd_fused_into_next, b._dx_relu_mask = True, True
            for a, b in zip(ls, ls[1:]):                 # ReLU -> MaxPool: relu' applied inside the pooling backward
                if isinstance(a, ReLUActivationLayer) and isinstance(b, MaxPoolingLayer):
                    a._bwd_fused_into_next, b._fused_relu_bwd = True, True
            for cv, a, b in zip(ls, ls[1:], ls[2:]):     # Conv(+ReLU fused) -> ReLU -> MaxPool: the pooling backward's output IS the
                if (isinstance(cv, _ConvBase) and isinstance(a, ReLUActivationLayer) and a._fused_into_prev and      # conv's output gradient;
                        isinstance(b, MaxPoolingLayer) and b._fused_relu_bwd):                                     # it also reduces it to db
                    b._db_sink = cv
        if self.skip_input_grad and ls and isinstance(ls[0], _ParamLayer):
            ls[0]._need_dx = False
        # first block Conv2D(3x3, s1) -> ReLU -> MaxPool(2,2) with no dX needed: one fused kernel each way, the un-pooled
        # activation and its gradient are never materialised
        if (self.fuse_activations and self.fuse_conv_pool and len(ls) >= 3 and isinstance(ls[0], Conv2DLayer) and not ls[0].transposed and
                not ls[0]._need_dx and isinstance(ls[1], ReLUActivationLayer) and isinstance(ls[2], MaxPoolingLayer) and
                int(ls[2].pool_size) == 2 and int(ls[2].stride) == 2 and ls[0]._k == (3, 3) and ls[0]._s == (1, 1) and
                ls[0].in_channels in (1, 3) and ls[0].out_channels * ls[0].in_channels * 9 <= 10000):
            ls[0]._fused_pool = (ls[1], ls[2])
            ls[0]._fused_act = ACT_NONE
            ls[1]._in_fused_block = ls[2]._in_fused_block = True
            ls[1]._fused_into_prev = ls[1]._bwd_fused_into_next = ls[2]._fused_relu_bwd = False

    # ------------------------------------------------------------------ forward / backward (okl.py:2871-2914)
    def forward(self, inputs: 'Tensor'):
        self._check_breakpoint('forward', inputs)
        if self.is_profiling:
            return self._profiled_forward(inputs)
        return self._forward(inputs)

    def _profiled_forward(self, inputs: 'Tensor'):
        self.profiler.enable()
        result = self._forward(inputs)
        self.profiler.disable()
        return result

    def _forward(self, inputs: 'Tensor', _fast=False):
        self._plan(_fast)
        self._run_hooks('pre_forward', inputs)
        self.layer_outputs = []
        pending_banks = self._prefetch_filter_banks() if _fast else False
        for i, layer in enumerate(self.layers):
            if pending_banks and getattr(layer, "_conv2", False):
                check(lib.okl_aux_join(get_ctx().h))           # the banks were refreshed beside the first layer's forward
                pending_banks = False
            inputs = layer.forward(inputs)
            self._log_layer_output(i, inputs)
        if pending_banks:
            check(lib.okl_aux_join(get_ctx().h))
        if self.temperature != 1.0:
            # okl.py:2890 `inputs.data = inputs.data / self.temperature` - in place on the last layer's output Tensor
            ctx = get_ctx()
            check(lib.okl_scale(ctx.h, inputs._dptr(), inputs.size, 1.0 / float(self.temperature)))
            inputs._touch()
        self._run_hooks('post_forward', inputs)
        return inputs

    def _prefetch_filter_banks(self):
        """BF16 tensor-core stacks: the bf16 filter banks of every convolution (one small kernel each, refreshed once per step after
        the update) are produced on an auxiliary lane while the first layer's forward runs, instead of one by one in front of
        their layers on the critical path.  Returns True when something was enqueued (the caller joins before the first use)."""
        lays = [l for l in self.layers[1:] if getattr(l, "_conv2", False) and getattr(l, "_bank_desc", None) is not None and l._aux_lane]
        if len(lays) < 2 or not getattr(self.layers[0], "_first_tc", False):
            return False
        ctx = get_ctx()
        check(lib.okl_aux_begin(ctx.h, 2))
        try:
            for l in lays:
                l._filter_banks(l._bank_desc)
        finally:
            check(lib.okl_aux_end(ctx.h))
        return True

    def backward(self, loss_gradient: 'Tensor', lr: float):
        self._check_breakpoint('backward', loss_gradient, lr)
        if self.is_profiling:
            return self._profiled_backward(loss_gradient, lr)
        return self._backward(loss_gradient, lr)

    def _profiled_backward(self, loss_gradient: 'Tensor', lr: float):
        self.profiler.enable()
        result = self._backward(loss_gradient, lr)
        self.profiler.disable()
        return result

    def _backward(self, loss_gradient: 'Tensor', lr: float):
        self._run_hooks('pre_backward', loss_gradient, lr)
        self.gradients = []
        # data parallel + arena: once the backward pass reaches the first parametric layer every other layer's gradients are
        # complete - their all-reduce starts now and overlaps that layer's backward
        arena = getattr(self, "_arena", None) if (self.early_exchange and get_ctx().nranks > 1) else None
        first = None
        if arena is not None:
            first = next((l for l in self.layers if isinstance(l, _ParamLayer)), None)
            if first is None or first._bucket is not arena.buckets[0] or not all(
                    l._defer_update and l._bucket.arena is arena for l in self.layers if isinstance(l, _ParamLayer)):
                first = None
        for i, layer in reversed(list(enumerate(self.layers))):
            if layer is first:
                arena.exchange_tail()
            loss_gradient = layer.backward(loss_gradient, lr)
            self._log_gradient(i, loss_gradient)
        apply_pending_updates(self.layers)                 # deferred updates (fast path / data parallel): one launch
        self._run_hooks('post_backward', loss_gradient, lr)
        self._log_parameters()

    # ------------------------------------------------------------------ train (okl.py:2916-3001)
    def train(self, inputs: 'Tensor', targets: 'Tensor', epochs: int, lr_scheduler, optimizer, batch_size: int, loss_function):
        self._check_breakpoint('train', inputs, targets, epochs, lr_scheduler.get_lr(), optimizer, batch_size, loss_function)
        if self.is_profiling:
            return self._profiled_train(inputs, targets, epochs, lr_scheduler, optimizer, batch_size, loss_function)
        return self._train(inputs, targets, epochs, lr_scheduler, optimizer, batch_size, loss_function)

    def _profiled_train(self, inputs, targets, epochs, lr_scheduler, optimizer, batch_size, loss_function):
        self.profiler.enable()
        result = self._train(inputs, targets, epochs, lr_scheduler, optimizer, batch_size, loss_function)
        self.profiler.disable()
        return result

    @staticmethod
    def _class_indices(targets, num_samples):
        """int32 device copy of the class targets in their current storage order (okl.py:1392 does ``.astype(int)``), cached on
        the tensor until its host data is replaced."""
        ctx = get_ctx()
        host = targets._host if targets._host is not None else targets.data
        key = (id(host), targets._dev_ver)
        cached = getattr(targets, "_cls_dev", None)
        if cached is not None and cached[0] == key and cached[1].nbytes >= max(num_samples, 1) * 4:
            return cached[1]
        buf = _DevBuf(ctx, max(num_samples, 1) * 4)
        th = np.ascontiguousarray(np.asarray(host).reshape(-1).astype(np.int64))
        check(lib.okl_h2d_i32(ctx.h, buf.ptr, th.ctypes.data, th.size, _lib.I64))
        targets._cls_dev = (key, buf)
        return buf

    def prepare(self, inputs: 'Tensor', targets: 'Tensor', loss_function=None):
        """Dataset placement ahead of ``train`` (optional; ``train`` does the same lazily on first use): uploads the inputs and
        the targets (class indices for CrossEntropyLoss) to the device and sizes the per-epoch index buffers, so that a following
        ``train`` call starts stepping immediately.  With ``stream_inputs`` the dataset stays in page-locked host memory and only
        the buffers are sized."""
        ctx = get_ctx()
        num_samples = inputs.shape[0]
        self._epoch_pinned(num_samples)
        idx_dev = getattr(self, "_idx_dev", None)
        if idx_dev is None or idx_dev.nbytes < max(num_samples, 1) * 4:
            self._idx_dev = _DevBuf(ctx, max(int(num_samples * 1.5), 1024) * 4)
        if not self.stream_inputs and getattr(inputs, "_lazy_perm", None) is None:
            inputs._dptr(NCX)
            if isinstance(loss_function, CrossEntropyLoss):
                self._class_indices(targets, num_samples)
            elif getattr(targets, "_lazy_perm", None) is None:
                targets._dptr(NCX)
        ctx.sync()

    def _epoch_pinned(self, num_samples):
        """(scalar slot, int32 index buffer) in page-locked host memory, cached on the network and grown on demand."""
        cur = getattr(self, "_pin_epoch", None)
        if cur is None or cur[1].size < num_samples:
            cap = max(1024, int(num_samples * 1.5))
            self._pin_epoch = cur = (pinned_empty((1,), np.int32), pinned_empty((cap,), np.int32))
        return cur

    def _one_step(self, batch_inputs, batch_targets, loss_function, lr, optimizer, fast):
        """forward + loss + backward (+ optimizer for layers exposing parameters/gradients): okl.py:2969-2988."""
        outputs = self._forward(batch_inputs, _fast=fast) if fast else self.forward(batch_inputs)
        last = self.layers[-1] if self.layers else None
        if isinstance(loss_function, CrossEntropyLoss):
            loss_function._fuse_softmax = last if (fast and isinstance(last, SoftmaxActivationLayer)) else None
        elif type(loss_function) is MSELoss:      # trailing ReLU: its backward is applied inside the loss kernel
            loss_function._fuse_relu = last if (fast and self.fuse_activations and isinstance(last, ReLUActivationLayer) and
                                                self.temperature == 1.0 and not last._in_fused_block) else None
        loss = loss_function.forward(outputs, batch_targets)
        loss_gradient = loss_function.backward(outputs, batch_targets)
        if fast:
            self._backward(loss_gradient, lr)
        else:
            self.backward(loss_gradient, lr)
        for i, layer in enumerate(self.layers):
            if hasattr(layer, 'parameters'):
                for key, param in layer.parameters.items():
                    optimizer.update(param, layer.gradients[key], f"layer_{i}_{key}")
        return outputs, loss

    def _train(self, inputs: 'Tensor', targets: 'Tensor', epochs: int, lr_scheduler, optimizer, batch_size: int,
               loss_function, clip_grad: float = None):
        ctx = get_ctx()
        num_samples = inputs.shape[0]
        num_batches = int(np.ceil(num_samples / batch_size))
        losses = []
        # parameters / accumulated gradients the user assigned on the host since the last step (``layer.weights.data = ...``,
        # ``.grad = None``, ``load``) must be on the device before a captured step is replayed - replay touches no Python state
        def flush_host_params():
            for layer in self.layers:
                b = getattr(layer, "_bucket", None)
                if b is not None and getattr(b, "ready", False):
                    for t in b.tensors:
                        t._dptr(NCX)
                        if getattr(t, "_bf16", None) is not None:
                            t._b16()
                    b._sync_user_grads()
        flush_host_params()
        epoch_hooks = bool(self.hooks.get('pre_epoch') or self.hooks.get('post_epoch'))
        fast = not self._observed()
        graph_ok = (fast and self.use_graph and not any(hasattr(l, 'parameters') for l in self.layers) and clip_grad is None
                    and hasattr(type(loss_function), "_loss_accum"))
        is_ce = isinstance(loss_function, CrossEntropyLoss)
        sample_shape = tuple(inputs.shape[1:])
        t_shape = tuple(targets.shape[1:])
        row_elems = _numel(sample_shape)
        t_row = _numel(t_shape)
        stream = bool(self.stream_inputs)

        # ---- dataset placement.  The dataset is never physically shuffled: every step gathers the rows of its batch through
        # the epoch's permutation (device index vector) into the step's static input buffer - from HBM, or, with
        # `stream_inputs`, straight from page-locked host memory over PCIe (the host->device copy is then inside the step).
        # The reordering of the caller's tensors (okl.py:2959-2960) is applied lazily when `.data` is next read.
        pend_x, pend_t = getattr(inputs, "_lazy_perm", None), getattr(targets, "_lazy_perm", None)
        if (pend_x is not None and pend_t is not None and len(pend_x) == num_samples == targets.shape[0] and
                np.array_equal(pend_x, pend_t)):
            perm_total = pend_x.copy()                                   # storage order still predates an earlier train()
            inputs._lazy_perm = targets._lazy_perm = None
        else:
            inputs._apply_lazy_perm()                                    # device row gather / host fancy index, whichever holds the data
            targets._apply_lazy_perm()
            perm_total = np.arange(num_samples)
        keep = []
        if stream:
            hx = np.asarray(inputs._host if (inputs._host is not None and not inputs._dev_valid) else inputs.data)
            if hx.dtype == np.float32 and pinned_base(hx) is not None:
                pin_x = hx.reshape(num_samples, row_elems)               # zero-copy: already pinned
            else:
                pin_x = pinned_empty((num_samples, row_elems), np.float32)
                np.copyto(pin_x, np.asarray(hx, dtype=np.float32).reshape(num_samples, row_elems))
            ht = np.asarray(targets._host if (targets._host is not None and not targets._dev_valid) else targets.data)
            t_dtype = np.int32 if is_ce else np.float32
            if ht.dtype == t_dtype and pinned_base(ht) is not None:
                pin_t = ht.reshape(num_samples, t_row)                   # zero-copy: already pinned, already the device dtype
            else:
                pin_t = pinned_empty((num_samples, t_row), t_dtype)
                np.copyto(pin_t, ht.reshape(num_samples, t_row).astype(t_dtype))
            x_src, t_src = pin_x.ctypes.data, pin_t.ctypes.data
            keep += [pin_x, pin_t]
        else:
            x_src = inputs._dptr(NCX)
            if is_ce:
                t_buf = self._class_indices(targets, num_samples)      # cached on the targets tensor (prepare() fills it ahead)
                t_src = t_buf.ptr
                keep.append(t_buf)
            else:
                t_src = targets._dptr(NCX)
        # device index vector of the epoch's permutation: cached on the network and grown on demand (no allocation inside the loop
        # of repeated train() calls)
        idx_dev = getattr(self, "_idx_dev", None)
        if idx_dev is None or idx_dev.nbytes < max(num_samples, 1) * 4:
            idx_dev = self._idx_dev = _DevBuf(ctx, max(int(num_samples * 1.5), 1024) * 4)

        # ---- gather-free batches (device-resident dataset, BF16 convolution stacks): a consumer that can address dataset rows through
        # the batch's index slice reads them in place - the batch is never copied.  Decided per part: x for a first layer on the
        # small-K tcgen05 kernels (or converted to bf16 channels-last by the prefetch, straight from the dataset), targets for MSELoss
        # on bf16 channels-last outputs.  A consumer that turns out not to be index-aware gathers the rows on first use (Tensor._row_view).
        self._plan(fast)
        first = self.layers[0] if self.layers else None
        bf16_stack = bool(ctx.precision == _lib.BF16 and _layers_mod._BF16_ONLY and self.prefetch_batches and not stream and
                          os.environ.get("OKL_ROW_VIEWS", "1") != "0" and isinstance(first, _ConvBase) and batch_size <= 65535)
        x_rows = t_rows = False
        if bf16_stack and fast and not first._tc_eligible() and not getattr(first, "_need_dx", True) and first._fused_pool is None and \
                (getattr(first, "_out_layout", NCX) == _lib.NXC or getattr(first, "_out_nxc_if_first", False)) and len(sample_shape) == first._nd + 1:
            dd, _ = first._desc((batch_size,) + sample_shape, NCX, _lib.NXC)
            x_rows = bool(lib.okl_conv_first_supported(ctx.h, C.byref(dd)))
        if bf16_stack and fast and isinstance(loss_function, MSELoss) and len(t_shape) >= 2 and t_shape[0] % 8 == 0:
            last = next((l for l in reversed(self.layers) if isinstance(l, _ParamLayer)), None)
            t_rows = bool(isinstance(last, _ConvBase) and (x_rows or last._tc_eligible()) and
                          all(isinstance(l, _Activation) for l in self.layers[self.layers.index(last) + 1:]))
        rows_key = (x_src if (x_rows or bf16_stack) else 0, t_src if t_rows else 0)     # captured steps reference the dataset itself

        # ---- per-(batch shape, loss) step state: static input buffers + the captured graph
        # the stock losses are stateless: a fresh CrossEntropyLoss() / MSELoss() per train() call (the usual way to call it) reuses
        # the captured step instead of capturing - and keeping alive - a new set of graphs and activation buffers every time
        loss_key = type(loss_function) if type(loss_function) in (CrossEntropyLoss, MSELoss) else id(loss_function)
        skey = (batch_size, sample_shape, t_shape, loss_key, tuple(id(l) for l in self.layers), fast,
                self.fuse_activations, self.skip_input_grad, ctx.precision, bool(self.prefetch_batches), self.fuse_head,
                self.fuse_conv_pool, self.use_aux_lanes, self.early_exchange, ctx.nranks,
                float(self.temperature),         # 1/temperature is baked into the captured step by value
                x_rows, t_rows, rows_key)
        st = self._graphs.pop(skey, None)
        if st is None:
            st = _StepState(ctx, batch_size, sample_shape, t_row)
            while len(self._graphs) >= 4:                       # bounded: drop the least recently used configuration
                ctx.sync()
                self._graphs.pop(next(iter(self._graphs)))
        self._graphs[skey] = st                                 # (re-)inserted last = most recently used
        if not hasattr(ctx, "lr_buf"):
            ctx.lr_buf = _DevBuf(ctx, 4)
        ctx.lr_dev = ctx.lr_buf.ptr
        self.last_step_losses = []
        custom_loss = not hasattr(type(loss_function), "_loss_accum")
        if not custom_loss:
            loss_function._loss_accum = st.eloss.ptr

        def make_targets(rows, slot, pipelined=False):
            if is_ce:
                t = Tensor(np.zeros((0,)), requires_grad=False)
                t._shape = (rows,) + t_shape
                t._i32_dev = slot.tin.ptr
                return t
            if t_rows and pipelined:
                return Tensor._row_view(t_src, slot.idx.ptr, (rows,) + t_shape, np.float32, slot.tin)
            return Tensor._view(slot.tin, 0, (rows,) + t_shape, np.float32, bound=False)

        # a first layer that runs channels-last on the tensor cores gets its batch in that layout straight from the prefetch
        x_nxc = bool(self.prefetch_batches and ctx.precision != _lib.FP32 and isinstance(first, _ConvBase) and first._nd <= 2 and first._tc_eligible() and
                     getattr(first, "_fused_pool", None) is None and len(sample_shape) >= 2 and sample_shape[0] > 1 and _numel(sample_shape[1:]) > 1)
        # BF16 mode: that copy is written as bf16 (permutation + conversion in one pass on the copy stream) - the operand format of the
        # second-generation convolution, which then reads the batch without any fp32 detour
        x_b16 = bool(x_nxc and ctx.precision == _lib.BF16 and _layers_mod._BF16_ONLY and first.in_channels % 64 == 0 and
                     first.out_channels % 64 == 0 and batch_size <= 65535)
        if x_nxc:
            for sl_ in st.slots:
                if getattr(sl_, "xnxc", None) is None or sl_.xnxc.nbytes < max(batch_size * row_elems, 1) * (2 if x_b16 else 4):
                    sl_.xnxc = _DevBuf(ctx, max(batch_size * row_elems, 1) * (2 if x_b16 else 4))

        def prefetch(batch):
            """gather batch `batch` of the current permutation into slot batch % 2 on the copy stream"""
            b0 = batch * batch_size
            rows = min(b0 + batch_size, num_samples) - b0
            sl = st.slots[batch & 1]
            x_skip = x_rows or (x_b16 and bf16_stack)     # bf16 channels-last batch converted straight from the dataset rows
            if stream and host_idx[0] is not None:        # rows move on the copy engines: they need the indices on the host too
                check(lib.okl_batch_host_index(ctx.h, host_idx[0] + b0 * 4))
            check(lib.okl_batch_prefetch(ctx.h, batch & 1, None if x_skip else sl.xbuf.ptr, x_src, row_elems, None if t_rows else sl.tin.ptr, t_src,
                                         t_row, idx_dev.ptr + b0 * 4, rows, sl.xnxc.ptr if x_nxc else None, sample_shape[0] if x_nxc else 0,
                                         1 if x_b16 else 0, sl.idx.ptr if (x_rows or t_rows) else None))

        host_idx = [None]                                 # address of the epoch's permutation in page-locked host memory (int32)
        lossf = C.c_float()

        def collect(slot_idx):
            check(lib.okl_loss_fetch_wait(ctx.h, slot_idx, C.byref(lossf)))
            self.last_step_losses.append(float(lossf.value))

        for epoch in range(epochs):
            self._run_hooks('pre_epoch', epoch, epochs)
            if epoch_hooks:                          # an epoch hook may have assigned layer.weights.data / .grad on the host
                flush_host_params()
            epoch_start_time = time.time()
            current_lr = lr_scheduler.step(epoch)
            if hasattr(optimizer, 'param_groups'):
                for param_group in optimizer.param_groups:
                    param_group['lr'] = current_lr
            elif hasattr(optimizer, 'lr'):
                optimizer.lr = current_lr
            elif hasattr(optimizer, 'set_lr'):
                optimizer.set_lr(current_lr)
            else:
                self.logger.warning("Unable to update optimizer's learning rate. Scheduler may not have an effect.")
            # per-epoch scalars and the permutation go through cached page-locked buffers with asynchronous copies: no host sync
            # in the epoch prologue (the read of the epoch loss at the end of every epoch is the one synchronisation point,
            # so the buffers are never rewritten while a copy is in flight)
            pin = self._epoch_pinned(num_samples)
            pin[0].view(np.float32)[0] = current_lr
            check(lib.okl_h2d_async_raw(ctx.h, ctx.lr_dev, pin[0].ctypes.data, 4))
            check(lib.okl_memset(ctx.h, st.eloss.ptr, 0, 4))
            host_losses = []

            # shuffle (same host RNG call as okl.py:2957-2958)
            indices = np.arange(num_samples)
            np.random.shuffle(indices)
            perm_total = perm_total[indices]
            np.copyto(pin[1][:num_samples], perm_total, casting="unsafe")
            host_idx[0] = pin[1].ctypes.data if pin[1].dtype == np.int32 else None
            if num_samples:
                check(lib.okl_h2d_async_raw(ctx.h, idx_dev.ptr, pin[1].ctypes.data, num_samples * 4))

            pipelined = bool(self.prefetch_batches)
            if num_batches and pipelined:
                prefetch(0)
            pending_loss = None                      # slot whose loss copy is in flight (sync_loss_every_step, lag 1)
            for batch in range(num_batches):
                b0 = batch * batch_size
                b1 = min(b0 + batch_size, num_samples)
                rows = b1 - b0
                full = rows == batch_size
                sl = st.slots[batch & 1] if pipelined else st.slots[0]
                if pipelined:
                    if batch + 1 < num_batches:
                        prefetch(batch + 1)          # overlaps this step; ordered after the previous user of that slot
                    check(lib.okl_batch_wait(ctx.h, batch & 1))
                else:
                    check(lib.okl_gather_batch(ctx.h, sl.xbuf.ptr, x_src, row_elems, sl.tin.ptr, t_src, t_row,
                                               idx_dev.ptr + b0 * 4, rows))
                if graph_ok and full and sl.graph is not None:
                    ctx.graph_launch(sl.graph.handle)
                    loss = sl.graph.loss
                else:
                    # a fresh view per step: a layer may re-lay-out ITS input tensor (channels-last for the tensor-core convs),
                    # which must never redirect or relabel the slot's own buffer - the gather keeps writing NC.. rows there
                    if x_rows and pipelined:
                        bx = Tensor._row_view(x_src, sl.idx.ptr, (rows,) + sample_shape, np.float32, sl.xbuf)
                    elif x_b16 and pipelined:
                        bx = Tensor._view(None, 0, (rows,) + sample_shape, np.float32, bound=False)
                        bx._layout, bx._b16_primary, bx._bf16, bx._f32_stamp = _lib.NXC, True, sl.xnxc, None
                        bx._bf16_ver = (bx._dev_ver, _lib.NXC)
                    elif x_nxc and pipelined:
                        bx = Tensor._view(sl.xnxc, 0, (rows,) + sample_shape, np.float32, bound=False)
                        bx._layout = _lib.NXC
                    else:
                        bx = Tensor._view(sl.xbuf, 0, (rows,) + sample_shape, np.float32, bound=False)
                    bt = make_targets(rows, sl, pipelined)
                    if graph_ok and full and st.warm:
                        keep = []
                        ctx._capture_keep = keep
                        ctx.graph_begin()
                        try:
                            outputs, loss = self._one_step(bx, bt, loss_function, current_lr, optimizer, fast)
                        finally:
                            handle = ctx.graph_end()
                            ctx._capture_keep = None
                        keep.extend([outputs, loss, bx, bt])
                        sl.graph = _StepGraph(handle, keep, 0)
                        sl.graph.loss = loss
                        ctx.graph_launch(sl.graph.handle)
                    else:
                        outputs, loss = self._one_step(bx, bt, loss_function, current_lr, optimizer, fast)
                        if full:
                            st.warm = True
                    if clip_grad is not None:
                        self._clip_gradients(clip_grad)
                if custom_loss:                      # duck-typed user loss: its value lives on the host (okl.py:2973)
                    host_losses.append(float(np.sum(getattr(loss, "data", loss))))
                if self.sync_loss_every_step:
                    if isinstance(loss, LossValue) and pipelined:
                        # the loss of every step reaches the host, one step late: its copy is enqueued behind the step and
                        # collected after the NEXT step has been enqueued, so the device never waits for the host
                        check(lib.okl_loss_fetch_async(ctx.h, batch & 1, loss._buf.ptr))
                        if pending_loss is not None:
                            collect(pending_loss)
                        pending_loss = batch & 1
                    else:
                        self.last_step_losses.append(float(loss))
            if pending_loss is not None:
                collect(pending_loss)

            if custom_loss:
                avg_loss = float(np.sum(host_losses)) / num_batches
            else:
                v = C.c_float()
                check(lib.okl_read_scalar(ctx.h, st.eloss.ptr, C.byref(v)))
                avg_loss = float(v.value) / num_batches
            losses.append(avg_loss)
            self.parameter_history.append(self.get_parameter_history())
            epoch_end_time = time.time()
            self.logger.info(f"Epoch {epoch + 1}/{epochs} - Loss: {avg_loss:.4f} - LR: {current_lr:.6f} - "
                             f"Time: {epoch_end_time - epoch_start_time:.2f}s")
            self._run_hooks('post_epoch', epoch, epochs, avg_loss)

        # the reference reorders the caller's tensors in place every epoch (okl.py:2959-2960)
        if epochs:
            ctx.sync()
            inputs._lazy_perm = perm_total
            targets._lazy_perm = perm_total.copy()
        del keep
        if not custom_loss:
            loss_function._loss_accum = None
        return losses

    def _clip_gradients(self, max_norm):
        """okl.py:3003-3016 - a no-op in the reference too, because parameters() returns []."""
        total_norm = 0.0
        for param in self.parameters():
            if param.grad is not None:
                total_norm += np.linalg.norm(param.grad.data) ** 2
        total_norm = np.sqrt(total_norm)
        if total_norm > max_norm:
            clip_coef = max_norm / (total_norm + 1e-6)
            for param in self.parameters():
                if param.grad is not None:
                    param.grad.data *= clip_coef

    # ------------------------------------------------------------------ save / load (okl.py:3021-3044)
    def save(self, file_path: str):
        params = [layer.get_params() for layer in self.layers]
        with open(file_path, 'wb') as f:
            pickle.dump((params, self.temperature), f)

    def load(self, file_path: str):
        with open(file_path, 'rb') as f:
            params, self.temperature = pickle.load(f)
        for layer, param in zip(self.layers, params):
            layer.set_params(param)

    def save_weights(self, file_path: str):
        params = [layer.get_params() for layer in self.layers]
        with open(file_path, 'wb') as f:
            pickle.dump(params, f)

    def load_weights(self, file_path: str):
        with open(file_path, 'rb') as f:
            params = pickle.load(f)
        for layer, param in zip(self.layers, params):
            layer.set_params(param)

    # ------------------------------------------------------------------ exact checkpoints (not in the reference)
    def save_checkpoint(self, file_path: str):
        """``save`` plus the state the reference's file format loses: every layer's ACCUMULATED gradients (the in-layer update
        rule never resets them, okl.py:39-45, so a run resumed from ``save`` / ``load`` diverges from the uninterrupted one).
        Values are the device's fp32 numbers widened to float64 - exact - so training resumed from this file continues bit for
        bit.  The ``params`` entry is what ``save`` writes, so ``load_weights``-style readers can still use the file."""
        state = []
        for layer in self.layers:
            b = getattr(layer, "_bucket", None)
            if b is None:
                state.append(None)
                continue
            b.materialize()
            stepped = bool(getattr(b, "_stepped", False))
            state.append([np.asarray(t.grad).copy() if (stepped or t._grad is not None) else None for t in b.tensors])
        with open(file_path, 'wb') as f:
            pickle.dump({"format": "okrolearn-b200-checkpoint-1", "params": [layer.get_params() for layer in self.layers],
                         "accumulated_gradients": state, "temperature": self.temperature}, f)

    def load_checkpoint(self, file_path: str):
        with open(file_path, 'rb') as f:
            ck = pickle.load(f)
        if not isinstance(ck, dict) or ck.get("format") != "okrolearn-b200-checkpoint-1":
            raise ValueError("not a checkpoint written by save_checkpoint")
        if len(ck["params"]) != len(self.layers):
            raise ValueError(f"checkpoint has {len(ck['params'])} layers, the network {len(self.layers)}")
        self.temperature = ck["temperature"]
        for layer, param, acc in zip(self.layers, ck["params"], ck["accumulated_gradients"]):
            layer.set_params(param)
            b = getattr(layer, "_bucket", None)
            if b is None or acc is None:
                continue
            b.materialize()
            for t, g in zip(b.tensors, acc):
                t.grad = None if g is None else np.asarray(g)
            if any(g is not None for g in acc):
                b._stepped = True               # a later `.grad = None` by the user then really clears the accumulator

    def set_temperature(self, temperature: float):
        self.temperature = temperature


class _BorrowedBuf:
    """Non-owning stand-in for _DevBuf (a pointer into a buffer someone else keeps alive)."""
    __slots__ = ("ptr", "nbytes")

    def __init__(self, ptr, nbytes=0):
        self.ptr, self.nbytes = ptr, nbytes
