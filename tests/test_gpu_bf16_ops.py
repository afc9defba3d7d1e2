"""Layout- and precision-crossing kernels of the BF16 pipeline (okl_bf16_ops.cu) through the C ABI: fp32 (N, C, S) -> bf16 (N, S, C)
in one pass, MSELoss on bf16 channels-last outputs against fp32 reference-layout targets; and the BF16 Conv1D train loop that uses
both (BASELINE configs[3]a shape family) against the oracle network."""
import ctypes as C
import logging

import numpy as np
import pytest

from conftest import rel_err
from oracle import okl_oracle as O

pytestmark = pytest.mark.gpu
logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)


def bf16_round(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def fetch_bf16(lib, check, ctx, ptr, n):
    raw = np.empty(n, dtype=np.uint16)
    check(lib.okl_d2h_async_raw(ctx.h, raw.ctypes.data, ptr, n * 2))
    ctx.sync()
    return (raw.astype(np.uint32) << 16).view(np.float32)


SHAPES = [(3, 64, (1024,)), (2, 8, (5, 7)), (4, 24, (130,)), (1, 64, (4, 6, 10)), (5, 40, (129,)), (2, 16, (2,))]


@pytest.mark.parametrize("N,Cc,sp", SHAPES)
def test_to_bf16_nxc_is_exact(okl, N, Cc, sp):
    from okrolearn_b200 import _lib
    from okrolearn_b200.tensor import _DevBuf
    lib, check, ctx = _lib.lib, _lib.check, okl.get_ctx()
    rng = np.random.default_rng(N * 7 + Cc)
    x = rng.standard_normal((N, Cc) + sp).astype(np.float32)
    S = int(np.prod(sp))
    t = okl.Tensor(x)
    out = _DevBuf(ctx, x.size * 2)
    check(lib.okl_to_bf16_nxc(ctx.h, t._dptr(_lib.NCX), None, out.ptr, N, Cc, S))
    got = fetch_bf16(lib, check, ctx, out.ptr, x.size).reshape(N, S, Cc)
    want = np.moveaxis(bf16_round(x).reshape(N, Cc, S), 1, 2)
    assert np.array_equal(got, want)
    # the Tensor-level entry point (what a BF16 convolution calls on a reference-layout input) gives the same bits and keeps the
    # fp32 copy in the caller's layout
    p16 = t._b16(_lib.NXC)
    assert t._layout == _lib.NCX
    assert np.array_equal(fetch_bf16(lib, check, ctx, p16, x.size).reshape(N, S, Cc), want)


@pytest.mark.parametrize("relu", [0, 1])
@pytest.mark.parametrize("N,Cc,sp", SHAPES)
def test_mse_bf16_channels_last_vs_oracle(okl, N, Cc, sp, relu):
    from okrolearn_b200 import _lib
    from okrolearn_b200.tensor import _DevBuf
    lib, check, ctx = _lib.lib, _lib.check, okl.get_ctx()
    rng = np.random.default_rng(N * 11 + Cc + relu)
    o = bf16_round(rng.standard_normal((N, Cc) + sp).astype(np.float32))
    if relu:
        o = np.maximum(o, 0)
    tg = rng.standard_normal((N, Cc) + sp).astype(np.float32)
    S = int(np.prod(sp))
    to, tt = okl.Tensor(o), okl.Tensor(tg)
    loss, g16 = _DevBuf(ctx, 4), _DevBuf(ctx, o.size * 2)
    check(lib.okl_mse_fwd_bwd_nxc_bf16(ctx.h, to._b16(_lib.NXC), tt._dptr(_lib.NCX), None, loss.ptr, g16.ptr, N, Cc, S, float(o.size), None, relu))
    lv = np.empty(1, dtype=np.float32)
    check(lib.okl_d2h_async_raw(ctx.h, lv.ctypes.data, loss.ptr, 4))
    ctx.sync()
    want_loss = O.mse_forward(o.astype(np.float64), tg.astype(np.float64))
    want_g = O.mse_backward(o.astype(np.float64), tg.astype(np.float64))
    if relu:
        want_g = want_g * (o > 0)
    got = np.moveaxis(fetch_bf16(lib, check, ctx, g16.ptr, o.size).reshape(N, S, Cc), 2, 1).reshape(o.shape)
    assert abs(float(lv[0]) - want_loss) <= 1e-5 * abs(want_loss)
    assert rel_err(got, want_g) < 5e-3                      # bf16 storage of the gradient: 2^-9 relative per element


@pytest.mark.parametrize("graph", [True, False])
def test_bf16_conv1d_train_uses_bf16_batches_and_loss(okl, graph):
    """Conv1D(64, 64, 3, 1, 1) + ReLU + MSE in BF16 mode: the prefetch hands the layer a bf16 channels-last batch, the loss consumes the
    bf16 outputs and emits the bf16 gradient (no fp32 activation / gradient tensor on the path); two epochs with different batches,
    graph replay and eager, against the oracle network."""
    okl.set_precision("bf16")
    try:
        rng = np.random.default_rng(123)
        N, B, L = 48, 8, 256
        X = rng.standard_normal((N, 64, L)).astype(np.float32)
        X.reshape(-1)[0] = abs(X.reshape(-1)[0]) + 0.1
        T = rng.standard_normal((N, 64, L)).astype(np.float32)
        W = (rng.standard_normal((64, 64, 3)) * 0.1).astype(np.float32)
        net = okl.NeuralNetwork()
        conv = okl.Conv1DLayer(64, 64, 3, 1, 1)
        conv.weights.data = W.copy()
        net.add(conv)
        net.add(okl.ReLUActivationLayer())
        net.use_graph = graph
        np.random.seed(5)
        ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.05),
                       optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.MSELoss())
        on = O.OracleNetwork([O.OracleConv(W, np.zeros((64, 1)), 1, 1, fast=True), O.OracleReLU()])
        np.random.seed(5)
        Xo, To, ol = X.copy(), T.copy(), []
        for ep in range(2):
            idx = np.arange(N)
            np.random.shuffle(idx)
            Xo, To = Xo[idx], To[idx]
            ol.append(sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleMSE(), 0.05) for s in range(0, N, B)) / (N // B))
        assert rel_err(ls, ol) < 1e-2
        assert rel_err(conv.weights.data, on.layers[0].W) < 2e-2
        assert rel_err(conv.biases.data, on.layers[0].b) < 2e-2
    finally:
        okl.set_precision("fp32")


@pytest.mark.parametrize("shape,B", [((6, 3, 4, 10, 12), 2), ((2, 3, 16, 112, 112), 1)])
def test_bf16_conv3d_video_block_train_vs_oracle(okl, shape, B):
    """BASELINE configs[3]b in BF16 mode: Conv3D(3, 64, 3, 1, 1) + ReLU + MSE through NeuralNetwork.train - the small-K tcgen05 kernels
    (operand built in shared memory from a staged fp32 patch, K = 81 + bias column), channels-last bf16 output, bf16 MSE, weight and bias
    gradient from the bf16 gradient - against the oracle network; a small clip set and the full 16 x 112 x 112 extent."""
    okl.set_precision("bf16")
    try:
        rng = np.random.default_rng(11)
        N = shape[0]
        X = rng.standard_normal(shape).astype(np.float32)
        X.reshape(-1)[0] = abs(X.reshape(-1)[0]) + 0.1
        T = rng.standard_normal((N, 64) + shape[2:]).astype(np.float32)
        W = (rng.standard_normal((64, 3, 3, 3, 3)) * 0.1).astype(np.float32)
        b = (rng.standard_normal((64, 1)) * 0.1).astype(np.float32)
        net = okl.NeuralNetwork()
        conv = okl.Conv3DLayer(3, 64, 3, 1, 1)
        conv.filters.data, conv.biases.data = W.copy(), b.copy()
        net.add(conv)
        net.add(okl.ReLUActivationLayer())
        np.random.seed(3)
        ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.05),
                       optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.MSELoss())
        assert getattr(conv, "_first_tc", False), "the small-K tcgen05 kernels were not selected"
        on = O.OracleNetwork([O.OracleConv(W.astype(np.float64), b.astype(np.float64), 1, 1, fast=True), O.OracleReLU()])
        np.random.seed(3)
        Xo, To, ol = X.copy(), T.copy(), []
        for ep in range(2):
            idx = np.arange(N)
            np.random.shuffle(idx)
            Xo, To = Xo[idx], To[idx]
            ol.append(sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleMSE(), 0.05) for s in range(0, N, B)) / (N // B))
        assert rel_err(ls, ol) < 1e-2
        assert rel_err(conv.filters.data, on.layers[0].W) < 2e-2
        assert rel_err(conv.biases.data, on.layers[0].b) < 2e-2
    finally:
        okl.set_precision("fp32")


def test_indexed_rows_match_gathered_copies(okl):
    """The ``*_rows_i32`` arguments (gather-free batches): reading dataset rows through an index gives exactly the bits of gathering
    the rows first - fp32 -> bf16 channels-last conversion, MSE against indexed targets, small-K convolution on indexed inputs."""
    from okrolearn_b200 import _lib
    from okrolearn_b200.tensor import _DevBuf
    lib, check, ctx = _lib.lib, _lib.check, okl.get_ctx()
    rng = np.random.default_rng(21)
    NS, B = 11, 4
    rows = np.array([7, 2, 10, 2], dtype=np.int32)
    idx = _DevBuf(ctx, B * 4)
    check(lib.okl_h2d_i32(ctx.h, idx.ptr, rows.astype(np.int64).ctypes.data, B, _lib.I64))
    # conversion
    X = rng.standard_normal((NS, 16, 37)).astype(np.float32)
    tX, tXg = okl.Tensor(X), okl.Tensor(X[rows])
    a, b = _DevBuf(ctx, B * 16 * 37 * 2), _DevBuf(ctx, B * 16 * 37 * 2)
    check(lib.okl_to_bf16_nxc(ctx.h, tX._dptr(_lib.NCX), idx.ptr, a.ptr, B, 16, 37))
    check(lib.okl_to_bf16_nxc(ctx.h, tXg._dptr(_lib.NCX), None, b.ptr, B, 16, 37))
    assert np.array_equal(fetch_bf16(lib, check, ctx, a.ptr, B * 16 * 37), fetch_bf16(lib, check, ctx, b.ptr, B * 16 * 37))
    # MSE with indexed targets
    o = bf16_round(rng.standard_normal((B, 16, 37)).astype(np.float32))
    to = okl.Tensor(o)
    l1, l2, g1, g2 = _DevBuf(ctx, 4), _DevBuf(ctx, 4), _DevBuf(ctx, o.size * 2), _DevBuf(ctx, o.size * 2)
    check(lib.okl_mse_fwd_bwd_nxc_bf16(ctx.h, to._b16(_lib.NXC), tX._dptr(_lib.NCX), idx.ptr, l1.ptr, g1.ptr, B, 16, 37, float(o.size), None, 0))
    check(lib.okl_mse_fwd_bwd_nxc_bf16(ctx.h, to._b16(_lib.NXC), tXg._dptr(_lib.NCX), None, l2.ptr, g2.ptr, B, 16, 37, float(o.size), None, 0))
    lv = np.empty(2, dtype=np.float32)
    check(lib.okl_d2h_async_raw(ctx.h, lv.ctypes.data, l1.ptr, 4))
    check(lib.okl_d2h_async_raw(ctx.h, lv.ctypes.data + 4, l2.ptr, 4))
    ctx.sync()
    assert lv[0] == lv[1]
    assert np.array_equal(fetch_bf16(lib, check, ctx, g1.ptr, o.size), fetch_bf16(lib, check, ctx, g2.ptr, o.size))
    # small-K convolution (2-D and 3-D) on indexed inputs
    for shape in ((NS, 3, 12, 16), (NS, 3, 4, 9, 16)):
        nd = len(shape) - 2
        Xc = rng.standard_normal(shape).astype(np.float32)
        W = okl.Tensor((rng.standard_normal((64, 3) + (3,) * nd) * 0.2).astype(np.float32))
        bb = okl.Tensor(np.zeros(64, dtype=np.float32))
        d = _lib.ConvDesc()
        d.nd, d.N, d.C, d.O = nd, B, 3, 64
        for i in range(nd):
            d.inp[i], d.k[i], d.s[i], d.p[i] = shape[2 + i], 3, 1, 1
        d.act, d.x_layout, d.y_layout = 1, _lib.NCX, _lib.NXC
        assert lib.okl_conv_first_supported(ctx.h, C.byref(d))
        tC, tCg = okl.Tensor(Xc), okl.Tensor(Xc[rows])
        n_out = B * 64 * int(np.prod(shape[2:]))
        y1, y2 = _DevBuf(ctx, n_out * 2), _DevBuf(ctx, n_out * 2)
        check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), tC._dptr(_lib.NCX), idx.ptr, W._dptr(), bb._dptr(), y1.ptr))
        check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), tCg._dptr(_lib.NCX), None, W._dptr(), bb._dptr(), y2.ptr))
        assert np.array_equal(fetch_bf16(lib, check, ctx, y1.ptr, n_out), fetch_bf16(lib, check, ctx, y2.ptr, n_out))
        d.act = 0
        dW1, dW2 = okl.Tensor._empty((64, 3) + (3,) * nd, np.float32), okl.Tensor._empty((64, 3) + (3,) * nd, np.float32)
        db1, db2 = okl.Tensor._empty((64,), np.float32), okl.Tensor._empty((64,), np.float32)
        check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), tC._dptr(_lib.NCX), idx.ptr, y1.ptr, dW1._ptr_raw(), db1._ptr_raw()))
        check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), tCg._dptr(_lib.NCX), None, y1.ptr, dW2._ptr_raw(), db2._ptr_raw()))
        for t in (dW1, dW2, db1, db2):
            t._touch()
        assert np.array_equal(dW1.data, dW2.data) and np.array_equal(db1.data, db2.data)


def test_gather_free_train_follows_a_new_dataset(okl):
    """A captured step that reads the dataset in place must not survive a change of dataset: train on A, then on B (same shapes) with
    the same network; the second run has to see B (the oracle does)."""
    okl.set_precision("bf16")
    try:
        rng = np.random.default_rng(31)
        N, B = 8, 2
        shape = (N, 3, 4, 10, 12)
        W = (rng.standard_normal((64, 3, 3, 3, 3)) * 0.1).astype(np.float32)
        net = okl.NeuralNetwork()
        conv = okl.Conv3DLayer(3, 64, 3, 1, 1)
        conv.filters.data = W.copy()
        net.add(conv)
        net.add(okl.ReLUActivationLayer())
        on = O.OracleNetwork([O.OracleConv(W.astype(np.float64), np.zeros((64, 1)), 1, 1, fast=True), O.OracleReLU()])
        for seed in (1, 2):
            X = rng.standard_normal(shape).astype(np.float32)
            X.reshape(-1)[0] = abs(X.reshape(-1)[0]) + 0.1
            T = rng.standard_normal((N, 64) + shape[2:]).astype(np.float32)
            np.random.seed(seed)
            ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=1, lr_scheduler=okl.LearningRateScheduler(0.05),
                           optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.MSELoss())
            np.random.seed(seed)
            idx = np.arange(N)
            np.random.shuffle(idx)
            Xo, To = X[idx], T[idx]
            ol = sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleMSE(), 0.05) for s in range(0, N, B)) / (N // B)
            assert rel_err(ls, [ol]) < 1e-2
        assert rel_err(conv.filters.data, on.layers[0].W) < 2e-2
    finally:
        okl.set_precision("fp32")


@pytest.mark.parametrize("graph", [True, False])
def test_bf16_conv3d_stack_train_vs_oracle(okl, graph):
    """A Conv3D stack whose inner layers run on the BF16 tensor-core kernels - Conv3D(3, 16) [small-K kernels] -> ReLU -> Conv3D(16, 32)
    -> ReLU -> Conv3D(32, 8) + MSE - through NeuralNetwork.train (fusion planner, CUDA-graph replay, auxiliary lanes) against the oracle
    network on the same shuffled batches; loss and every layer's updated parameters within the BF16 tolerance."""
    okl.set_precision("bf16")
    try:
        rng = np.random.default_rng(21)
        N, B, shape = 8, 4, (3, 4, 16, 16)
        X = rng.standard_normal((N,) + shape).astype(np.float32)
        X.reshape(-1)[0] = abs(X.reshape(-1)[0]) + 0.1
        T = rng.standard_normal((N, 8) + shape[1:]).astype(np.float32)
        Ws = [(rng.standard_normal((16, 3, 3, 3, 3)) * 0.2).astype(np.float32), (rng.standard_normal((32, 16, 3, 3, 3)) * 0.07).astype(np.float32),
              (rng.standard_normal((8, 32, 3, 3, 3)) * 0.05).astype(np.float32)]
        net = okl.NeuralNetwork()
        net.use_graph = graph
        convs = [okl.Conv3DLayer(3, 16, 3, 1, 1), okl.Conv3DLayer(16, 32, 3, 1, 1), okl.Conv3DLayer(32, 8, 3, 1, 1)]
        for c, W in zip(convs, Ws):
            c.filters.data = W.copy()
        for l in (convs[0], okl.ReLUActivationLayer(), convs[1], okl.ReLUActivationLayer(), convs[2]):
            net.add(l)
        np.random.seed(5)
        ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.02),
                       optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.MSELoss())
        assert convs[1]._conv2 and convs[2]._conv2, "the inner Conv3D layers were not taken by the tcgen05 kernels"
        z = lambda n: np.zeros((n, 1))                                   # noqa: E731
        on = O.OracleNetwork([O.OracleConv(Ws[0].astype(np.float64), z(16), 1, 1, fast=True), O.OracleReLU(),
                              O.OracleConv(Ws[1].astype(np.float64), z(32), 1, 1, fast=True), O.OracleReLU(),
                              O.OracleConv(Ws[2].astype(np.float64), z(8), 1, 1, fast=True)])
        np.random.seed(5)
        Xo, To, ol = X.copy(), T.copy(), []
        for ep in range(2):
            idx = np.arange(N)
            np.random.shuffle(idx)
            Xo, To = Xo[idx], To[idx]
            ol.append(sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleMSE(), 0.02) for s in range(0, N, B)) / (N // B))
        assert rel_err(ls, ol) < 2e-2
        for c, k in zip(convs, (0, 2, 4)):
            assert rel_err(c.filters.data, on.layers[k].W) < 2e-2
            assert rel_err(c.biases.data, on.layers[k].b) < 5e-2
    finally:
        okl.set_precision("fp32")
