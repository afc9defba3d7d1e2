"""GPU tests of the train-step loop (NeuralNetwork) against the reference-generated golden run, the oracle network,
and size-independent properties at BASELINE.json's full configuration sizes."""
import logging
import os
import tempfile

import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import okl_oracle as O

pytestmark = pytest.mark.gpu
logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)
TOL32 = 1e-5


@pytest.fixture(autouse=True)
def _fp32(okl):
    okl.set_precision("fp32")
    yield


def dense_net(okl, g):
    net = okl.NeuralNetwork(temperature=1.0)
    d1, d2 = okl.DenseLayer(6, 8), okl.DenseLayer(8, 3)
    d1.weights.data, d2.weights.data = g["Wa"].copy(), g["Wb"].copy()
    net.add(d1)
    net.add(okl.TanhActivationLayer())
    net.add(d2)
    net.add(okl.SoftmaxActivationLayer())
    return net, d1, d2


@pytest.mark.parametrize("mode", ["graph", "eager", "hooks", "stream"])
def test_train_dense_matches_reference_run(okl, mode):
    """tests/golden/train_dense.npz was produced by the shipped reference NeuralNetwork.train (3 epochs, batch 10 of 25
    samples incl. a short last batch, StepLR, CE loss, host shuffle with np.random.seed(11))."""
    g = golden("train_dense")
    net, d1, d2 = dense_net(okl, g)
    net.use_graph = mode in ("graph", "stream")
    net.stream_inputs = mode == "stream"
    seen = []
    if mode == "hooks":
        net.add_hook("post_forward", lambda out: seen.append(out.shape))
        net.add_hook("post_epoch", lambda e, n, l: seen.append(("epoch", e, round(l, 3))))
    X, T = okl.Tensor(g["X"].copy()), okl.Tensor(g["T"].copy())
    np.random.seed(int(g["seed"]))
    losses = net.train(X, T, epochs=3, lr_scheduler=okl.StepLRScheduler(0.05, 1, 0.5), optimizer=okl.SGDOptimizer(lr=0.05),
                       batch_size=10, loss_function=okl.CrossEntropyLoss())
    assert rel_err(losses, g["losses"]) < 2e-5
    assert rel_err(d1.weights.data, g["W1"]) < TOL32 and rel_err(d1.biases.data, g["b1"]) < 2e-5
    assert rel_err(d2.weights.data, g["W2"]) < TOL32 and rel_err(d2.biases.data, g["b2"]) < 2e-5
    # the caller's tensors are reordered in place, as okl.py:2959-2960 does
    assert rel_err(X.data, g["X_after"]) < 1e-6 and np.array_equal(np.asarray(T.data).astype(int), g["T_after"])
    if mode == "hooks":
        assert (10, 3) in seen and (5, 3) in seen and ("epoch", 2, round(float(g["losses"][2]), 3)) in seen


def cnn_small(okl, g):
    net = okl.NeuralNetwork()
    conv = okl.Conv2DLayer(1, 4, 3)
    conv.filters.data = g["Wc1"].copy()
    d1, d2 = okl.DenseLayer(100, 16), okl.DenseLayer(16, 10)
    d1.weights.data, d2.weights.data = g["Wd1"].copy(), g["Wd2"].copy()
    for l in (conv, okl.ReLUActivationLayer(), okl.MaxPoolingLayer(2, 2), okl.FlattenLayer(), d1, okl.ReLUActivationLayer(), d2,
              okl.SoftmaxActivationLayer()):
        net.add(l)
    return net, conv, d1, d2


@pytest.mark.parametrize("fuse", [False, True])
def test_cnn_two_steps_layer_api(okl, fuse):
    """MNIST-shaped CNN (Conv2D->ReLU->MaxPool->Flatten->Dense->ReLU->Dense->Softmax + CE), two steps through the
    reference-shaped forward/backward calls; golden from the oracle (repaired conv/pool backward) cross-checked with torch."""
    g = golden("train_cnn_small")
    net, conv, d1, d2 = cnn_small(okl, g)
    net.fuse_activations = fuse
    ce = okl.CrossEntropyLoss()
    losses = []
    for _ in range(2):
        out = net._forward(okl.Tensor(g["X"]), _fast=fuse)
        losses.append(float(ce.forward(out, okl.Tensor(g["T"]))))
        net._backward(ce.backward(out, None), float(g["lr"]))
    assert rel_err(losses, g["losses"]) < 2e-5
    assert rel_err(conv.filters.data, g["conv_W"]) < TOL32 and rel_err(conv.biases.data, g["conv_b"]) < 2e-5
    assert rel_err(d1.weights.data, g["d1_W"]) < TOL32 and rel_err(d2.weights.data, g["d2_W"]) < TOL32
    assert rel_err(d1.biases.data, g["d1_b"]) < 2e-5 and rel_err(d2.biases.data, g["d2_b"]) < 2e-5


def test_cnn_train_graph_equals_eager_and_oracle(okl):
    """Full-batch train() with CUDA-graph replay == eager == oracle network, on an MNIST-shaped CNN (batch 16, 6 steps)."""
    rng = np.random.default_rng(21)
    N, B = 48, 16
    X = rng.standard_normal((N, 1, 12, 12)).astype(np.float32)
    T = rng.integers(0, 10, N)
    Wc, W1, W2 = rng.standard_normal((4, 1, 3, 3)) * 0.3, rng.standard_normal((100, 16)) * 0.1, rng.standard_normal((16, 10)) * 0.1
    g = {"Wc1": Wc, "Wd1": W1, "Wd2": W2}
    results = {}
    for mode in ("graph", "eager"):
        net, conv, d1, d2 = cnn_small(okl, g)
        net.use_graph = mode == "graph"
        np.random.seed(5)
        ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.05),
                       optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.CrossEntropyLoss())
        results[mode] = (ls, conv.filters.data.copy(), d1.weights.data.copy(), d2.weights.data.copy())
    on = O.OracleNetwork([O.OracleConv(Wc, np.zeros((4, 1)), fast=True), O.OracleReLU(), O.OracleMaxPool(2, 2), O.OracleFlatten(),
                          O.OracleDense(W1, np.zeros((1, 16))), O.OracleReLU(), O.OracleDense(W2, np.zeros((1, 10))),
                          O.OracleSoftmax()])
    np.random.seed(5)
    Xo, To, ol = X.astype(np.float64), T.copy(), []
    for ep in range(2):
        idx = np.arange(N)
        np.random.shuffle(idx)
        Xo, To = Xo[idx], To[idx]
        ol.append(sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleCrossEntropy(), 0.05) for s in range(0, N, B)) / (N // B))
    for mode in ("graph", "eager"):
        ls, cw, w1, w2 = results[mode]
        assert rel_err(ls, ol) < 2e-5, mode
        assert rel_err(cw, on.layers[0].W) < TOL32 and rel_err(w1, on.layers[4].W) < TOL32 and rel_err(w2, on.layers[6].W) < TOL32
    assert rel_err(results["graph"][2], results["eager"][2]) < 1e-6


def test_mse_regression_net_and_temperature(okl):
    rng = np.random.default_rng(8)
    X, Y = rng.standard_normal((32, 5)), rng.standard_normal((32, 2))
    W1, W2 = rng.standard_normal((5, 7)) * 0.3, rng.standard_normal((7, 2)) * 0.3
    net = okl.NeuralNetwork(temperature=2.0)
    a, b = okl.DenseLayer(5, 7), okl.DenseLayer(7, 2)
    a.weights.data, b.weights.data = W1.copy(), W2.copy()
    net.add(a)
    net.add(okl.SigmoidActivationLayer())
    net.add(b)
    np.random.seed(1)
    ls = net.train(okl.Tensor(X.copy()), okl.Tensor(Y.copy()), 2, okl.LearningRateScheduler(0.1), okl.AdamOptimizer(), 8, okl.MSELoss())
    on = O.OracleNetwork([O.OracleDense(W1, np.zeros((1, 7))), O.OracleSigmoid(), O.OracleDense(W2, np.zeros((1, 2)))], temperature=2.0)
    np.random.seed(1)
    Xo, Yo, ol = X.copy(), Y.copy(), []
    for ep in range(2):
        idx = np.arange(32)
        np.random.shuffle(idx)
        Xo, Yo = Xo[idx], Yo[idx]
        ol.append(sum(on.train_step(Xo[s:s + 8], Yo[s:s + 8], O.OracleMSE(), 0.1) for s in range(0, 32, 8)) / 4)
    assert rel_err(ls, ol) < 2e-5
    assert rel_err(a.weights.data, on.layers[0].W) < TOL32 and rel_err(b.weights.data, on.layers[2].W) < TOL32
    out = net.forward(okl.Tensor(X))
    assert rel_err(out.data, on.forward(X)) < TOL32              # final output divided by temperature (okl.py:2890)


def test_save_load_roundtrip_and_debug_mode(okl):
    rng = np.random.default_rng(9)
    net = okl.NeuralNetwork()
    net.add(okl.Conv2DLayer(1, 2, 3))
    net.add(okl.ReLUActivationLayer())
    net.add(okl.FlattenLayer())
    net.add(okl.DenseLayer(2 * 4 * 4, 3))
    x = okl.Tensor(rng.standard_normal((2, 1, 6, 6)))
    y0 = net.forward(x).data.copy()
    with tempfile.TemporaryDirectory() as d:
        net.save(os.path.join(d, "m.pkl"))
        net2 = okl.NeuralNetwork()
        net2.add(okl.Conv2DLayer(1, 2, 3))
        net2.add(okl.ReLUActivationLayer())
        net2.add(okl.FlattenLayer())
        net2.add(okl.DenseLayer(2 * 4 * 4, 3))
        net2.load(os.path.join(d, "m.pkl"))
        assert rel_err(net2.forward(x).data, y0) < 1e-6
    net.set_debug_mode(True)
    net.forward(x)
    net.backward(okl.Tensor(np.ones((2, 3))), 0.01)
    assert len(net.get_layer_outputs()) == 4 and len(net.get_gradients()) == 4 and len(net.parameter_history) == 1


def test_full_size_mnist_cnn_step_properties(okl):
    """BASELINE config 2 at full size (batch 256): properties that need no CPU reference at that size -
    probabilities sum to 1, the loss decreases under repeated steps on one batch, update linearity
    (W2 - W1 == 2 (W1 - W0) + lr * (g2 - g1) is implied by the accumulate rule: checked as gacc consistency)."""
    rng = np.random.default_rng(1234)
    X = rng.standard_normal((256, 1, 28, 28)).astype(np.float32)
    T = rng.integers(0, 10, 256)
    np.random.seed(0)
    net = okl.NeuralNetwork()
    conv, d1, d2 = okl.Conv2DLayer(1, 32, 3), okl.DenseLayer(5408, 128), okl.DenseLayer(128, 10)
    for l in (conv, okl.ReLUActivationLayer(), okl.MaxPoolingLayer(2, 2), okl.FlattenLayer(), d1, okl.ReLUActivationLayer(), d2,
              okl.SoftmaxActivationLayer()):
        net.add(l)
    W0 = d1.weights.data.copy()
    ls = net.train(okl.Tensor(X), okl.Tensor(T), epochs=4, lr_scheduler=okl.LearningRateScheduler(0.01),
                   optimizer=okl.SGDOptimizer(), batch_size=256, loss_function=okl.CrossEntropyLoss())
    assert all(np.isfinite(ls)) and ls[-1] < ls[0]
    p = net.forward(okl.Tensor(X)).data
    assert p.shape == (256, 10) and np.allclose(p.sum(1), 1.0, atol=1e-5) and (p >= 0).all()
    # accumulated rule: W4 = W0 - lr * (4 g1 + 3 g2 + 2 g3 + g4)  =>  gacc = g1+g2+g3+g4 is finite and non-trivial
    assert np.isfinite(d1.weights.grad).all() and np.abs(d1.weights.grad).max() > 0
    assert np.abs(d1.weights.data - W0).max() > 0
    # one oracle-checked slice at full width: the first 8 samples through the same weights
    on = O.OracleNetwork([O.OracleConv(conv.filters.data, conv.biases.data, fast=True), O.OracleReLU(), O.OracleMaxPool(2, 2),
                          O.OracleFlatten(), O.OracleDense(d1.weights.data, d1.biases.data), O.OracleReLU(),
                          O.OracleDense(d2.weights.data, d2.biases.data), O.OracleSoftmax()])
    assert rel_err(p[:8], on.forward(X[:8].astype(np.float64))) < 1e-4


@pytest.mark.parametrize("graph", [True, False])
def test_fused_classifier_head_equals_unfused_and_oracle(okl, graph):
    """Dense(N<=16) -> Softmax -> CrossEntropy as ONE kernel (okl_dense_softmax_ce) must give the same parameters as the separate
    dense / CE / dense-backward kernels (same summation orders => bit-identical) and match the oracle; short last batch included."""
    rng = np.random.default_rng(33)
    N, B = 56, 16                                   # 3 full batches + one of 8 rows
    X = rng.standard_normal((N, 20)).astype(np.float32)
    T = rng.integers(0, 10, N)
    W1, W2 = rng.standard_normal((20, 16)) * 0.3, rng.standard_normal((16, 10)) * 0.3
    res = {}
    for fuse in (True, False):
        net = okl.NeuralNetwork()
        d1, d2 = okl.DenseLayer(20, 16), okl.DenseLayer(16, 10)
        d1.weights.data, d2.weights.data = W1.copy(), W2.copy()
        for l in (d1, okl.ReLUActivationLayer(), d2, okl.SoftmaxActivationLayer()):
            net.add(l)
        net.fuse_head, net.use_graph = fuse, graph
        np.random.seed(3)
        ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.1),
                       optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.CrossEntropyLoss())
        assert d2._fused_head == fuse
        res[fuse] = (np.array(ls), d1.weights.data.copy(), d1.biases.data.copy(), d2.weights.data.copy(), d2.biases.data.copy())
    for a, b in zip(res[True][1:], res[False][1:]):
        assert np.array_equal(a, b)
    assert rel_err(res[True][0], res[False][0]) < 1e-6
    on = O.OracleNetwork([O.OracleDense(W1, np.zeros((1, 16))), O.OracleReLU(), O.OracleDense(W2, np.zeros((1, 10))), O.OracleSoftmax()])
    np.random.seed(3)
    Xo, To, ol = X.astype(np.float64), T.copy(), []
    for ep in range(2):
        idx = np.arange(N)
        np.random.shuffle(idx)
        Xo, To = Xo[idx], To[idx]
        ol.append(sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleCrossEntropy(), 0.1) for s in range(0, N, B)) / 4)
    assert rel_err(res[True][0], ol) < 2e-5
    assert rel_err(res[True][1], on.layers[0].W) < TOL32 and rel_err(res[True][3], on.layers[2].W) < TOL32
    assert rel_err(res[True][2], on.layers[0].b) < 2e-5 and rel_err(res[True][4], on.layers[2].b) < 2e-5


def test_fused_head_output_still_readable(okl):
    """The head's probabilities are materialised on demand when something other than CrossEntropyLoss reads them first."""
    rng = np.random.default_rng(4)
    X = rng.standard_normal((8, 12))
    net = okl.NeuralNetwork()
    d = okl.DenseLayer(12, 5)
    net.add(d)
    net.add(okl.SoftmaxActivationLayer())
    out = net._forward(okl.Tensor(X), _fast=True)
    assert d._fused_head
    z = X @ d.weights.data + d.biases.data
    assert rel_err(out.data, O.softmax_forward(z)) < TOL32


def test_pipelined_train_loop_equals_inline_loop(okl):
    """Double-buffered batch prefetch (copy stream) + lagged per-step loss read-back must give exactly what the inline loop
    (gather on the compute stream, blocking loss read) gives: same parameters bit for bit, same per-step losses, including a short
    last batch and two epochs (slot parity crosses the epoch boundary); host-resident (pinned) and device-resident datasets."""
    rng = np.random.default_rng(17)
    N, B = 88, 16                                   # 5 full batches + one of 8 rows
    X = rng.standard_normal((N, 1, 12, 12)).astype(np.float32)
    T = rng.integers(0, 10, N)
    g = {"Wc1": rng.standard_normal((4, 1, 3, 3)) * 0.3, "Wd1": rng.standard_normal((100, 16)) * 0.1, "Wd2": rng.standard_normal((16, 10)) * 0.1}
    res = {}
    for mode in ("inline", "prefetch", "prefetch_stream"):
        net, conv, d1, d2 = cnn_small(okl, g)
        net.prefetch_batches = mode != "inline"
        net.stream_inputs = mode == "prefetch_stream"
        net.sync_loss_every_step = True
        np.random.seed(11)
        if mode == "prefetch_stream":
            pX = okl.pinned_empty(X.shape, np.float32)
            pX[...] = X
            tx = okl.Tensor.wrap(pX)
        else:
            tx = okl.Tensor(X.copy())
        ls = net.train(tx, okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.05), optimizer=okl.SGDOptimizer(),
                       batch_size=B, loss_function=okl.CrossEntropyLoss())
        assert len(net.last_step_losses) == 12      # every step of both epochs
        res[mode] = (np.array(ls), np.array(net.last_step_losses), conv.filters.data.copy(), d1.weights.data.copy(), d2.weights.data.copy())
    for mode in ("prefetch", "prefetch_stream"):
        for a, b in zip(res[mode], res["inline"]):
            assert np.array_equal(a, b), mode
    assert abs(res["inline"][1][6:].mean() - res["inline"][0][-1]) < 1e-6  # epoch loss = mean of its step losses


@pytest.mark.parametrize("nd,fuse", [(3, True), (3, False), (1, True)])
def test_conv_relu_mse_regression_train(okl, nd, fuse):
    """Conv3D(3, 8, 3, 1, 1) / Conv1D(3, 8, 3, 1, 1) + ReLU trained against a random target with MSELoss (BASELINE configs[3] shape
    family).  Fast path: ReLU in the conv epilogue, its backward inside the MSE kernel, tap-sliced small-K weight gradient (K + 1 =
    82 > one 28-tap slice for the 3-D layer); must match the oracle network and the unfused path."""
    okl.set_precision("fp32")
    rng = np.random.default_rng(50 + nd)
    sp = (5, 6, 7) if nd == 3 else (40,)
    N, B = 6, 3
    X = rng.standard_normal((N, 3) + sp).astype(np.float32)
    X.reshape(-1)[0] = abs(X.reshape(-1)[0]) + 0.1
    T = rng.standard_normal((N, 8) + sp).astype(np.float32)
    W = (rng.standard_normal((8, 3) + (3,) * nd) * 0.2).astype(np.float32)
    net = okl.NeuralNetwork()
    conv = okl.Conv3DLayer(3, 8, 3, 1, 1) if nd == 3 else okl.Conv1DLayer(3, 8, 3, 1, 1)
    (conv.filters if nd == 3 else conv.weights).data = W.copy()
    net.add(conv)
    net.add(okl.ReLUActivationLayer())
    net.fuse_activations = fuse
    np.random.seed(2)
    ls = net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=2, lr_scheduler=okl.LearningRateScheduler(0.05),
                   optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.MSELoss())
    on = O.OracleNetwork([O.OracleConv(W, np.zeros((8, 1)), 1, 1, fast=True), O.OracleReLU()])
    np.random.seed(2)
    Xo, To, ol = X.copy(), T.copy(), []
    for ep in range(2):
        idx = np.arange(N)
        np.random.shuffle(idx)
        Xo, To = Xo[idx], To[idx]
        ol.append(sum(on.train_step(Xo[s:s + B], To[s:s + B], O.OracleMSE(), 0.05) for s in range(0, N, B)) / 2)
    assert rel_err(ls, ol) < 2e-5
    assert rel_err((conv.filters if nd == 3 else conv.weights).data, on.layers[0].W) < 5e-5
    assert rel_err(conv.biases.data, on.layers[0].b) < 5e-5


def test_checkpoint_resume_is_bit_exact(okl):
    """save_checkpoint / load_checkpoint keep the accumulated gradients, so (1 epoch, checkpoint, fresh network,
    1 epoch) ends with exactly the parameters of 2 uninterrupted epochs - which ``save`` / ``load`` (the reference's format) cannot."""
    import os
    import tempfile
    rng = np.random.default_rng(44)
    N, B = 48, 16
    X = rng.standard_normal((N, 1, 12, 12)).astype(np.float32)
    T = rng.integers(0, 10, N)
    g = {"Wc1": rng.standard_normal((4, 1, 3, 3)) * 0.3, "Wd1": rng.standard_normal((100, 16)) * 0.1, "Wd2": rng.standard_normal((16, 10)) * 0.1}

    def run(net, x, t, epochs):
        return net.train(x, t, epochs=epochs, lr_scheduler=okl.LearningRateScheduler(0.05), optimizer=okl.SGDOptimizer(), batch_size=B,
                         loss_function=okl.CrossEntropyLoss())

    netA, convA, d1A, d2A = cnn_small(okl, g)
    np.random.seed(6)
    run(netA, okl.Tensor(X.copy()), okl.Tensor(T.copy()), 2)
    ref_w = [convA.filters.data.copy(), convA.biases.data.copy(), d1A.weights.data.copy(), d2A.weights.data.copy(), d2A.biases.data.copy()]

    with tempfile.TemporaryDirectory() as d:
        for mode in ("checkpoint", "save"):
            netB, *_ = cnn_small(okl, g)
            netC, convC, d1C, d2C = cnn_small(okl, g)         # built BEFORE seeding: layer constructors draw from np.random
            xs, ts = okl.Tensor(X.copy()), okl.Tensor(T.copy())
            np.random.seed(6)
            run(netB, xs, ts, 1)
            path = os.path.join(d, mode + ".pkl")
            (netB.save_checkpoint if mode == "checkpoint" else netB.save)(path)
            (netC.load_checkpoint if mode == "checkpoint" else netC.load)(path)
            run(netC, xs, ts, 1)                      # same (reordered) dataset tensors, NumPy RNG stream continues
            got = [convC.filters.data, convC.biases.data, d1C.weights.data, d2C.weights.data, d2C.biases.data]
            if mode == "checkpoint":
                for a, b in zip(got, ref_w):
                    assert np.array_equal(a, b)
            else:                                     # the reference's format drops the accumulators: the resumed run differs
                assert not np.array_equal(got[2], ref_w[2])
        with pytest.raises(ValueError):
            netC.load_checkpoint(os.path.join(d, "save.pkl"))


def test_host_side_parameter_edits_reach_replayed_graphs(okl):
    """train() (graphs captured) -> the user resets the parameters and clears the accumulated gradients on the host -> train() again
    with a FRESH loss object (reuses the captured step) must equal a fresh network trained once: replay must see the edits."""
    rng = np.random.default_rng(71)
    N, B = 64, 16
    X = rng.standard_normal((N, 1, 12, 12)).astype(np.float32)
    T = rng.integers(0, 10, N)
    g = {"Wc1": rng.standard_normal((4, 1, 3, 3)) * 0.3, "Wd1": rng.standard_normal((100, 16)) * 0.1, "Wd2": rng.standard_normal((16, 10)) * 0.1}

    def run(net):
        np.random.seed(13)
        return net.train(okl.Tensor(X.copy()), okl.Tensor(T.copy()), epochs=1, lr_scheduler=okl.LearningRateScheduler(0.05),
                         optimizer=okl.SGDOptimizer(), batch_size=B, loss_function=okl.CrossEntropyLoss())

    netA, convA, d1A, d2A = cnn_small(okl, g)
    netB, convB, d1B, d2B = cnn_small(okl, g)
    lossA1 = run(netA)
    n_graphs = len(netA._graphs)
    convA.filters.data, d1A.weights.data, d2A.weights.data = g["Wc1"].copy(), g["Wd1"].copy(), g["Wd2"].copy()
    convA.biases.data, d1A.biases.data, d2A.biases.data = np.zeros((4, 1)), np.zeros((1, 16)), np.zeros((1, 10))
    for t in (convA.filters, convA.biases, d1A.weights, d1A.biases, d2A.weights, d2A.biases):
        t.grad = None
    lossA2 = run(netA)
    lossB = run(netB)
    assert len(netA._graphs) == n_graphs                       # the fresh CrossEntropyLoss() reused the captured step
    assert np.array_equal(np.array(lossA2), np.array(lossB)) and np.array_equal(np.array(lossA1), np.array(lossB))
    for a, b in ((convA.filters, convB.filters), (d1A.weights, d1B.weights), (d2A.weights, d2B.weights), (d2A.biases, d2B.biases)):
        assert np.array_equal(a.data, b.data)


# ---------------------------------------------------------------------------------------------- round-2 regressions
def _small_dense(okl, seed=3):
    rng = np.random.default_rng(seed)
    g = {"Wa": rng.standard_normal((6, 8)) * 0.3, "Wb": rng.standard_normal((8, 3)) * 0.3}
    return dense_net(okl, g) + (g,)


def test_device_consumers_see_the_reordered_dataset_after_train(okl):
    """okl.py:2959-2960 permutes inputs.data / targets.data eagerly, so ``net.forward(X)`` after ``train`` lines up with
    ``T.data``.  Here the reorder is pending until first use - and a DEVICE consumer (forward) must honour it as well as .data."""
    net, d1, d2, g = _small_dense(okl)
    rng = np.random.default_rng(0)
    X0, T0 = rng.standard_normal((25, 6)), rng.integers(0, 3, 25)
    X, T = okl.Tensor(X0.copy()), okl.Tensor(T0.copy())
    np.random.seed(4)
    net.train(X, T, 2, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 10, okl.CrossEntropyLoss())
    out_dev = net.forward(X).data                     # device consumer first: no .data read of X before it
    np.random.seed(4)
    perm = np.arange(25)
    for _ in range(2):
        idx = np.arange(25)
        np.random.shuffle(idx)
        perm = perm[idx]
    assert rel_err(X.data, X0[perm]) < 1e-6 and np.array_equal(np.asarray(T.data).astype(int), T0[perm])
    on = O.OracleNetwork([O.OracleDense(d1.weights.data, d1.biases.data), O.OracleTanh(), O.OracleDense(d2.weights.data, d2.biases.data),
                          O.OracleSoftmax()])
    assert rel_err(out_dev, on.forward(X0[perm])) < 2e-5
    # predictions and labels are paired the way the reference pairs them
    assert out_dev.shape[0] == np.asarray(T.data).shape[0]


def test_data_setter_drops_a_pending_reorder(okl):
    net, d1, d2, g = _small_dense(okl)
    rng = np.random.default_rng(1)
    X, T = okl.Tensor(rng.standard_normal((25, 6))), okl.Tensor(rng.integers(0, 3, 25))
    net.train(X, T, 1, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 10, okl.CrossEntropyLoss())
    new = rng.standard_normal((12, 6))                # smaller dataset on the same Tensor objects
    X.data = new
    assert np.array_equal(X.data, new)                # not shuffled by the stale permutation
    T.data = rng.integers(0, 3, 12)
    losses = net.train(X, T, 1, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 5, okl.CrossEntropyLoss())
    assert np.isfinite(losses).all()
    # only ONE of the two replaced: the other's pending order must be applied, not reused for both
    X2, T2 = okl.Tensor(rng.standard_normal((20, 6))), okl.Tensor(rng.integers(0, 3, 20))
    net.train(X2, T2, 1, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 10, okl.CrossEntropyLoss())
    t_after = np.asarray(T2.data).copy()
    X2.data = rng.standard_normal((20, 6))
    keep = np.asarray(X2.data).copy()
    np.random.seed(8)
    net.train(X2, T2, 1, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 10, okl.CrossEntropyLoss())
    np.random.seed(8)
    idx = np.arange(20)
    np.random.shuffle(idx)
    assert rel_err(X2.data, keep[idx]) < 1e-6 and np.array_equal(np.asarray(T2.data), t_after[idx])


def test_temperature_change_is_not_served_from_a_stale_graph(okl):
    rng = np.random.default_rng(2)
    X0, T0 = rng.standard_normal((20, 6)), rng.integers(0, 3, 20)
    res = {}
    for mode in ("graph", "eager"):
        net, d1, d2, g = _small_dense(okl)
        net.use_graph = mode == "graph"
        for temp in (1.0, 2.0):
            net.set_temperature(temp)
            np.random.seed(3)
            net.train(okl.Tensor(X0.copy()), okl.Tensor(T0.copy()), 2, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 10,
                      okl.CrossEntropyLoss())
        res[mode] = (d1.weights.data.copy(), d2.weights.data.copy())
    assert rel_err(res["graph"][0], res["eager"][0]) < 1e-6 and rel_err(res["graph"][1], res["eager"][1]) < 1e-6


def test_inplace_edit_of_any_element_reaches_the_device(okl):
    """``layer.weights.data[i, j] = 0`` on the array returned by ``.data`` (the reference's live array) - every element counts."""
    lay = okl.DenseLayer(300, 200)                    # 60 000 elements: most are not on a 1024-point sample grid
    x = okl.Tensor(np.ones((1, 300)))
    lay.forward(x).data
    w = lay.weights.data
    w[137, 91] += 5.0                                 # touches neither a strided sample nor the last element
    y = lay.forward(x).data
    assert abs(y[0, 91] - (w[:, 91].sum())) < 1e-3


def test_epoch_hook_assignment_survives_graph_replay(okl):
    rng = np.random.default_rng(5)
    X0, T0 = rng.standard_normal((20, 6)), rng.integers(0, 3, 20)
    res = {}
    for mode in ("graph", "eager"):
        net, d1, d2, g = _small_dense(okl)
        net.use_graph = mode == "graph"
        Wnew = rng.standard_normal((8, 3)) * 0.2 if mode == "graph" else res["_W"]
        res["_W"] = Wnew

        def hook(epoch, epochs, _d2=d2, _W=Wnew):
            if epoch == 1:
                _d2.weights.data = _W.copy()
        net.add_hook("pre_epoch", hook)
        np.random.seed(6)
        net.train(okl.Tensor(X0.copy()), okl.Tensor(T0.copy()), 3, okl.LearningRateScheduler(0.05), okl.SGDOptimizer(lr=0.05), 10,
                  okl.CrossEntropyLoss())
        res[mode] = (d1.weights.data.copy(), d2.weights.data.copy())
    assert rel_err(res["graph"][0], res["eager"][0]) < 1e-6 and rel_err(res["graph"][1], res["eager"][1]) < 1e-6


def test_stream_mode_staged_gather_matches_resident_dataset(okl):
    """Batches of >= 1 MB in stream mode cross PCIe through the host-gathered staging ring (okl_batch_host_index + one DMA copy per part)
    instead of the zero-copy gather kernel: two epochs incl. a short last batch and a wrap of the staging ring must leave exactly the
    parameters, losses and reordered dataset of the device-resident run (same seed, same shuffles, FP32 mode: bit-identical kernels)."""
    rng = np.random.default_rng(5)
    n, feat, classes, B = 1100, 4096, 10, 100               # 100 x 4096 x 4 B = 1.6 MB per batch, 11 batches per epoch > ring of 8
    X = rng.standard_normal((n, feat)).astype(np.float32)
    T = rng.integers(0, classes, n).astype(np.int32)
    W1, W2 = rng.standard_normal((feat, 32)) * 0.02, rng.standard_normal((32, classes)) * 0.1
    results = []
    for stream in (False, True):
        net = okl.NeuralNetwork()
        d1, d2 = okl.DenseLayer(feat, 32), okl.DenseLayer(32, classes)
        d1.weights.data, d2.weights.data = W1.copy(), W2.copy()
        for l in (d1, okl.TanhActivationLayer(), d2, okl.SoftmaxActivationLayer()):
            net.add(l)
        net.stream_inputs, net.sync_loss_every_step = stream, stream
        if stream:
            pX, pT = okl.pinned_empty(X.shape, np.float32), okl.pinned_empty(T.shape, np.int32)
            pX[...], pT[...] = X, T
            tX, tT = okl.Tensor.wrap(pX), okl.Tensor.wrap(pT)
        else:
            tX, tT = okl.Tensor(X.copy()), okl.Tensor(T.copy())
        np.random.seed(3)
        losses = net.train(tX, tT, epochs=2, lr_scheduler=okl.LearningRateScheduler(0.05), optimizer=okl.SGDOptimizer(lr=0.05), batch_size=B,
                           loss_function=okl.CrossEntropyLoss())
        results.append((np.asarray(losses), d1.weights.data.copy(), d2.weights.data.copy(), np.asarray(tX.data).copy(), np.asarray(tT.data).copy(),
                        list(net.last_step_losses)))
    a, b = results
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4].astype(int), b[4].astype(int))
    assert len(b[5]) == 2 * 11                               # every step's loss reached the host
