import sys, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
import okrolearn_b200.okrolearn as okl
from oracle import okl_oracle as O
from conftest import rel_err
okl.set_precision("bf16")
for chans in ((16, 32, 48), (24, 40, 16)):
    c1n, c2n, c3n = chans
    rng = np.random.default_rng(c1n)
    N = 8
    X = rng.standard_normal((N, 3, 16, 16)).astype(np.float32)
    T = rng.integers(0, 10, N)
    W1, W2, W3 = rng.standard_normal((c1n, 3, 3, 3)) * 0.2, rng.standard_normal((c2n, c1n, 3, 3)) * 0.08, rng.standard_normal((c3n, c2n, 3, 3)) * 0.08
    W4 = rng.standard_normal((c3n * 4 * 4, 10)) * 0.05
    net = okl.NeuralNetwork()
    l1, l2, l3, d = okl.Conv2DLayer(3, c1n, 3, 1, 1), okl.Conv2DLayer(c1n, c2n, 3, 1, 1), okl.Conv2DLayer(c2n, c3n, 3, 1, 1), okl.DenseLayer(c3n * 16, 10)
    l1.filters.data, l2.filters.data, l3.filters.data, d.weights.data = W1.copy(), W2.copy(), W3.copy(), W4.copy()
    for l in (l1, okl.ReLUActivationLayer(), l2, okl.ReLUActivationLayer(), okl.MaxPoolingLayer(2, 2), l3, okl.ReLUActivationLayer(),
              okl.MaxPoolingLayer(2, 2), okl.FlattenLayer(), d, okl.SoftmaxActivationLayer()):
        net.add(l)
    z = lambda n: np.zeros((n, 1))
    on = O.OracleNetwork([O.OracleConv(W1, z(c1n), 1, 1, fast=True), O.OracleReLU(), O.OracleConv(W2, z(c2n), 1, 1, fast=True), O.OracleReLU(),
                          O.OracleMaxPool(2, 2), O.OracleConv(W3, z(c3n), 1, 1, fast=True), O.OracleReLU(), O.OracleMaxPool(2, 2), O.OracleFlatten(),
                          O.OracleDense(W4, np.zeros((1, 10))), O.OracleSoftmax()])
    ce = okl.CrossEntropyLoss()
    for fast in (False, True, True):
        out = net._forward(okl.Tensor(X), _fast=fast)
        loss = float(ce.forward(out, okl.Tensor(T)))
        net._backward(ce.backward(out, None), 0.02)
        ref = on.train_step(X.astype(np.float64), T, O.OracleCrossEntropy(), 0.02)
        print(chans, fast, 'loss', loss, ref, 'conv2', l2._conv2, l3._conv2, 'first_tc', getattr(l1, '_first_tc', None))
        for lay, k, nm in ((l1, 0, 'l1'), (l2, 2, 'l2'), (l3, 5, 'l3')):
            print('   ', nm, 'W', rel_err(lay.filters.data, on.layers[k].W), 'dW', rel_err(lay.filters.data - [W1, W2, W3][[0,2,5].index(k)], on.layers[k].W - [W1, W2, W3][[0,2,5].index(k)]), 'b', rel_err(lay.biases.data, on.layers[k].b))
        print('    dense W', rel_err(d.weights.data, on.layers[9].W), 'b', rel_err(d.biases.data, on.layers[9].b))
