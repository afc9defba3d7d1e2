"""GPU parity tests (run with ``-m gpu`` on the B200): the CUDA path, called through the C ABI by the reference-shaped
Python layers, against (1) the committed golden vectors generated from the reference and (2) the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): 1e-5 relative (max|a-b| / max|b|) in FP32 mode, 2e-2 in TF32/BF16 mode.
"""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import okl_oracle as O

pytestmark = pytest.mark.gpu

TOL32 = 1e-5
CONV_CASES = ["conv2d_cfg1_small", "conv2d_s2p1", "conv2d_rect", "conv2d_c16", "conv1d_p0", "conv1d_s2p2",
              "conv3d_p0", "conv3d_p1", "conv3d_s2"]


@pytest.fixture(autouse=True)
def _fp32(okl):
    okl.set_precision("fp32")
    yield


def make_conv(okl, g):
    W = g["W"]
    nd = W.ndim - 2
    s, p = tuple(int(v) for v in g["stride"]), tuple(int(v) for v in g["padding"])
    k = tuple(W.shape[2:])
    if nd == 1:
        lay = okl.Conv1DLayer(W.shape[1], W.shape[0], k[0], s[0], p[0])
        lay.weights.data = W.copy()
    elif nd == 2:
        lay = okl.Conv2DLayer(W.shape[1], W.shape[0], k, s, p)
        lay.filters.data = W.copy()
    else:
        lay = okl.Conv3DLayer(W.shape[1], W.shape[0], k, s, p)
        lay.filters.data = W.copy()
    lay.biases.data = g["b"].copy()
    return lay


def test_dense_golden(okl):
    g = golden("dense")
    lay = okl.DenseLayer(12, 7)
    lay.weights.data = g["W"].copy()
    lay.biases.data = g["b"].copy()
    y = lay.forward(okl.Tensor(g["x"]))
    assert y.data.dtype == np.float64 and rel_err(y.data, g["y"]) < TOL32
    dx1 = lay.backward(okl.Tensor(g["g1"]), float(g["lr"]))
    assert rel_err(dx1.data, g["dx1"]) < TOL32
    assert rel_err(lay.weights.data, g["W1"]) < TOL32 and rel_err(lay.biases.data, g["b1"]) < TOL32
    lay.forward(okl.Tensor(g["x"]))
    dx2 = lay.backward(okl.Tensor(g["g2"]), float(g["lr"]))
    assert rel_err(dx2.data, g["dx2"]) < TOL32
    # accumulated-gradient rule: W2 = W1 - lr (g1 + g2)
    assert rel_err(lay.weights.data, g["W2"]) < TOL32 and rel_err(lay.biases.data, g["b2"]) < TOL32
    assert lay.weights.grad.shape == (12, 7) and lay.biases.grad.shape == (1, 7)
    assert sorted(lay.get_params()) == ["biases", "weights"]


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv_golden(okl, name):
    g = golden(name)
    lay = make_conv(okl, g)
    y = lay.forward(okl.Tensor(g["x"]))
    assert y.data.dtype == g["y"].dtype                      # float64 for Conv2D, float32 for Conv1D/3D
    assert rel_err(y.data, g["y"]) < TOL32
    W0, b0 = g["W"].astype(np.float64), g["b"].astype(np.float64)
    dX = lay.backward(okl.Tensor(g["g"]), 0.1)
    assert rel_err(dX.data, g["dX"]) < TOL32
    Wt = lay.weights if hasattr(lay, "weights") else lay.filters
    assert rel_err(Wt.grad, g["dW"]) < TOL32
    assert lay.biases.grad.shape == (g["W"].shape[0], 1) and rel_err(lay.biases.grad, g["db"]) < TOL32
    assert rel_err(Wt.data, W0 - 0.1 * g["dW"]) < TOL32
    assert rel_err(lay.biases.data, b0 - 0.1 * g["db"]) < TOL32


def test_conv_kat_inputs_of_reference_tests(okl):
    g = golden("kat_conv_reference_tests")
    c2 = okl.Conv2DLayer(3, 2, (3, 3), stride=1, padding=0)
    c2.filters.data = g["W2"]                                 # int64 arrays, as tests/conv2d.py:11-19 assigns
    c2.biases.data = g["b2"]
    y2 = c2.forward(okl.Tensor(g["x2"])).data
    assert np.allclose(y2, g["y2"], rtol=1e-6)
    dX = c2.backward(okl.Tensor(np.ones_like(y2)), lr=0.01)   # the part of tests/conv2d.py:54-60 the reference never reaches
    assert c2.filters.grad is not None and c2.biases.grad is not None and dX.data.shape == g["x2"].shape
    c1 = okl.Conv1DLayer(3, 2, 3, stride=1, padding=0)
    c1.weights.data = g["W1"]
    c1.biases.data = g["b1"]
    assert np.allclose(c1.forward(okl.Tensor(g["x1"])).data, g["y1"], rtol=1e-6)


def test_activations_golden(okl):
    g = golden("activations")
    x, gr = g["x"], g["g"]
    r = okl.ReLUActivationLayer()
    assert rel_err(r.forward(okl.Tensor(x)).data, g["relu_y"]) < 1e-7
    xin = okl.Tensor(x)
    r.forward(xin)
    assert rel_err(r.backward(okl.Tensor(gr), 0.0).data, g["relu_dx"]) < 1e-7
    assert rel_err(xin.grad, g["relu_dx"]) < 1e-7            # okl.py:110 also sets inputs.grad
    s = okl.SigmoidActivationLayer()
    assert rel_err(s.forward(okl.Tensor(x)).data, g["sigmoid_y"]) < 1e-6
    assert rel_err(s.backward(okl.Tensor(gr)).data, g["sigmoid_dx"]) < 1e-6
    t = okl.TanhActivationLayer()
    assert rel_err(t.forward(okl.Tensor(x)).data, g["tanh_y"]) < 1e-6
    assert rel_err(t.backward(okl.Tensor(gr), 0.0).data, g["tanh_dx"]) < 1e-6
    sm = okl.SoftmaxActivationLayer()
    assert rel_err(sm.forward(okl.Tensor(x)).data, g["softmax_y"]) < 1e-6
    assert rel_err(sm.backward(okl.Tensor(gr)).data, g["softmax_dx"]) < 1e-5
    kat = okl.SoftmaxActivationLayer().forward(okl.Tensor(g["softmax_kat_x"])).data        # tests/dot_test.py:28-32
    assert np.allclose(kat, [[0.09003057, 0.24472847, 0.66524096]], atol=1e-6)
    # x.flat[0] <= 0: the shipped np.vectorize path truncates to int64; this path returns the intended max(0, x)
    xn = x.copy()
    xn.flat[0] = -0.5
    yn = okl.ReLUActivationLayer().forward(okl.Tensor(xn)).data
    assert yn.dtype == np.float64 and rel_err(yn, np.maximum(xn, 0)) < 1e-7


def test_pooling_golden(okl):
    g = golden("pooling")
    x = g["x"]
    for kind, cls in (("max", okl.MaxPoolingLayer), ("avg", okl.AveragePoolingLayer)):
        for ps, st in ((2, 2), (3, 2), (2, 1), (3, 3)):
            k = f"{kind}_{ps}_{st}_"
            lay = cls(ps, st)
            y = lay.forward(okl.Tensor(x))
            assert y.data.dtype == np.float64 and rel_err(y.data, g[k + "y"]) < 1e-6
            dx = lay.backward(okl.Tensor(g[k + "g"]))
            assert rel_err(dx.data, g[k + "dx"]) < 1e-6, k
    assert okl.MaxPoolingLayer(2, 2).forward(okl.Tensor(g["kat_x"])).data.reshape(2, 2).tolist() == g["kat_max"].tolist()
    assert okl.AveragePoolingLayer(2, 2).forward(okl.Tensor(g["kat_x"])).data.reshape(2, 2).tolist() == g["kat_avg"].tolist()
    # ties: every maximum of a window receives the gradient (R4)
    mp = okl.MaxPoolingLayer(2, 2)
    mp.forward(okl.Tensor(np.zeros((1, 1, 2, 2))))
    assert mp.backward(okl.Tensor(np.full((1, 1, 1, 1), 3.0))).data.tolist() == [[[[3.0, 3.0], [3.0, 3.0]]]]
    assert okl.MaxPoolingLayer(3, 2).forward(okl.Tensor(np.zeros((1, 2, 7, 7)))).shape == (1, 2, 3, 3)


def test_losses_golden(okl):
    g = golden("losses")
    ce = okl.CrossEntropyLoss()
    p, t = okl.Tensor(g["p"]), okl.Tensor(g["t"])
    v = ce.forward(p, t)
    assert abs(float(v) - float(g["ce"])) < 1e-5 * abs(float(g["ce"]))
    assert rel_err(ce.backward(p, t).data, g["ce_grad"]) < 1e-6
    ms = okl.MSELoss()
    o, tt = okl.Tensor(g["o"]), okl.Tensor(g["tt"])
    assert abs(float(ms.forward(o, tt)) - float(g["mse"])) < 1e-5 * float(g["mse"])
    assert rel_err(ms.backward(o, tt).data, g["mse_grad"]) < 1e-6


def test_optimizers_golden(okl):
    g = golden("optim_sched")
    ad, sg = okl.AdamOptimizer(lr=0.01), okl.SGDOptimizer(lr=0.01, momentum=0.9)
    wa, ws = g["w0"].copy(), g["w0"].copy()
    for gr in g["g"]:
        ad.update(wa, gr, "k")
        sg.update(ws, gr, "k")
    assert ad.t == 3
    assert rel_err(wa, g["adam_w"]) < 1e-5 and rel_err(ws, g["sgd_w"]) < 1e-6


def test_tensor_ops(okl):
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal((5, 7)), rng.standard_normal((7, 4))
    ta, tb = okl.Tensor(a), okl.Tensor(b)
    assert rel_err(ta.dot(tb).data, a @ b) < TOL32
    assert rel_err((ta + okl.Tensor(a * 2)).data, 3 * a) < 1e-6
    assert rel_err((ta + okl.Tensor(b[:1, :1] * 0 + 1.5)).data, a + 1.5) < 1e-6
    assert rel_err((ta + okl.Tensor(a[:1])).data, a + a[:1]) < 1e-6
    assert rel_err((ta * okl.Tensor(a)).data, a * a) < 1e-6
    assert rel_err((ta * 2.5).data, a * 2.5) < 1e-6
    assert rel_err(ta.transpose().data, a.T) < 1e-7
    assert rel_err(ta.reshape(7, 5).data, a.reshape(7, 5)) < 1e-7 and ta.reshape((35,)).shape == (35,)
    out = ta.dot(tb)
    out.backward()
    assert rel_err(ta.grad, np.ones((5, 4)) @ b.T) < TOL32 and rel_err(tb.grad, a.T @ np.ones((5, 4))) < TOL32
    c = ta.copy()
    assert c is not ta and np.array_equal(c.data, ta.data)
    assert repr(okl.Tensor([1, 2])) == "Tensor([1 2])"
    with pytest.raises(RuntimeError):
        okl.Tensor([1.0]).backward()
    # write-through of assignments and augmented assignments; in-place edits of a snapshot are re-uploaded
    t = okl.Tensor(a)
    t.data = a * 2
    t.data -= a
    assert rel_err(t.dot(tb).data, a @ b) < TOL32
    snap = t.data
    snap[0, 0] += 100.0
    assert abs(t.dot(tb).data[0, 0] - ((a @ b)[0, 0] + 100.0 * b[0, 0])) < 1e-3


def test_shape_errors_are_valueerrors(okl):
    with pytest.raises(ValueError):
        okl.DenseLayer(4, 3).forward(okl.Tensor(np.zeros((2, 5))))
    with pytest.raises(ValueError):
        okl.Conv2DLayer(3, 2, 3).forward(okl.Tensor(np.zeros((2, 4, 8, 8))))
    with pytest.raises(ValueError):
        okl.Conv2DLayer(3, 2, 5).forward(okl.Tensor(np.zeros((2, 3, 4, 4))))
    with pytest.raises(ValueError):
        okl.MaxPoolingLayer().forward(okl.Tensor(np.zeros((4, 4, 3))))
    d = okl.DenseLayer(4, 3)
    d.forward(okl.Tensor(np.zeros((2, 4))))
    with pytest.raises(ValueError):
        d.backward(okl.Tensor(np.zeros((2, 5))), 0.1)


@pytest.mark.parametrize("M,K,N", [(1, 1, 1), (3, 5, 2), (64, 64, 64), (65, 129, 67), (256, 5408, 128), (256, 128, 10),
                                   (300, 1000, 200), (1024, 1024, 1024)])
def test_dense_vs_oracle_sizes(okl, M, K, N):
    rng = np.random.default_rng(M * 7 + K * 3 + N)
    x, W, b = rng.standard_normal((M, K)), rng.standard_normal((K, N)) * 0.1, rng.standard_normal((1, N))
    g = rng.standard_normal((M, N))
    lay = okl.DenseLayer(K, N)
    lay.weights.data, lay.biases.data = W, b
    y = lay.forward(okl.Tensor(x))
    assert rel_err(y.data, O.dense_forward(x, W, b)) < TOL32
    dX = lay.backward(okl.Tensor(g), 0.01)
    dW, db, dXo = O.dense_backward(x, W, g)
    assert rel_err(dX.data, dXo) < TOL32
    assert rel_err(lay.weights.grad, dW) < TOL32 and rel_err(lay.biases.grad, db) < TOL32


CONV_SEEDED = [
    # nd, N, C, spatial, O, k, s, p
    (2, 32, 1, (28, 28), 16, 3, 1, 0),          # BASELINE config 1 at full size
    (2, 8, 3, (32, 32), 64, 3, 1, 1),           # CIFAR layer 1
    (2, 4, 64, (16, 16), 64, 3, 1, 1),
    (2, 2, 32, (9, 9), 48, 3, 2, 1),
    (2, 3, 5, (12, 10), 7, (5, 3), (2, 1), (2, 0)),
    (1, 4, 64, (128,), 64, 3, 1, 1),            # Conv1D sequence shape, shortened
    (1, 2, 3, (50,), 5, 7, 3, 3),
    (3, 1, 3, (6, 20, 20), 8, 3, 1, 1),         # Conv3D video block, shortened
    (3, 2, 2, (5, 9, 8), 4, (2, 3, 3), (1, 2, 2), (0, 1, 1)),
]


@pytest.mark.parametrize("nd,N,C,sp,Oc,k,s,p", CONV_SEEDED)
def test_conv_vs_oracle_seeded(okl, nd, N, C, sp, Oc, k, s, p):
    rng = np.random.default_rng(1234 + N + C + Oc)
    kt = O._tup(k, nd)
    x = rng.standard_normal((N, C) + sp).astype(np.float32)
    W = rng.standard_normal((Oc, C) + kt) * 0.1
    b = rng.standard_normal((Oc, 1)) * 0.1
    g = {"W": W.astype(np.float64 if nd == 2 else np.float32), "b": b, "stride": np.array(O._tup(s, nd)),
         "padding": np.array(O._tup(p, nd))}
    lay = make_conv(okl, g)
    x64 = x.astype(np.float64)
    W64 = g["W"].astype(np.float64)
    y = lay.forward(okl.Tensor(x))
    yo = O.conv_forward_fast(x64, W64, b, s, p, dtype=np.float64)
    assert rel_err(y.data, yo) < TOL32
    gr = rng.standard_normal(yo.shape).astype(np.float32)
    dX = lay.backward(okl.Tensor(gr), 0.05)
    dW, db, dXo = O.conv_backward_fast(x64, W64, gr.astype(np.float64), s, p)
    assert rel_err(dX.data, dXo) < TOL32
    Wt = lay.weights if hasattr(lay, "weights") else lay.filters
    assert rel_err(Wt.grad, dW) < TOL32 and rel_err(lay.biases.grad, db) < TOL32
    assert rel_err(Wt.data, W64 - 0.05 * dW) < TOL32


def test_empty_batch(okl):
    assert okl.DenseLayer(4, 2).forward(okl.Tensor(np.zeros((0, 4)))).data.shape == (0, 2)
    assert okl.Conv2DLayer(2, 3, 3).forward(okl.Tensor(np.zeros((0, 2, 5, 5)))).data.shape == (0, 3, 3, 3)


def test_layouts_channels_last_roundtrip(okl):
    """The channels-last (NXC) physical layout is invisible at the host boundary."""
    from okrolearn_b200 import _lib
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 6, 5, 7))
    t = okl.Tensor(x)
    t._dptr(_lib.NXC)
    assert t._layout == _lib.NXC
    t._host = None
    assert rel_err(t.data, x) < 1e-7
    # pooling and activations work in either layout and agree
    a = okl.Tensor(x)
    a._dptr(_lib.NXC)
    y1 = okl.MaxPoolingLayer(2, 2).forward(a).data
    assert rel_err(y1, O.maxpool_forward(x, 2, 2)) < 1e-7
    # conv through the strided FFMA kernel with channels-last input and output
    lay = okl.Conv2DLayer(6, 4, 3, 1, 1)
    lay._out_layout = _lib.NXC
    W, b = lay.filters.data.copy(), lay.biases.data.copy()
    c = okl.Tensor(x)
    c._dptr(_lib.NXC)
    y = lay.forward(c)
    assert y._layout == _lib.NXC
    assert rel_err(y.data, O.conv_forward_fast(x, W, b, 1, 1)) < TOL32
    gr = rng.standard_normal(y.shape)
    dX = lay.backward(okl.Tensor(gr), 0.0)
    dW, db, dXo = O.conv_backward_fast(x, W, gr, 1, 1)
    assert rel_err(dX.data, dXo) < TOL32 and rel_err(lay.filters.grad, dW) < TOL32


def test_unfold_fold_layers(okl):
    """Unfold / Fold (okl.py:525-636) on the device against the reference's own vectors and the oracle on larger random shapes."""
    g = golden("unfold_fold")
    g2 = golden("unfold_fold2")
    x = g["x"]
    assert rel_err(okl.Unfold(2, 2).forward(okl.Tensor(x)).data, g["unfold_k2s2"]) < TOL32
    assert rel_err(okl.Unfold(3, 1).forward(okl.Tensor(x)).data, g["unfold_k3s1"]) < TOL32
    up = okl.Unfold(3, 2, 1)
    yp = up.forward(okl.Tensor(x))
    assert yp.data.dtype == np.float64 and rel_err(yp.data, g2["unfold_k3s2p1"]) < TOL32
    with pytest.raises(ValueError):
        up.backward(okl.Tensor(np.zeros(yp.shape)))                 # okl.py:633 raises for padding > 0
    u3 = okl.Unfold(3, 1, 0)
    u3.forward(okl.Tensor(x))
    assert rel_err(u3.backward(okl.Tensor(g2["g_k3s1"])).data, g2["dx_k3s1"]) < TOL32
    u2 = okl.Unfold(2, 2, 0)
    u2.forward(okl.Tensor(x))
    assert rel_err(u2.backward(okl.Tensor(g2["g_k2s2"])).data, g2["dx_k2s2"]) < TOL32
    fo = okl.Fold((6, 6), 2, 2, 0)
    assert rel_err(fo.forward(okl.Tensor(g2["fold_in"])).data, g2["fold_out"]) < TOL32
    assert rel_err(fo.backward(okl.Tensor(g2["fold_g"])).data, g2["fold_bwd"]) < TOL32
    with pytest.raises(ValueError):
        okl.Fold((6, 6), 4, 2, 0).forward(okl.Tensor(g2["fold_in"]))
    # larger random shapes vs the oracle; W.reshape(O, -1) @ unfold(x) == conv(x) ties the K ordering to the conv layers
    rng = np.random.default_rng(12)
    xb = rng.standard_normal((3, 5, 19, 23))
    for k, s, p in ((3, 1, 0), (4, 2, 0), (3, 2, 2), (5, 3, 1)):
        assert rel_err(okl.Unfold(k, s, p).forward(okl.Tensor(xb)).data, O.unfold_forward_loops(xb, k, s, p)) < TOL32
    ub = okl.Unfold(4, 2, 0)
    cols = ub.forward(okl.Tensor(xb))
    gb = rng.standard_normal(cols.shape)
    assert rel_err(ub.backward(okl.Tensor(gb)).data, O.unfold_backward(gb, xb.shape, 4, 2, 0)) < TOL32
    Wc = rng.standard_normal((7, 5, 3, 3))
    conv = okl.Conv2DLayer(5, 7, 3)
    conv.filters.data = Wc.copy()
    cols3 = okl.Unfold(3, 1, 0).forward(okl.Tensor(xb)).data
    yc = conv.forward(okl.Tensor(xb)).data
    assert rel_err((Wc.reshape(7, -1) @ cols3).reshape(yc.shape), yc) < 2e-5


def test_other_activation_layers(okl):
    """the remaining elementwise / row-wise activation layers on the device vs the reference's vectors."""
    g = golden("activations2")
    x, gg = g["x"], g["g"]
    for nm, lay in (("elu", okl.ELUActivationLayer(0.7)), ("selu", okl.SELUActivationLayer()), ("leaky", okl.LeakyReLUActivationLayer(0.05)),
                    ("swish", okl.SwishActivationLayer()), ("gelu", okl.GELUActivationLayer()), ("softsign", okl.SoftsignActivationLayer()),
                    ("hardtanh", okl.HardTanhActivationLayer())):
        xt = okl.Tensor(x.copy())
        y = lay.forward(xt)
        assert rel_err(y.data, g[nm + "_y"]) < 2e-6, nm
        dx = lay.backward(okl.Tensor(gg.copy()), 0.1)
        assert rel_err(dx.data, g[nm + "_dx"]) < 2e-6, nm
        if nm not in ("softsign",):
            assert rel_err(np.asarray(xt.grad), g[nm + "_dx"]) < 2e-6, nm      # inputs.grad side effect (okl.py:81, 151, 187, 219, 304)
        else:
            assert xt.grad is None
    pl = okl.PReLUActivationLayer(0.05)
    xt = okl.Tensor(x.copy())
    assert rel_err(pl.forward(xt).data, g["prelu_y"]) < 2e-6
    assert rel_err(pl.backward(okl.Tensor(gg.copy()), 0.1).data, g["prelu_dx"]) < 2e-6
    assert abs(pl.d_alpha - float(g["prelu_d_alpha"])) < 1e-4 and abs(pl.alpha - float(g["prelu_alpha_after"])) < 1e-5
    assert pl.get_params() == {"alpha": pl.alpha}
    sm = okl.SoftminActivationLayer()
    assert rel_err(sm.forward(okl.Tensor(x.copy())).data, g["softmin_y"]) < 2e-6
    assert rel_err(sm.backward(okl.Tensor(gg.copy())).data, g["softmin_dx"]) < 2e-5
    ls = okl.LogSoftmaxActivationLayer()
    yl = ls.forward(okl.Tensor(x.copy()))
    assert yl.data.dtype == np.float64 and rel_err(yl.data, g["logsoftmax_y"]) < 2e-6
    assert rel_err(ls.backward(okl.Tensor(gg.copy())).data, g["logsoftmax_dx"]) < 2e-5
    li = okl.LinearActivationLayer()
    xt = okl.Tensor(x.copy())
    assert li.forward(xt) is xt
    assert rel_err(li.backward(okl.Tensor(gg.copy()), 0.1).data, gg) < 1e-7 and rel_err(np.asarray(xt.grad), gg) < 1e-7
    # a Dense -> GELU -> Dense -> LogSoftmax stack trains through NeuralNetwork.forward / backward with these layers
    rng = np.random.default_rng(5)
    net = okl.NeuralNetwork()
    d1, d2 = okl.DenseLayer(7, 9), okl.DenseLayer(9, 4)
    W1, W2 = d1.weights.data.copy(), d2.weights.data.copy()
    for l in (d1, okl.GELUActivationLayer(), d2, okl.LogSoftmaxActivationLayer()):
        net.add(l)
    out = net.forward(okl.Tensor(x))
    h = O.gelu(x @ W1)
    assert rel_err(out.data, O.logsoftmax_forward(h @ W2)) < 2e-5
    net.backward(okl.Tensor(gg[:, :4].copy()), 0.01)
    dz2 = O.logsoftmax_backward(O.logsoftmax_forward(h @ W2), gg[:, :4])
    assert rel_err(d2.weights.data, W2 - 0.01 * (h.T @ dz2)) < 2e-5
    dz1 = (dz2 @ W2.T) * O.gelu_d(x @ W1)
    assert rel_err(d1.weights.data, W1 - 0.01 * (x.T @ dz1)) < 2e-5


def test_batchnorm_layer(okl):
    """BatchNormLayer on the device vs the reference's vectors: two training steps (outputs, input gradients, the
    accumulated-gradient update of gamma / beta, running statistics), eval mode, and the params round trip."""
    g = golden("batchnorm")
    bn = okl.BatchNormLayer(5)
    for step in (0, 1):
        y = bn.forward(okl.Tensor(g[f"x{step}"]))
        assert y.data.dtype == np.float64 and rel_err(y.data, g[f"y{step}"]) < 2e-6
        dx = bn.backward(okl.Tensor(g[f"g{step}"]), 0.05)
        assert rel_err(dx.data, g[f"dx{step}"]) < 2e-5
        assert rel_err(bn.gamma.data, g[f"gamma{step}"]) < 2e-6 and rel_err(bn.beta.data, g[f"beta{step}"]) < 2e-6
        assert rel_err(bn.running_mean, g[f"rm{step}"]) < 2e-6 and rel_err(bn.running_var, g[f"rv{step}"]) < 2e-6
    bn.training = False
    assert rel_err(bn.forward(okl.Tensor(g["x0"])).data, g["y_eval"]) < 2e-6
    p = bn.get_params()
    bn2 = okl.BatchNormLayer(5)
    bn2.set_params(p)
    bn2.training = False
    assert rel_err(bn2.forward(okl.Tensor(g["x0"])).data, g["y_eval"]) < 2e-6
    # inside a network, between two Dense layers, on larger random data vs the oracle
    rng = np.random.default_rng(31)
    X, G = rng.standard_normal((64, 40)), rng.standard_normal((64, 24))
    net = okl.NeuralNetwork()
    d1, b1, d2 = okl.DenseLayer(40, 33), okl.BatchNormLayer(33), okl.DenseLayer(33, 24)
    W1, W2 = d1.weights.data.copy(), d2.weights.data.copy()
    for l in (d1, b1, okl.ReLUActivationLayer(), d2):
        net.add(l)
    on = O.OracleNetwork([O.OracleDense(W1, np.zeros((1, 33))), O.OracleBatchNorm(33), O.OracleReLU(), O.OracleDense(W2, np.zeros((1, 24)))])
    for _ in range(2):
        out = net.forward(okl.Tensor(X))
        oo = on.forward(X)
        assert rel_err(out.data, oo) < 5e-5
        net.backward(okl.Tensor(G), 0.01)
        on.backward(G, 0.01)
    assert rel_err(d1.weights.data, on.layers[0].W) < 5e-5 and rel_err(b1.gamma.data, on.layers[1].gamma) < 5e-5
    assert rel_err(b1.running_var, on.layers[1].running_var) < 5e-5


def test_regularization_wrappers_and_dropout(okl):
    """L1 / L2 / L3RegularizationLayer around BatchNormLayer (two steps: the penalty is added to the accumulated gradients and applied
    once more, okl.py:1070-1078) and DropoutLayer with the reference's seeded host mask, vs vectors from the running reference."""
    g = golden("reg_dropout")
    for nm, cls in (("l1", okl.L1RegularizationLayer), ("l2", okl.L2RegularizationLayer), ("l3", okl.L3RegularizationLayer)):
        lay = cls(okl.BatchNormLayer(5), 0.01)
        for step in (0, 1):
            lay.forward(okl.Tensor(g[f"x{step}"]))
            dx = lay.backward(okl.Tensor(g[f"g{step}"]), 0.05)
            assert rel_err(dx.data, g[f"{nm}_dx{step}"]) < 3e-5, nm
        assert rel_err(lay.layer.gamma.data, g[f"{nm}_gamma"]) < 3e-6 and rel_err(lay.layer.beta.data, g[f"{nm}_beta"]) < 3e-6, nm
        assert rel_err(np.asarray(lay.layer.gamma.grad).reshape(-1), g[f"{nm}_ggrad"].reshape(-1)) < 3e-6, nm
        assert rel_err(np.asarray(lay.layer.beta.grad).reshape(-1), g[f"{nm}_bgrad"].reshape(-1)) < 3e-6, nm
    np.random.seed(123)
    dl = okl.DropoutLayer(0.3)
    y = dl.forward(okl.Tensor(g["x0"]))
    assert np.array_equal(dl.mask, g["drop_mask"])            # same host RNG call -> the same units are dropped
    assert rel_err(y.data, g["drop_y"]) < 1e-6
    assert rel_err(dl.backward(okl.Tensor(g["g0"]), 0.05).data, g["drop_dx"]) < 1e-6
    xt = okl.Tensor(g["x0"])
    assert dl.forward(xt, training=False) is xt


def test_conv2d_transposed_layer(okl):
    """Conv2DLayer(transposed=True) = dgrad / fprop / wgrad of the mirrored convolution; forward vs the reference's
    outputs, backward vs the torch-checked sums, the in-layer accumulated update, and the padding > 0 error behaviour."""
    g = golden("conv2d_transposed")
    for ci in range(3):
        Ci, Co, kh, kw, sh, sw, ph, pw = [int(v) for v in g[f"cfg{ci}"]]
        lay = okl.Conv2DLayer(Ci, Co, (kh, kw), (sh, sw), (ph, pw), transposed=True)
        assert lay.filters.shape == (Ci, Co, kh, kw)
        lay.filters.data, lay.biases.data = g[f"F{ci}"].copy(), g[f"b{ci}"].copy()
        y = lay.forward(okl.Tensor(g[f"x{ci}"]))
        assert y.data.dtype == np.float64 and rel_err(y.data, g[f"y{ci}"]) < TOL32
        dX = lay.backward(okl.Tensor(g[f"g{ci}"]), 0.1)
        assert rel_err(dX.data, g[f"dX{ci}"]) < TOL32
        assert rel_err(lay.filters.grad, g[f"dF{ci}"]) < TOL32 and rel_err(np.asarray(lay.biases.grad).reshape(-1), g[f"db{ci}"].reshape(-1)) < TOL32
        assert rel_err(lay.filters.data, g[f"F{ci}"] - 0.1 * g[f"dF{ci}"]) < TOL32
        assert rel_err(lay.biases.data, g[f"b{ci}"] - 0.1 * g[f"db{ci}"]) < TOL32
    with pytest.raises(ValueError):
        okl.Conv2DLayer(2, 3, 3, 1, 1, transposed=True).forward(okl.Tensor(np.zeros((1, 2, 4, 4))))
    # encoder / decoder stack through NeuralNetwork: Conv2D(s2) -> ReLU -> transposed Conv2D(s2) restores the extent
    rng = np.random.default_rng(8)
    X = rng.standard_normal((4, 3, 9, 9))
    X.reshape(-1)[0] = abs(X.reshape(-1)[0]) + 0.1
    enc, dec = okl.Conv2DLayer(3, 6, 3, 2, 0), okl.Conv2DLayer(6, 3, 3, 2, 0, transposed=True)
    We, Wd = enc.filters.data.copy(), dec.filters.data.copy()
    net = okl.NeuralNetwork()
    for l in (enc, okl.ReLUActivationLayer(), dec):
        net.add(l)
    out = net.forward(okl.Tensor(X))
    h = O.relu_forward(O.conv_forward_fast(X, We, np.zeros((6, 1)), 2, 0, dtype=np.float64))
    yo = O.conv2d_transposed_forward(h, Wd, np.zeros((3, 1)), 2, 0)
    assert out.shape == X.shape and rel_err(out.data, yo) < 2e-5
    G = rng.standard_normal(X.shape)
    net.backward(okl.Tensor(G), 0.05)
    dF, db, dh = O.conv2d_transposed_backward(h, Wd, G, 2, 0)
    assert rel_err(dec.filters.data, Wd - 0.05 * dF) < 2e-5
    dWe, dbe, _ = O.conv_backward_fast(X, We, dh * (h > 0), 2, 0)
    assert rel_err(enc.filters.data, We - 0.05 * dWe) < 2e-5


def test_unpooling_and_padding_layers(okl):
    """Max / AverageUnpoolingLayer (overlapping stamps: last wins; gaps stay 0) and Zero / ReflectionPaddingLayer vs the
    running reference's vectors."""
    g = golden("unpool_pad")
    x = g["x"]
    for nm, cls in (("max", okl.MaxUnpoolingLayer), ("avg", okl.AverageUnpoolingLayer)):
        for tag, ps, st in (("a", 2, 2), ("b", 3, 2), ("c", 2, 3)):
            lay = cls(ps, st)
            xt = okl.Tensor(x.copy())
            assert rel_err(lay.forward(xt).data, g[f"{nm}_{tag}_y"]) < TOL32, (nm, tag)
            dx = lay.backward(okl.Tensor(g[f"{nm}_{tag}_g"]))
            assert rel_err(dx.data, g[f"{nm}_{tag}_dx"]) < TOL32, (nm, tag)
            assert rel_err(np.asarray(xt.grad), g[f"{nm}_{tag}_dx"]) < TOL32
    for nm, cls in (("zero", okl.ZeroPaddingLayer), ("reflect", okl.ReflectionPaddingLayer)):
        for tag, pd in (("a", 1), ("b", (2, 1)), ("c", (1, 2, 3, 1))):
            lay = cls(pd)
            y = lay.forward(okl.Tensor(x.copy()))
            assert y.shape == g[f"{nm}_{tag}_y"].shape and rel_err(y.data, g[f"{nm}_{tag}_y"]) < 1e-7, (nm, tag)
            dx = lay.backward(okl.Tensor(g[f"{nm}_{tag}_g"]))
            assert dx.shape == g[f"{nm}_{tag}_dx"].shape and rel_err(dx.data, g[f"{nm}_{tag}_dx"]) < 1e-7, (nm, tag)
    with pytest.raises(ValueError):
        okl.ZeroPaddingLayer((1, 2, 3))
    with pytest.raises(NotImplementedError):
        okl.PaddingLayer(1).forward(okl.Tensor(x))


def test_rnn_layer(okl):
    """RNNLayer (one Elman step on the Dense kernels) unrolled over three steps vs the running reference: hidden
    states, both returned gradients, the accumulated (never applied) weight gradients, and the inputs.grad / prev_hidden.grad side
    effects."""
    g = golden("rnn")
    lay = okl.RNNLayer(6, 5)
    lay.Wh.data, lay.Wx.data, lay.b.data = g["Wh"].copy(), g["Wx"].copy(), g["b"].copy()
    hp = None
    for t in range(3):
        xt = okl.Tensor(g[f"x{t}"])
        ht = None if hp is None else okl.Tensor(hp)
        h = lay.forward(xt, ht)
        assert h.data.dtype == np.float64 and rel_err(h.data, g[f"h{t}"]) < 2e-6
        dx, dp = lay.backward(okl.Tensor(g[f"g{t}"]), 0.1)
        assert rel_err(dx.data, g[f"dx{t}"]) < 2e-5 and rel_err(dp.data, g[f"dprev{t}"]) < 2e-5
        assert rel_err(np.asarray(xt.grad), g[f"dx{t}"]) < 2e-5
        hp = g[f"h{t}"]
    assert rel_err(lay.Wh.grad, g["gWh"]) < 2e-5 and rel_err(lay.Wx.grad, g["gWx"]) < 2e-5 and rel_err(lay.b.grad, g["gb"]) < 2e-5
    assert rel_err(lay.Wh.data, g["Wh"]) < 1e-7 and rel_err(lay.Wx.data, g["Wx"]) < 1e-7       # backward never updates (okl.py:1669-1674)
    assert set(lay.get_params()) == {"Wh", "Wx", "b"}


def test_lstm_layer(okl):
    """LSTMLayer (gates on the Dense kernels + pointwise cell kernels) unrolled over three steps vs the running
    reference: hidden / cell states, the two returned gradients, the eight accumulated (never applied) parameter gradients."""
    g = golden("lstm")
    lay = okl.LSTMLayer(6, 5)
    for nm in ("Wf", "Wi", "Wc", "Wo", "bf", "bi", "bc", "bo"):
        getattr(lay, nm).data = g[nm].copy()
    hp = cp = None
    for t in range(3):
        h, c = lay.forward(okl.Tensor(g[f"x{t}"]), None if hp is None else okl.Tensor(hp), None if cp is None else okl.Tensor(cp))
        assert rel_err(h.data, g[f"h{t}"]) < 3e-6 and rel_err(c.data, g[f"c{t}"]) < 3e-6
        dh, dc = lay.backward(okl.Tensor(g[f"gh{t}"]), okl.Tensor(g[f"gc{t}"]), 0.1)
        assert rel_err(dh.data, g[f"dh{t}"]) < 3e-5 and rel_err(dc.data, g[f"dc{t}"]) < 3e-5
        hp, cp = g[f"h{t}"], g[f"c{t}"]
    for nm in ("Wf", "Wi", "Wc", "Wo", "bf", "bi", "bc", "bo"):
        assert rel_err(np.asarray(getattr(lay, nm).grad).reshape(g["g" + nm].shape), g["g" + nm]) < 3e-5, nm
        assert rel_err(getattr(lay, nm).data, g[nm]) < 1e-7, nm                 # backward never updates the weights
    assert set(lay.get_params()) == {"Wf", "Wi", "Wc", "Wo", "bf", "bi", "bc", "bo"}


def test_gru_layer(okl):
    """GRULayer unrolled over three steps vs the running reference (hidden states, dL/dx, dL/dh_prev, the six accumulated
    parameter gradients; weights untouched by backward)."""
    g = golden("gru")
    lay = okl.GRULayer(6, 5)
    for nm in ("Wz", "Wr", "Wh", "bz", "br", "bh"):
        getattr(lay, nm).data = g[nm].copy()
    hp = None
    for t in range(3):
        h = lay.forward(okl.Tensor(g[f"x{t}"]), None if hp is None else okl.Tensor(hp))
        assert rel_err(h.data, g[f"h{t}"]) < 3e-6
        dx, dp = lay.backward(okl.Tensor(g[f"gh{t}"]), 0.1)
        assert rel_err(dx.data, g[f"dx{t}"]) < 3e-5 and rel_err(dp.data, g[f"dprev{t}"]) < 3e-5
        hp = g[f"h{t}"]
    for nm in ("Wz", "Wr", "Wh", "bz", "br", "bh"):
        assert rel_err(np.asarray(getattr(lay, nm).grad).reshape(g["g" + nm].shape), g["g" + nm]) < 3e-5, nm
        assert rel_err(getattr(lay, nm).data, g[nm]) < 1e-7, nm


def test_conv_random_shapes_fp32(okl):
    """Randomised parity of Conv1D / Conv2D / Conv3D (forward, dX, dW, db, in-layer update) in FP32 mode against the oracle: strides
    1-3, paddings 0-2, rectangular kernels, odd channel counts - beyond the enumerated golden cases."""
    rng = np.random.default_rng(4242)
    for case in range(18):
        nd = int(rng.integers(1, 4))
        N, C, Oc = int(rng.integers(1, 5)), int(rng.integers(1, 7)), int(rng.integers(1, 9))
        k = tuple(int(v) for v in rng.integers(1, 4, nd))
        s = tuple(int(v) for v in rng.integers(1, 4, nd))
        p = tuple(int(v) for v in rng.integers(0, 3, nd))
        sp = tuple(int(kk + rng.integers(0, 9)) for kk in k)
        x = rng.standard_normal((N, C) + sp)
        W = rng.standard_normal((Oc, C) + k) * 0.3
        b = rng.standard_normal((Oc, 1)) * 0.2
        if nd == 1:
            lay = okl.Conv1DLayer(C, Oc, k[0], s[0], p[0])
            lay.weights.data = W.copy()
        elif nd == 2:
            lay = okl.Conv2DLayer(C, Oc, k, s, p)
            lay.filters.data = W.copy()
        else:
            lay = okl.Conv3DLayer(C, Oc, k, s, p)
            lay.filters.data = W.copy()
        lay.biases.data = b.copy()
        Wt = lay.weights if nd == 1 else lay.filters
        tag = (case, nd, N, C, Oc, k, s, p, sp)
        y = lay.forward(okl.Tensor(x))
        yo = O.conv_forward_fast(x, W, b, s, p, dtype=np.float64)
        assert y.shape == yo.shape and rel_err(y.data, yo) < 2e-5, tag
        g = rng.standard_normal(yo.shape)
        dX = lay.backward(okl.Tensor(g), 0.05)
        dW, db, dXo = O.conv_backward_fast(x, W, g, s, p)
        assert rel_err(dX.data, dXo) < 2e-5, tag
        assert rel_err(Wt.grad, dW) < 2e-5 and rel_err(np.asarray(lay.biases.grad).reshape(-1), np.asarray(db).reshape(-1)) < 2e-5, tag
        assert rel_err(Wt.data, W - 0.05 * dW) < 2e-5, tag


def test_dense_random_shapes_and_empty_batches(okl):
    """Randomised DenseLayer parity in FP32 mode (tiny-N head kernel, FFMA tiles, odd sizes) and the empty-batch edge cases of Dense and
    Conv2D (zero rows in, zero rows out, zero gradients, parameters unchanged)."""
    rng = np.random.default_rng(99)
    for case in range(14):
        M, K, N = int(rng.integers(1, 70)), int(rng.integers(1, 90)), int(rng.integers(1, 40))
        x, W, b = rng.standard_normal((M, K)), rng.standard_normal((K, N)) * 0.3, rng.standard_normal((1, N)) * 0.2
        lay = okl.DenseLayer(K, N)
        lay.weights.data, lay.biases.data = W.copy(), b.copy()
        y = lay.forward(okl.Tensor(x))
        assert rel_err(y.data, x @ W + b) < 2e-5, (case, M, K, N)
        g = rng.standard_normal((M, N))
        dX = lay.backward(okl.Tensor(g), 0.05)
        assert rel_err(dX.data, g @ W.T) < 2e-5, (case, M, K, N)
        assert rel_err(lay.weights.data, W - 0.05 * (x.T @ g)) < 2e-5 and rel_err(lay.biases.data, b - 0.05 * g.sum(0, keepdims=True)) < 2e-5
    d = okl.DenseLayer(7, 5)
    W0 = d.weights.data.copy()
    y0 = d.forward(okl.Tensor(np.zeros((0, 7))))
    assert y0.shape == (0, 5) and y0.data.shape == (0, 5)
    dx0 = d.backward(okl.Tensor(np.zeros((0, 5))), 0.1)
    assert dx0.shape == (0, 7) and np.array_equal(d.weights.data.astype(np.float32), W0.astype(np.float32))
    c = okl.Conv2DLayer(2, 3, 3, 1, 1)
    F0 = c.filters.data.copy()
    yc = c.forward(okl.Tensor(np.zeros((0, 2, 6, 6))))
    assert yc.shape == (0, 3, 6, 6)
    dxc = c.backward(okl.Tensor(np.zeros((0, 3, 6, 6))), 0.1)
    assert dxc.shape == (0, 2, 6, 6) and np.array_equal(c.filters.data.astype(np.float32), F0.astype(np.float32))
