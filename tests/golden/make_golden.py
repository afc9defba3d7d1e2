"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE in the build container.

Run once here (``python tests/golden/make_golden.py``); the GPU box has no /root/reference, so the
``.npz`` files it writes are committed.  For every case it
  1. runs the shipped reference code (forward everywhere; backward where the shipped code runs),
  2. asserts the oracle restatement (oracle/okl_oracle.py) reproduces it,
  3. for conv/pool backward (shipped code raises ValueError) takes the R1-R5 oracle
     and asserts it against torch.autograd on CPU in fp64,
  4. stores inputs + expected outputs.  Key ``src_<name>`` says "reference" or "oracle+torch".
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference():
    plt = types.ModuleType("matplotlib.pyplot")
    def _stub_attr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None

    plt.__getattr__ = _stub_attr
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    fl = types.ModuleType("flask")
    fl.Flask = type("Flask", (), {"__init__": lambda s, *a, **k: None})
    fl.request = None
    fl.jsonify = lambda *a, **k: None
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "flask": fl})
    sys.path.insert(0, "/root/reference/src")
    import okrolearn.okrolearn as ref  # must come before okrolearn.tensor (circular import)
    import okrolearn.optimizers as ropt
    return ref, ropt


import torch  # noqa: E402  (before the stubs go in: torch inspects sys.modules)
import torch.nn.functional as F  # noqa: E402

ref, ropt = load_reference()
from oracle import okl_oracle as O  # noqa: E402

T = ref.Tensor
out = {}


def close(a, b, tol=1e-12, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30) if a.size else 0.0
    assert err <= tol, (what, err)
    return err


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(f"wrote {name}.npz  ({', '.join(arrs)})")


rng = np.random.default_rng(1234)

# ----------------------------------------------------------------------------- Dense
x = rng.standard_normal((8, 12))
W = rng.standard_normal((12, 7)) * 0.1
b = rng.standard_normal((1, 7)) * 0.1
g1 = rng.standard_normal((8, 7))
g2 = rng.standard_normal((8, 7))
lay = ref.DenseLayer(12, 7)
lay.weights.data = W.copy()
lay.biases.data = b.copy()
y = lay.forward(T(x)).data
close(O.dense_forward(x, W, b), y, what="dense fwd")
dx1 = lay.backward(T(g1), 0.05).data
W1, b1 = lay.weights.data.copy(), lay.biases.data.copy()
lay.forward(T(x))
dx2 = lay.backward(T(g2), 0.05).data
W2, b2 = lay.weights.data.copy(), lay.biases.data.copy()
od = O.OracleDense(W, b)
od.forward(x)
close(od.backward(g1, 0.05), dx1, what="dense dX1")
close(od.W, W1, what="dense W1")
od.forward(x)
close(od.backward(g2, 0.05), dx2, what="dense dX2")
close(od.W, W2, what="dense W2")
close(od.b, b2, what="dense b2")
save("dense", x=x, W=W, b=b, g1=g1, g2=g2, lr=np.float64(0.05), y=y, dx1=dx1, W1=W1, b1=b1, dx2=dx2, W2=W2, b2=b2,
     src="reference")

# ----------------------------------------------------------------------------- Conv forward (reference) + backward (oracle+torch)
conv_cases = {
    # name: (nd, N, C, spatial, O, k, stride, pad)
    "conv2d_cfg1_small": (2, 4, 1, (10, 10), 6, 3, 1, 0),
    "conv2d_s2p1": (2, 3, 3, (9, 11), 5, 3, 2, 1),
    "conv2d_rect": (2, 2, 4, (8, 7), 3, (3, 2), (1, 2), (1, 0)),
    "conv2d_c16": (2, 2, 16, (8, 8), 16, 3, 1, 1),
    "conv1d_p0": (1, 3, 2, (17,), 4, 3, 1, 0),
    "conv1d_s2p2": (1, 2, 5, (20,), 3, 5, 2, 2),
    "conv3d_p0": (3, 2, 2, (5, 6, 7), 3, 3, 1, 0),
    "conv3d_p1": (3, 1, 3, (4, 6, 6), 4, 3, 1, 1),
    "conv3d_s2": (3, 2, 2, (7, 7, 6), 3, (3, 3, 2), (2, 2, 1), (1, 1, 0)),
}
torch_conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}
for name, (nd, N, C, sp, Oc, k, s, p) in conv_cases.items():
    kt = O._tup(k, nd)
    wdt = np.float64 if nd == 2 else np.float32
    x = rng.standard_normal((N, C) + sp).astype(np.float32)
    Wt = (rng.standard_normal((Oc, C) + kt) * 0.1).astype(wdt)
    bt = (rng.standard_normal((Oc, 1)) * 0.1).astype(wdt)
    if nd == 1:
        lay = ref.Conv1DLayer(C, Oc, k, s, p)
        lay.weights.data = Wt.copy()
    elif nd == 2:
        lay = ref.Conv2DLayer(C, Oc, k, s, p)
        lay.filters.data = Wt.copy()
    else:
        lay = ref.Conv3DLayer(C, Oc, k, s, p)
        lay.filters.data = Wt.copy()
    lay.biases.data = bt.copy()
    y = lay.forward(T(x)).data
    yo = O.conv_forward_loops(x, Wt, bt, s, p)
    assert yo.dtype == y.dtype, (name, yo.dtype, y.dtype)
    close(yo, y, 1e-6 if nd != 2 else 1e-13, what=name + " fwd")
    close(O.conv_forward_fast(x, Wt, bt, s, p), y, 1e-5 if nd != 2 else 1e-12, what=name + " fwd fast")
    # the shipped backward must raise (documents why the gradient oracle is the repaired restatement)
    g = rng.standard_normal(y.shape).astype(np.float32)
    try:
        lay.backward(T(g), 0.01)
        raised = False
    except ValueError:
        raised = True
    assert raised, name + ": shipped conv backward unexpectedly ran"
    # R1-R3 oracle in fp64, pinned against torch.autograd fp64
    x64, W64, b64, g64 = x.astype(np.float64), Wt.astype(np.float64), bt.astype(np.float64), g.astype(np.float64)
    # run the loops in float64 regardless of nd by going through the 2-D dtype rule manually
    dW, db, dX = O.conv_backward_loops(x64, W64, g64, s, p)
    tx = torch.tensor(x64, requires_grad=True)
    tw = torch.tensor(W64, requires_grad=True)
    tb = torch.tensor(b64.reshape(-1), requires_grad=True)
    ty = torch_conv[nd](tx, tw, tb, stride=s, padding=p)
    close(ty.detach().numpy(), y, 1e-6 if nd != 2 else 1e-13, what=name + " torch fwd")
    ty.backward(torch.tensor(g64))
    close(dW, tw.grad.numpy(), 1e-12, what=name + " dW")
    close(db.reshape(-1), tb.grad.numpy(), 1e-12, what=name + " db")
    close(dX, tx.grad.numpy(), 1e-12, what=name + " dX")
    dWf, dbf, dXf = O.conv_backward_fast(x64, W64, g64, s, p)
    close(dWf, dW, 1e-12, what=name + " dW fast")
    close(dXf, dX, 1e-12, what=name + " dX fast")
    save(name, x=x, W=Wt, b=bt, g=g, y=y, dW=dW, db=db, dX=dX, stride=np.array(O._tup(s, nd)),
         padding=np.array(O._tup(p, nd)), src_y="reference", src_grads="oracle+torch")

# ----------------------------------------------------------------------------- the reference's own conv KAT inputs
# tests/conv2d.py:8-48: filters = two 3x3 ramps (1..9 and 9..1) repeated over 3 channels, input = arange(1..48)
lay = ref.Conv2DLayer(3, 2, (3, 3), stride=1, padding=0)
ramp = np.arange(1, 10).reshape(3, 3)
Wk = np.stack([np.stack([ramp] * 3), np.stack([ramp[::-1, ::-1]] * 3)])
bk = np.array([[1], [2]])
xk = np.arange(1, 49).reshape(1, 3, 4, 4)
lay.filters.data = Wk
lay.biases.data = bk
yk = lay.forward(T(xk)).data
close(O.conv_forward_loops(xk, Wk, bk), yk, what="kat conv2d")
close(F.conv2d(torch.tensor(xk, dtype=torch.float64), torch.tensor(Wk, dtype=torch.float64),
               torch.tensor(bk.reshape(-1), dtype=torch.float64)).numpy(), yk, what="kat conv2d torch")
wrong_expected = np.array([[[[1428, 1531], [1840, 1943]], [[1053, 948], [633, 528]]]])
assert not np.allclose(yk, wrong_expected)
# tests/conv1d.py:8-31: weights = the same two ramps as (2,3,3), input = arange(1..15) as (1,3,5)
l1 = ref.Conv1DLayer(3, 2, 3, stride=1, padding=0)
Wk1 = np.stack([ramp, ramp[::-1, ::-1]])
xk1 = np.arange(1, 16).reshape(1, 3, 5)
l1.weights.data = Wk1
l1.biases.data = bk
yk1 = l1.forward(T(xk1)).data
close(O.conv_forward_loops(xk1, Wk1, bk), yk1, what="kat conv1d")
wrong_expected1 = np.array([[[339, 369, 399], [255, 276, 297]]])
assert not np.allclose(yk1, wrong_expected1)
save("kat_conv_reference_tests", x2=xk, W2=Wk, b2=bk, y2=yk, test_file_expected2_WRONG=wrong_expected,
     x1=xk1, W1=Wk1, b1=bk, y1=yk1, test_file_expected1_WRONG=wrong_expected1,
     note="tests/conv2d.py:43-48 and tests/conv1d.py:30-31 expected arrays disagree with the reference code and with "
          "torch; y1/y2 are the reference code's outputs")

# ----------------------------------------------------------------------------- activations
xa = rng.standard_normal((6, 9))
xa.flat[0] = abs(xa.flat[0]) + 0.1
ga = rng.standard_normal((6, 9))
r = ref.ReLUActivationLayer()
yr = r.forward(T(xa)).data
close(O.relu_forward(xa), yr, what="relu fwd")
close(O.relu_forward_as_shipped(xa), yr, what="relu fwd shipped")
dr = r.backward(T(ga), 0.0).data
close(O.relu_backward(xa, ga), dr, what="relu bwd")
xneg = xa.copy()
xneg.flat[0] = -0.5
yneg = ref.ReLUActivationLayer().forward(T(xneg)).data
assert yneg.dtype == np.int64 and O.relu_forward_as_shipped(xneg).dtype == np.int64
s_ = ref.SigmoidActivationLayer()
ys = s_.forward(T(xa)).data
close(O.sigmoid_forward(xa), ys, what="sigmoid fwd")
ds = s_.backward(T(ga)).data
close(O.sigmoid_backward(ys, ga), ds, what="sigmoid bwd")
t_ = ref.TanhActivationLayer()
yt = t_.forward(T(xa)).data
close(O.tanh_forward(xa), yt, what="tanh fwd")
dt_ = t_.backward(T(ga), 0.0).data
close(O.tanh_backward(xa, ga), dt_, what="tanh bwd")
sm = ref.SoftmaxActivationLayer()
ysm = sm.forward(T(xa)).data
close(O.softmax_forward(xa), ysm, what="softmax fwd")
dsm = sm.backward(T(ga)).data
close(O.softmax_backward_jacobian(ysm, ga), dsm, what="softmax bwd jac")
close(O.softmax_backward(ysm, ga), dsm, 1e-12, what="softmax bwd closed")
kat_sm = ref.SoftmaxActivationLayer().forward(T(np.array([[1.0, 2.0, 3.0]]))).data      # tests/dot_test.py:28-32
close(kat_sm, np.array([[0.09003057, 0.24472847, 0.66524096]]), 1e-7, what="softmax KAT")
save("activations", x=xa, g=ga, relu_y=yr, relu_dx=dr, relu_neg_first_dtype=str(yneg.dtype), sigmoid_y=ys,
     sigmoid_dx=ds, tanh_y=yt, tanh_dx=dt_, softmax_y=ysm, softmax_dx=dsm, softmax_kat_x=np.array([[1.0, 2.0, 3.0]]),
     softmax_kat_y=kat_sm, src="reference")

# ----------------------------------------------------------------------------- pooling
xp_ = rng.standard_normal((3, 4, 9, 8))
for kind, cls, fo, bo in (("max", ref.MaxPoolingLayer, O.maxpool_forward, O.maxpool_backward),
                          ("avg", ref.AveragePoolingLayer, O.avgpool_forward, O.avgpool_backward)):
    for (ps, st) in ((2, 2), (3, 2), (2, 1), (3, 3)):
        lay = cls(ps, st)
        y = lay.forward(T(xp_)).data
        close(fo(xp_, ps, st), y, what=f"{kind}pool fwd")
        g = rng.standard_normal(y.shape)
        try:
            lay.backward(T(g))
            raised = False
        except ValueError:
            raised = True
        assert raised, "shipped pool backward unexpectedly ran on NCHW"
        dX = bo(xp_, g, ps, st)
        if st >= ps:   # non-overlapping windows: equals torch.autograd (no ties in continuous random data)
            tx = torch.tensor(xp_, requires_grad=True)
            ty = (F.max_pool2d if kind == "max" else F.avg_pool2d)(tx, ps, st)
            ty.backward(torch.tensor(g))
            close(dX, tx.grad.numpy(), 1e-12, what=f"{kind}pool bwd")
        out[f"{kind}_{ps}_{st}_y"] = y
        out[f"{kind}_{ps}_{st}_g"] = g
        out[f"{kind}_{ps}_{st}_dx"] = dX
# the reference's printed pool KATs (tests/polling_test.py, avg_polling_test.py): arange(1..16) 4x4
xk4 = np.arange(1, 17, dtype=np.float64).reshape(1, 1, 4, 4)
close(ref.MaxPoolingLayer(2, 2).forward(T(xk4)).data.reshape(2, 2), [[6, 8], [14, 16]], what="maxpool KAT")
close(ref.AveragePoolingLayer(2, 2).forward(T(xk4)).data.reshape(2, 2), [[3.5, 5.5], [11.5, 13.5]], what="avgpool KAT")
# tie rule (R4): all maxima receive the gradient
xt = np.zeros((1, 1, 2, 2))
close(O.maxpool_backward(xt, np.full((1, 1, 1, 1), 3.0)), np.full((1, 1, 2, 2), 3.0), what="tie rule")
save("pooling", x=xp_, kat_x=xk4, kat_max=np.array([[6, 8], [14, 16.]]), kat_avg=np.array([[3.5, 5.5], [11.5, 13.5]]),
     src_fwd="reference", src_bwd="oracle+torch", **out)

# ----------------------------------------------------------------------------- losses
p = O.softmax_forward(rng.standard_normal((10, 6)))
p[0, 0] = 0.0  # exercises the clip
t = rng.integers(0, 6, 10)
ce = ref.CrossEntropyLoss()
lv = ce.forward(T(p), T(t))
lg = ce.backward(T(p), T(t)).data
close(O.cross_entropy_forward(p, t), lv, what="ce fwd")
close(O.cross_entropy_backward(p, t), lg, what="ce bwd")
o_ = rng.standard_normal((10, 6))
tt = rng.standard_normal((10, 6))
ms = ref.MSELoss()
mv = ms.forward(T(o_), T(tt))
mg = ms.backward(T(o_), T(tt)).data
close(O.mse_forward(o_, tt), mv, what="mse fwd")
close(O.mse_backward(o_, tt), mg, what="mse bwd")
save("losses", p=p, t=t, ce=np.float64(lv), ce_grad=lg, o=o_, tt=tt, mse=np.float64(mv), mse_grad=mg, src="reference")

# ----------------------------------------------------------------------------- Unfold / Fold (tests/fold_test.py:33 round trip)
xu = rng.standard_normal((2, 3, 6, 6))
u = ref.Unfold(kernel_size=2, stride=2, padding=0).forward(T(xu)).data
close(O.unfold_forward(xu, 2, 2), u, what="unfold")
f = ref.Fold(output_size=(6, 6), kernel_size=2, stride=2, padding=0).forward(T(u)).data
close(f, xu, what="fold(unfold(x)) == x")
close(O.fold_forward(u, (6, 6), 2), xu, what="oracle fold")
u3 = ref.Unfold(kernel_size=3, stride=1, padding=0).forward(T(xu)).data
close(O.unfold_forward(xu, 3, 1), u3, what="unfold k3")
Wc = rng.standard_normal((5, 3, 3, 3))
yc = ref.Conv2DLayer(3, 5, 3)
yc.filters.data = Wc
close((Wc.reshape(5, -1) @ u3).reshape(2, 5, 4, 4), yc.forward(T(xu)).data, 1e-12, what="W.reshape @ unfold == conv")
save("unfold_fold", x=xu, unfold_k2s2=u, unfold_k3s1=u3, src="reference")
# Unfold with padding (the reference's top-left-aligned border patches), Unfold.backward (p = 0), Fold forward / backward
uf = ref.Unfold(3, 2, 1)
u_p = uf.forward(T(xu)).data
close(O.unfold_forward_loops(xu, 3, 2, 1), u_p, 0, what="unfold padding quirk")
uf0 = ref.Unfold(3, 1, 0)
u0 = uf0.forward(T(xu)).data
gu = rng.standard_normal(u0.shape)
dxu = uf0.backward(T(gu)).data
close(O.unfold_backward(gu, xu.shape, 3, 1, 0), dxu, 1e-13, what="unfold backward")
uf2 = ref.Unfold(2, 2, 0)
u2 = uf2.forward(T(xu)).data
gu2 = rng.standard_normal(u2.shape)
dxu2 = uf2.backward(T(gu2)).data
close(O.unfold_backward(gu2, xu.shape, 2, 2, 0), dxu2, 1e-13, what="unfold backward s2")
try:
    uf.backward(T(rng.standard_normal(u_p.shape)))
    unfold_pad_bwd_raises = False
except ValueError:
    unfold_pad_bwd_raises = True
assert unfold_pad_bwd_raises
fo = ref.Fold((6, 6), 2, 2, 0)
f_out = fo.forward(T(u2)).data
close(O.fold_forward_ref(u2, (6, 6), 2, 2, 0), f_out, 0, what="fold forward")
gf = rng.standard_normal(f_out.shape)
f_bwd = fo.backward(T(gf)).data
close(O.fold_backward(gf, u2.shape, (6, 6), 2, 2, 0), f_bwd, 0, what="fold backward")
save("unfold_fold2", x=xu, unfold_k3s2p1=u_p, g_k3s1=gu, dx_k3s1=dxu, g_k2s2=gu2, dx_k2s2=dxu2, fold_in=u2, fold_out=f_out, fold_g=gf,
     fold_bwd=f_bwd, src="reference")

# ----------------------------------------------------------------------------- the other activations
xa = rng.standard_normal((6, 7)) * 2.0
xa.flat[0] = -0.3        # np.vectorize infers the dtype of `lambda x: 1 if x > 0 else alpha` from element 0: keep it float
ga = rng.standard_normal((6, 7))
acts = {}
for nm, lay, f, d in (("elu", ref.ELUActivationLayer(0.7), lambda v: O.elu(v, 0.7), lambda v: O.elu_d(v, 0.7)),
                      ("selu", ref.SELUActivationLayer(), O.selu, O.selu_d),
                      ("leaky", ref.LeakyReLUActivationLayer(0.05), lambda v: O.leaky(v, 0.05), lambda v: O.leaky_d(v, 0.05)),
                      ("swish", ref.SwishActivationLayer(), O.swish, O.swish_d),
                      ("gelu", ref.GELUActivationLayer(), O.gelu, O.gelu_d),
                      ("softsign", ref.SoftsignActivationLayer(), O.softsign, O.softsign_d),
                      ("hardtanh", ref.HardTanhActivationLayer(), O.hardtanh, O.hardtanh_d)):
    y = lay.forward(T(xa.copy())).data
    dx = lay.backward(T(ga.copy()), 0.1)
    dx = dx.data if hasattr(dx, "data") else dx
    close(f(xa), y, 1e-13, what=nm + " forward")
    close(ga * d(xa), dx, 1e-13, what=nm + " backward")
    acts[nm + "_y"], acts[nm + "_dx"] = np.asarray(y, dtype=np.float64), np.asarray(dx, dtype=np.float64)
pl = ref.PReLUActivationLayer(0.05)
ypl = pl.forward(T(xa.copy())).data
dpl = pl.backward(T(ga.copy()), 0.1).data
close(O.leaky(xa, 0.05), ypl, 1e-13, what="prelu forward")
close(ga * O.leaky_d(xa, 0.05), dpl, 1e-13, what="prelu backward")
close(np.array(O.prelu_alpha_grad(xa, ga)), np.array(pl.d_alpha), 1e-12, what="prelu d_alpha")
close(np.array(0.05 - 0.1 * O.prelu_alpha_grad(xa, ga)), np.array(pl.alpha), 1e-12, what="prelu alpha update")
sm_l = ref.SoftminActivationLayer()
ysm = sm_l.forward(T(xa.copy())).data
dsm = sm_l.backward(T(ga.copy())).data
close(O.softmin_forward(xa), ysm, 1e-13, what="softmin forward")
close(O.softmax_backward(ysm, ga), dsm, 1e-12, what="softmin backward")
ls_l = ref.LogSoftmaxActivationLayer()
yls = ls_l.forward(T(xa.copy())).data
dls = ls_l.backward(T(ga.copy())).data
close(O.logsoftmax_forward(xa), yls, 1e-13, what="logsoftmax forward")
close(O.logsoftmax_backward(yls, ga), dls, 1e-12, what="logsoftmax backward")
save("activations2", x=xa, g=ga, prelu_y=ypl, prelu_dx=dpl, prelu_d_alpha=np.array(pl.d_alpha), prelu_alpha_after=np.array(pl.alpha),
     softmin_y=ysm, softmin_dx=dsm, logsoftmax_y=yls, logsoftmax_dx=dls, src="reference", **acts)

# ----------------------------------------------------------------------------- Conv2DLayer(transposed=True)
import torch
import torch.nn.functional as TF
tconv = {}
for ci, (Ci, Co, k, st, pd, hw) in enumerate(((3, 4, 3, 1, 0, (5, 6)), (2, 5, (3, 2), 2, 0, (6, 5)), (4, 3, 4, (2, 3), 0, (4, 4)))):
    lay = ref.Conv2DLayer(Ci, Co, k, st, pd, transposed=True)
    Ft = rng.standard_normal(lay.filters.data.shape) * 0.3
    bt = rng.standard_normal((Co, 1)) * 0.2
    lay.filters.data, lay.biases.data = Ft.copy(), bt.copy()
    xt_ = rng.standard_normal((2, Ci) + hw)
    yr = lay.forward(T(xt_.copy())).data
    close(O.conv2d_transposed_forward(xt_, Ft, bt, st, pd), yr, 1e-13, what=f"transposed conv forward {ci}")
    gt_ = rng.standard_normal(yr.shape)
    try:
        lay.backward(T(gt_.copy()), 0.1)
        raised = False
    except ValueError:
        raised = True
    assert raised, "shipped backward_transposed was expected to raise (okl.py:904-912 broadcast)"
    xt64 = torch.tensor(xt_, requires_grad=True)
    Ft64 = torch.tensor(Ft, requires_grad=True)
    bt64 = torch.tensor(bt.reshape(-1), requires_grad=True)
    yt = TF.conv_transpose2d(xt64, Ft64, bt64, stride=st, padding=pd)
    close(yt.detach().numpy(), yr, 1e-12, what=f"transposed conv forward vs torch {ci}")
    yt.backward(torch.tensor(gt_))
    dF, db_, dX_ = O.conv2d_transposed_backward(xt_, Ft, gt_, st, pd)
    close(dF, Ft64.grad.numpy(), 1e-12, what=f"transposed dF vs torch {ci}")
    close(db_.reshape(-1), bt64.grad.numpy(), 1e-12, what=f"transposed db vs torch {ci}")
    close(dX_, xt64.grad.numpy(), 1e-12, what=f"transposed dX vs torch {ci}")
    tconv.update({f"x{ci}": xt_, f"F{ci}": Ft, f"b{ci}": bt, f"y{ci}": yr, f"g{ci}": gt_, f"dF{ci}": dF, f"db{ci}": db_, f"dX{ci}": dX_,
                  f"cfg{ci}": np.array([Ci, Co] + list(O._tup(k, 2)) + list(O._tup(st, 2)) + list(O._tup(pd, 2)))})
try:                                              # padding > 0: the shipped forward allocates the CROPPED output and then scatters into it
    ref.Conv2DLayer(2, 3, 3, 1, 1, transposed=True).forward(T(rng.standard_normal((1, 2, 4, 4))))
    raised = False
except ValueError:
    raised = True
assert raised, "forward_transposed with padding > 0 was expected to raise"
save("conv2d_transposed", src="reference forward; backward = repaired loops cross-checked with torch.autograd", **tconv)

# ----------------------------------------------------------------------------- un-pooling / padding layers
xu_ = rng.standard_normal((2, 3, 4, 5))
up = {}
for nm, cls, avg in (("max", ref.MaxUnpoolingLayer, False), ("avg", ref.AverageUnpoolingLayer, True)):
    for tag, ps, st in (("a", 2, 2), ("b", 3, 2), ("c", 2, 3)):
        lay = cls(ps, st)
        y = lay.forward(T(xu_.copy())).data
        close(O.unpool_forward(xu_, ps, st, avg), y, 0, what=f"{nm} unpool forward {tag}")
        gg_ = rng.standard_normal(y.shape)
        d = lay.backward(T(gg_.copy())).data
        close(O.unpool_backward(gg_, xu_.shape, ps, st, avg), d, 1e-13, what=f"{nm} unpool backward {tag}")
        up.update({f"{nm}_{tag}_y": y, f"{nm}_{tag}_g": gg_, f"{nm}_{tag}_dx": d})
for nm, cls, refl in (("zero", ref.ZeroPaddingLayer, False), ("reflect", ref.ReflectionPaddingLayer, True)):
    for tag, pd in (("a", 1), ("b", (2, 1)), ("c", (1, 2, 3, 1))):
        lay = cls(pd)
        y = lay.forward(T(xu_.copy())).data
        close(O.pad_forward(xu_, lay.padding, refl), y, 0, what=f"{nm} pad forward {tag}")
        gg_ = rng.standard_normal(y.shape)
        d = lay.backward(T(gg_.copy())).data
        close(O.pad_backward(gg_, lay.padding), d, 0, what=f"{nm} pad backward {tag}")
        up.update({f"{nm}_{tag}_y": y, f"{nm}_{tag}_g": gg_, f"{nm}_{tag}_dx": d})
save("unpool_pad", x=xu_, src="reference", **up)

# ----------------------------------------------------------------------------- RNNLayer
rl = ref.RNNLayer(6, 5)
rWh, rWx, rb = rng.standard_normal((5, 5)) * 0.4, rng.standard_normal((6, 5)) * 0.4, rng.standard_normal((1, 5)) * 0.1
rl.Wh.data, rl.Wx.data, rl.b.data = rWh.copy(), rWx.copy(), rb.copy()
orl = O.OracleRNN(rWh, rWx, rb)
xs_ = [rng.standard_normal((4, 6)) for _ in range(3)]
gs_ = [rng.standard_normal((4, 5)) for _ in range(3)]
rn = {}
hp, hpo = None, None
for t_, (xv, gv) in enumerate(zip(xs_, gs_)):
    h_r = rl.forward(T(xv.copy()), None if hp is None else T(hp.copy())).data
    h_o = orl.forward(xv, hpo)
    close(h_o, h_r, 1e-13, what=f"rnn forward {t_}")
    dx_r, dp_r = rl.backward(T(gv.copy()), 0.1)
    dx_o, dp_o = orl.backward(gv)
    close(dx_o, dx_r.data, 1e-13, what=f"rnn dx {t_}")
    close(dp_o, dp_r.data, 1e-13, what=f"rnn dprev {t_}")
    rn.update({f"x{t_}": xv, f"g{t_}": gv, f"h{t_}": h_r, f"dx{t_}": dx_r.data, f"dprev{t_}": dp_r.data})
    hp, hpo = h_r, h_o
close(orl.gWh, rl.Wh.grad, 1e-13, what="rnn Wh.grad")
close(orl.gWx, rl.Wx.grad, 1e-13, what="rnn Wx.grad")
close(orl.gb, rl.b.grad, 1e-13, what="rnn b.grad")
close(rWh, rl.Wh.data, 0, what="rnn weights untouched by backward")
save("rnn", Wh=rWh, Wx=rWx, b=rb, gWh=np.asarray(rl.Wh.grad), gWx=np.asarray(rl.Wx.grad), gb=np.asarray(rl.b.grad), src="reference", **rn)

# ----------------------------------------------------------------------------- LSTMLayer
ll = ref.LSTMLayer(6, 5)
lW = [rng.standard_normal((11, 5)) * 0.4 for _ in range(4)]
lb = [rng.standard_normal((1, 5)) * 0.1 for _ in range(4)]
for nm_, wv in zip(("Wf", "Wi", "Wc", "Wo"), lW):
    getattr(ll, nm_).data = wv.copy()
for nm_, bv in zip(("bf", "bi", "bc", "bo"), lb):
    getattr(ll, nm_).data = bv.copy()
ol = O.OracleLSTM(lW, lb, 6)
ls_ = {}
hp = cp = hpo = cpo = None
for t_ in range(3):
    xv, gh, gc = rng.standard_normal((4, 6)), rng.standard_normal((4, 5)), rng.standard_normal((4, 5))
    h_r, c_r = ll.forward(T(xv.copy()), None if hp is None else T(hp.copy()), None if cp is None else T(cp.copy()))
    h_o, c_o = ol.forward(xv, hpo, cpo)
    close(h_o, h_r.data, 1e-13, what=f"lstm hidden {t_}")
    close(c_o, c_r.data, 1e-13, what=f"lstm cell {t_}")
    dh_r, dc_r = ll.backward(T(gh.copy()), T(gc.copy()), 0.1)
    dh_o, dc_o = ol.backward(gh, gc)
    close(dh_o, dh_r.data, 1e-13, what=f"lstm dprev_h {t_}")
    close(dc_o, dc_r.data, 1e-13, what=f"lstm dprev_c {t_}")
    ls_.update({f"x{t_}": xv, f"gh{t_}": gh, f"gc{t_}": gc, f"h{t_}": h_r.data, f"c{t_}": c_r.data, f"dh{t_}": dh_r.data, f"dc{t_}": dc_r.data})
    hp, cp, hpo, cpo = h_r.data, c_r.data, h_o, c_o
for k_, nm_ in enumerate(("Wf", "Wi", "Wc", "Wo")):
    close(ol.gW[k_], getattr(ll, nm_).grad, 1e-13, what="lstm " + nm_ + ".grad")
    close(lW[k_], getattr(ll, nm_).data, 0, what="lstm weights untouched")
    ls_["g" + nm_] = np.asarray(getattr(ll, nm_).grad)
for k_, nm_ in enumerate(("bf", "bi", "bc", "bo")):
    close(ol.gb[k_], getattr(ll, nm_).grad, 1e-13, what="lstm " + nm_ + ".grad")
    ls_["g" + nm_] = np.asarray(getattr(ll, nm_).grad)
save("lstm", Wf=lW[0], Wi=lW[1], Wc=lW[2], Wo=lW[3], bf=lb[0], bi=lb[1], bc=lb[2], bo=lb[3], src="reference", **ls_)

# ----------------------------------------------------------------------------- GRULayer
gl = ref.GRULayer(6, 5)
gW = [rng.standard_normal((11, 5)) * 0.4 for _ in range(3)]
gb_ = [rng.standard_normal((1, 5)) * 0.1 for _ in range(3)]
for nm_, wv in zip(("Wz", "Wr", "Wh"), gW):
    getattr(gl, nm_).data = wv.copy()
for nm_, bv in zip(("bz", "br", "bh"), gb_):
    getattr(gl, nm_).data = bv.copy()
og = O.OracleGRU(gW, gb_, 6)
gr_ = {}
hp = hpo = None
for t_ in range(3):
    xv, gh = rng.standard_normal((4, 6)), rng.standard_normal((4, 5))
    h_r = gl.forward(T(xv.copy()), None if hp is None else T(hp.copy())).data
    h_o = og.forward(xv, hpo)
    close(h_o, h_r, 1e-13, what=f"gru hidden {t_}")
    dx_r, dp_r = gl.backward(T(gh.copy()), 0.1)
    dx_o, dp_o = og.backward(gh)
    close(dx_o, dx_r.data, 1e-13, what=f"gru dx {t_}")
    close(dp_o, dp_r.data, 1e-13, what=f"gru dprev {t_}")
    gr_.update({f"x{t_}": xv, f"gh{t_}": gh, f"h{t_}": h_r, f"dx{t_}": dx_r.data, f"dprev{t_}": dp_r.data})
    hp, hpo = h_r, h_o
for k_, nm_ in enumerate(("Wz", "Wr", "Wh")):
    close(og.gW[k_], getattr(gl, nm_).grad, 1e-13, what="gru " + nm_ + ".grad")
    close(gW[k_], getattr(gl, nm_).data, 0, what="gru weights untouched")
    gr_["g" + nm_] = np.asarray(getattr(gl, nm_).grad)
for k_, nm_ in enumerate(("bz", "br", "bh")):
    close(og.gb[k_], getattr(gl, nm_).grad, 1e-13, what="gru " + nm_ + ".grad")
    gr_["g" + nm_] = np.asarray(getattr(gl, nm_).grad)
save("gru", Wz=gW[0], Wr=gW[1], Wh=gW[2], bz=gb_[0], br=gb_[1], bh=gb_[2], src="reference", **gr_)

# ----------------------------------------------------------------------------- BatchNormLayer
xb1, xb2 = rng.standard_normal((12, 5)) * 1.5 + 0.3, rng.standard_normal((12, 5)) * 0.7 - 0.2
gb1, gb2 = rng.standard_normal((12, 5)), rng.standard_normal((12, 5))
rb, ob = ref.BatchNormLayer(5), O.OracleBatchNorm(5)
bn = {}
for step, (xv, gv) in enumerate(((xb1, gb1), (xb2, gb2))):
    yr, yo = rb.forward(T(xv.copy())).data, ob.forward(xv)
    close(yo, yr, 1e-13, what=f"batchnorm forward {step}")
    dr, do = rb.backward(T(gv.copy()), 0.05).data, ob.backward(gv, 0.05)
    close(do, dr, 1e-12, what=f"batchnorm backward {step}")
    close(ob.gamma, rb.gamma.data, 1e-13, what="batchnorm gamma")
    close(ob.beta, rb.beta.data, 1e-13, what="batchnorm beta")
    bn.update({f"y{step}": yr, f"dx{step}": dr, f"gamma{step}": rb.gamma.data.copy(), f"beta{step}": rb.beta.data.copy(),
               f"rm{step}": rb.running_mean.copy(), f"rv{step}": rb.running_var.copy()})
rb.training = ob.training = False
close(ob.forward(xb1), rb.forward(T(xb1.copy())).data, 1e-13, what="batchnorm eval forward")
bn["y_eval"] = rb.forward(T(xb1.copy())).data
save("batchnorm", x0=xb1, x1=xb2, g0=gb1, g1=gb2, src="reference", **bn)
# regularisation wrappers around BatchNorm (two steps each) and Dropout with a seeded mask
reg = {}
for nm, cls in (("l1", ref.L1RegularizationLayer), ("l2", ref.L2RegularizationLayer), ("l3", ref.L3RegularizationLayer)):
    lay = cls(ref.BatchNormLayer(5), 0.01)
    for step, (xv, gv) in enumerate(((xb1, gb1), (xb2, gb2))):
        lay.forward(T(xv.copy()))
        reg[f"{nm}_dx{step}"] = lay.backward(T(gv.copy()), 0.05).data
    reg[f"{nm}_gamma"], reg[f"{nm}_beta"] = lay.layer.gamma.data.copy(), lay.layer.beta.data.copy()
    reg[f"{nm}_ggrad"], reg[f"{nm}_bgrad"] = np.asarray(lay.layer.gamma.grad).copy(), np.asarray(lay.layer.beta.grad).copy()
np.random.seed(123)
dl = ref.DropoutLayer(0.3)
yd = dl.forward(T(xb1.copy())).data
dd = dl.backward(T(gb1.copy()), 0.05).data
save("reg_dropout", x0=xb1, x1=xb2, g0=gb1, g1=gb2, drop_y=yd, drop_dx=dd, drop_mask=dl.mask, src="reference", **reg)

# ----------------------------------------------------------------------------- optimizers / schedulers
w0 = rng.standard_normal((5, 4))
gs = [rng.standard_normal((5, 4)) for _ in range(3)]
ad = ropt.AdamOptimizer(lr=0.01)
wa = w0.copy()
st = O.AdamState()
wo = w0.copy()
for g in gs:
    ad.update(wa, g, "k")
    wo = O.adam_update(wo, g, "k", st, lr=0.01)
close(wo, wa, what="adam")
sg = ropt.SGDOptimizer(lr=0.01, momentum=0.9)
ws = w0.copy()
v = None
wso = w0.copy()
for g in gs:
    try:
        sg.update(ws, g, "k")
    except AttributeError:
        pass            # optimizers.py:146 touches ndarray.grad after the update was applied
    wso, v = O.sgd_update(wso, v, g, 0.01, 0.9)
close(wso, ws, what="sgd")
sched = {
    "const": [ref.LearningRateScheduler(0.1).step(e) for e in range(6)],
    "step": [ref.StepLRScheduler(0.1, 2, 0.5).step(e) for e in range(6)],
    "time": [ref.TimeBasedDecayScheduler(0.1, 0.3).step(e) for e in range(6)],
    "exp": [ref.ExponentialDecayScheduler(0.1, 0.3).step(e) for e in range(6)],
    "lin": [ref.LinearDecayScheduler(0.1, 0.01, 4).step(e) for e in range(6)],
}
for e in range(6):
    close(O.lr_step(0.1, 2, 0.5, e), sched["step"][e])
    close(O.lr_time_based(0.1, 0.3, e), sched["time"][e])
    close(O.lr_exponential(0.1, 0.3, e), sched["exp"][e])
    close(O.lr_linear(0.1, 0.01, 4, e), sched["lin"][e])
save("optim_sched", w0=w0, g=np.stack(gs), adam_w=wa, sgd_w=ws, **{"sched_" + k: np.array(v_) for k, v_ in sched.items()},
     src="reference")

# ----------------------------------------------------------------------------- NeuralNetwork.train on a Dense net (shipped code end to end)
import logging  # noqa: E402

logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)
np.random.seed(7)
Xd = rng.standard_normal((25, 6))
Xd[:, 0] = np.abs(Xd[:, 0]) + 0.1
Td = rng.integers(0, 3, 25)
Wa, Wb = rng.standard_normal((6, 8)) * 0.3, rng.standard_normal((8, 3)) * 0.3
Wa[0, :] = np.abs(Wa[0, :])
net = ref.NeuralNetwork(temperature=1.0)
d1, d2 = ref.DenseLayer(6, 8), ref.DenseLayer(8, 3)
d1.weights.data, d2.weights.data = Wa.copy(), Wb.copy()
net.add(d1)
net.add(ref.TanhActivationLayer())
net.add(d2)
net.add(ref.SoftmaxActivationLayer())
xin, tin = T(Xd.copy()), T(Td.copy())
np.random.seed(11)
losses = net.train(xin, tin, epochs=3, lr_scheduler=ref.StepLRScheduler(0.05, 1, 0.5),
                   optimizer=ropt.SGDOptimizer(lr=0.05), batch_size=10, loss_function=ref.CrossEntropyLoss())
# oracle replay with the same shuffles
np.random.seed(11)
on = O.OracleNetwork([O.OracleDense(Wa, np.zeros((1, 8))), O.OracleTanh(), O.OracleDense(Wb, np.zeros((1, 3))), O.OracleSoftmax()])
Xo, To = Xd.copy(), Td.copy()
ol = []
for ep in range(3):
    lr = O.lr_step(0.05, 1, 0.5, ep)
    idx = np.arange(25)
    np.random.shuffle(idx)
    Xo, To = Xo[idx], To[idx]
    tot = 0.0
    for s0 in range(0, 25, 10):
        tot += on.train_step(Xo[s0:s0 + 10], To[s0:s0 + 10], O.OracleCrossEntropy(), lr)
    ol.append(tot / 3)
close(ol, losses, 1e-12, what="train losses")
close(on.layers[0].W, d1.weights.data, 1e-12, what="train W1")
close(on.layers[2].W, d2.weights.data, 1e-12, what="train W2")
close(Xo, xin.data, what="dataset reordered in place")
save("train_dense", X=Xd, T=Td, Wa=Wa, Wb=Wb, losses=np.array(losses), W1=d1.weights.data, b1=d1.biases.data,
     W2=d2.weights.data, b2=d2.biases.data, X_after=xin.data, T_after=tin.data, seed=np.int64(11), src="reference")

# ----------------------------------------------------------------------------- small MNIST-shaped CNN, two train steps (oracle; conv/pool bwd repaired)
N = 8
Xc = rng.standard_normal((N, 1, 12, 12)).astype(np.float32)
Tc = rng.integers(0, 10, N)
Wc1 = rng.standard_normal((4, 1, 3, 3)) * 0.3
Wd1 = rng.standard_normal((4 * 5 * 5, 16)) * 0.1
Wd2 = rng.standard_normal((16, 10)) * 0.1


def mk(fast):
    return O.OracleNetwork([O.OracleConv(Wc1, np.zeros((4, 1)), 1, 0, fast=fast), O.OracleReLU(), O.OracleMaxPool(2, 2),
                            O.OracleFlatten(), O.OracleDense(Wd1, np.zeros((1, 16))), O.OracleReLU(),
                            O.OracleDense(Wd2, np.zeros((1, 10))), O.OracleSoftmax()])


na, nb = mk(False), mk(True)
l_a = [na.train_step(Xc.astype(np.float64), Tc, O.OracleCrossEntropy(), 0.05) for _ in range(2)]
l_b = [nb.train_step(Xc.astype(np.float64), Tc, O.OracleCrossEntropy(), 0.05) for _ in range(2)]
close(l_a, l_b, 1e-12, what="cnn loops vs fast")
close(na.layers[0].W, nb.layers[0].W, 1e-12, what="cnn conv W loops vs fast")
# cross-check against torch autograd with the accumulate-forever rule
tw = [torch.tensor(a, requires_grad=True) for a in (Wc1, np.zeros(4), Wd1, np.zeros(16), Wd2, np.zeros(10))]
acc = [torch.zeros_like(w) for w in tw]
tl = []
for _ in range(2):
    h = F.max_pool2d(F.relu(F.conv2d(torch.tensor(Xc.astype(np.float64)), tw[0], tw[1])), 2, 2).reshape(N, -1)
    h = F.relu(h @ tw[2] + tw[3]) @ tw[4] + tw[5]
    # The reference feeds (clip(p) - onehot)/N (okl.py:1399-1403) through the softmax Jacobian
    # (okl.py:376-393) - not the analytic dCE/dlogits - so mimic exactly that on the logits.
    pr = torch.softmax(h, 1).detach()
    prc = pr.clamp(1e-12, 1 - 1e-12)
    loss = -torch.log(prc[range(N), torch.tensor(Tc)]).mean()
    gp = prc.clone()
    gp[range(N), torch.tensor(Tc)] -= 1
    gp /= N
    glogits = pr * (gp - (gp * pr).sum(1, keepdim=True))
    grads = torch.autograd.grad(h, tw, grad_outputs=glogits)
    tl.append(float(loss))
    with torch.no_grad():
        for w, a, g in zip(tw, acc, grads):
            a += g
            w -= 0.05 * a
close(tl, l_a, 1e-9, what="cnn vs torch loss")
close(tw[0].detach().numpy(), na.layers[0].W, 1e-9, what="cnn vs torch conv W")
close(tw[2].detach().numpy(), na.layers[4].W, 1e-9, what="cnn vs torch dense W")
save("train_cnn_small", X=Xc, T=Tc, Wc1=Wc1, Wd1=Wd1, Wd2=Wd2, lr=np.float64(0.05), losses=np.array(l_a),
     conv_W=na.layers[0].W, conv_b=na.layers[0].b, d1_W=na.layers[4].W, d1_b=na.layers[4].b, d2_W=na.layers[6].W,
     d2_b=na.layers[6].b, probs=na.acts[-1], src="oracle+torch")
print("all golden vectors generated and cross-checked")
