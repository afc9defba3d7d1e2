"""GPU tests of the tcgen05 (TF32 / BF16) tensor-core path against the fp64 oracle at the north_star tolerance 2e-2
(relative, max|a-b| / max|b|); in practice TF32 lands near 1e-3 and BF16 near 1e-2, asserted a bit tighter below."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from oracle import okl_oracle as O

pytestmark = pytest.mark.gpu
TOL = {0: 3e-3, 1: 1.5e-2}


def tc_gemm(okl, kind, a_mn, b_mn, A, B, bias=None, act=0):
    """A: (M,K) logical, B: (K,N) logical; storage chosen by the major flags."""
    from okrolearn_b200 import _lib
    ctx = okl.get_ctx()
    M, K = A.shape
    N = B.shape[1]
    As = np.ascontiguousarray(A.T if a_mn else A, dtype=np.float32)          # [K][M] or [M][K]
    Bs = np.ascontiguousarray(B if b_mn else B.T, dtype=np.float32)          # [K][N] or [N][K]
    tA, tB = okl.Tensor(As), okl.Tensor(Bs)
    D = okl.Tensor._empty((M, N), np.float32)
    tb = okl.Tensor(bias.astype(np.float32)) if bias is not None else None
    _lib.check(_lib.lib.okl_tc_gemm(ctx.h, kind, a_mn, b_mn, M, N, K, tA._dptr(), As.shape[1], tB._dptr(), Bs.shape[1],
                                    D._ptr_raw(), N, tb._dptr() if tb is not None else None, act))
    return D.data


SHAPES = [(128, 128, 32), (128, 256, 64), (256, 128, 5408), (256, 5408, 128), (5408, 128, 256), (200, 72, 100), (1024, 1024, 1024),
          (130, 36, 40), (64, 64, 64), (384, 640, 96), (8192, 512, 1024)]


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_all_majors(okl, kind, a_mn, b_mn, M, N, K):
    q = 8 if kind == 1 else 4
    lda = M if a_mn else K
    ldb = N if b_mn else K
    if lda % q or ldb % q:
        pytest.skip("leading dimension not a multiple of 16 bytes (TMA)")
    epc = 64 if kind == 1 else 32
    if (M if a_mn else K) < epc or (N if b_mn else K) < epc:
        pytest.skip("operand smaller than one 128-byte TMA box: FFMA kernels' job")
    rng = np.random.default_rng(M + 3 * N + 7 * K + a_mn * 11 + b_mn * 13)
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    bias = rng.standard_normal(N)
    ref = A @ B + bias
    got = tc_gemm(okl, kind, a_mn, b_mn, A, B, bias, 0)
    assert rel_err(got, ref) < TOL[kind]
    got = tc_gemm(okl, kind, a_mn, b_mn, A, B, None, 1)
    assert rel_err(got, np.maximum(A @ B, 0)) < TOL[kind]


@pytest.mark.parametrize("prec,tol", [("tf32", 3e-3), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("M,K,N", [(256, 5408, 128), (64, 128, 256), (300, 1000, 200), (1024, 1024, 1024), (256, 128, 10)])
def test_dense_layer_tensor_core_modes(okl, prec, tol, M, K, N):
    okl.set_precision(prec)
    try:
        rng = np.random.default_rng(M + K + N)
        x, W, b = rng.standard_normal((M, K)), rng.standard_normal((K, N)) * 0.1, rng.standard_normal((1, N))
        g = rng.standard_normal((M, N))
        lay = okl.DenseLayer(K, N)
        lay.weights.data, lay.biases.data = W, b
        y = lay.forward(okl.Tensor(x))
        assert rel_err(y.data, O.dense_forward(x, W, b)) < tol
        dX = lay.backward(okl.Tensor(g), 0.01)
        dW, db, dXo = O.dense_backward(x, W, g)
        assert rel_err(dX.data, dXo) < tol
        assert rel_err(lay.weights.grad, dW) < tol and rel_err(lay.biases.grad, db) < 1e-5
    finally:
        okl.set_precision("fp32")


def test_tc_gemm_is_deterministic(okl):
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((256, 5408)), rng.standard_normal((5408, 128))
    r1 = tc_gemm(okl, 0, 0, 1, A, B)
    r2 = tc_gemm(okl, 0, 0, 1, A, B)
    assert np.array_equal(r1, r2)            # split-K partials are reduced in a fixed order
