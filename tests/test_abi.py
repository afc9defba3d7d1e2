"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what include/okl.h declares, the ctypes
table matches the header, and the host-side mirror of the reference API behaves (no compute calls - there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "okl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int)\s+(okl_\w+)\s*\(", src, flags=re.M)
    assert len(names) > 40
    return names


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "okrolearn_b200", "libokl_b200.so"))
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/okl.h but not exported"


def test_ctypes_table_matches_header():
    from okrolearn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == sorted(header_functions())


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "okl.h")).read()
    for needle in ("okl.py:18-54", "okl.py:669-764", "okl.py:1715-1831", "okl.py:355-399", "optimizers.py:28-42"):
        assert needle in src


def test_version_and_error_strings():
    from okrolearn_b200 import _lib
    assert b"sm_100a" in _lib.lib.okl_version()
    assert isinstance(_lib.last_error(), str)


def test_no_gpu_fails_loudly_without_fallback():
    from okrolearn_b200 import _lib
    probe = ctypes.c_void_p()
    rc = _lib.lib.okl_ctx_create(0, 0, ctypes.byref(probe))
    if rc == 0:
        _lib.lib.okl_ctx_destroy(probe)
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.check(rc)
    import okrolearn_b200.okrolearn as m
    layer = m.DenseLayer(4, 3)
    with pytest.raises(RuntimeError):
        layer.forward(m.Tensor(np.zeros((2, 4))))


def test_conv_extent_and_validation_host_side():
    from okrolearn_b200 import _lib
    d = _lib.ConvDesc()
    d.nd, d.N, d.C, d.O = 2, 4, 3, 8
    for i, (n, k, s, p) in enumerate(((9, 3, 2, 1), (11, 3, 2, 1))):
        d.inp[i], d.k[i], d.s[i], d.p[i] = n, k, s, p
    out = (ctypes.c_int * 3)()
    assert _lib.lib.okl_conv_out_extent(ctypes.byref(d), out) == 0
    assert (out[0], out[1]) == (5, 6)                 # floor((9+2-3)/2)+1, floor((11+2-3)/2)+1
    d.k[0] = 20
    assert _lib.lib.okl_conv_out_extent(ctypes.byref(d), out) == _lib.OKL_ERR_INVALID
    with pytest.raises(ValueError, match="kernel larger"):
        _lib.check(_lib.OKL_ERR_INVALID)


def test_reference_api_surface():
    import okrolearn_b200.okrolearn as m
    for name in ("Tensor DenseLayer Conv1DLayer Conv2DLayer Conv3DLayer ReLUActivationLayer SigmoidActivationLayer "
                 "TanhActivationLayer SoftmaxActivationLayer MaxPoolingLayer AveragePoolingLayer FlattenLayer CrossEntropyLoss "
                 "MSELoss LearningRateScheduler StepLRScheduler TimeBasedDecayScheduler ExponentialDecayScheduler "
                 "LinearDecayScheduler SGDOptimizer AdamOptimizer NeuralNetwork").split():
        assert hasattr(m, name), name
    net = m.NeuralNetwork(temperature=2.0)
    for meth in ("add remove add_hook remove_hook forward backward train save load save_weights load_weights set_temperature "
                 "set_debug_mode get_layer_outputs get_gradients").split():
        assert callable(getattr(net, meth))
    assert sorted(net.hooks) == sorted(['pre_forward', 'post_forward', 'pre_backward', 'post_backward', 'pre_epoch', 'post_epoch'])
    with pytest.raises(ValueError):
        net.add_hook("nope", lambda: None)


def test_layer_construction_matches_reference_shapes_and_dtypes():
    import okrolearn_b200.okrolearn as m
    np.random.seed(0)
    d = m.DenseLayer(5, 3)
    assert d.weights.data.shape == (5, 3) and d.weights.data.dtype == np.float64 and d.biases.data.shape == (1, 3)
    np.random.seed(0)
    assert np.array_equal(d.weights.data, np.random.randn(5, 3) * 0.1)                 # same RNG draw as okl.py:26
    c1 = m.Conv1DLayer(2, 4, 3, stride=2, padding=1)
    assert c1.weights.data.shape == (4, 2, 3) and c1.weights.data.dtype == np.float32 and c1.biases.data.shape == (4, 1)
    c2 = m.Conv2DLayer(3, 8, (3, 2), stride=(1, 2), padding=1)
    assert c2.filters.data.shape == (8, 3, 3, 2) and c2.filters.data.dtype == np.float64
    assert c2.kernel_size == (3, 2) and c2.stride == (1, 2) and c2.padding == (1, 1)
    c3 = m.Conv3DLayer(1, 2, 3)
    assert c3.filters.data.shape == (2, 1, 3, 3, 3) and c3.filters.data.dtype == np.float32
    assert sorted(c2.get_params()) == ['biases', 'filters'] and sorted(c1.get_params()) == ['biases', 'weights']
    c2.filters.data = np.ones((8, 3, 3, 2), dtype=np.int64)                             # tests/conv2d.py:11 assigns ints
    assert c2.filters.data.dtype == np.int64
    t = m.Tensor([[1, 2], [3, 4]])
    assert t.data.dtype == np.int64 and t.shape == (2, 2) and t.grad is None and t.backward_fn is None


def test_schedulers_match_golden():
    import okrolearn_b200.okrolearn as m
    from conftest import golden
    g = golden("optim_sched")
    for e in range(6):
        assert m.LearningRateScheduler(0.1).step(e) == g["sched_const"][e]
        assert abs(m.StepLRScheduler(0.1, 2, 0.5).step(e) - g["sched_step"][e]) < 1e-16
        assert abs(m.TimeBasedDecayScheduler(0.1, 0.3).step(e) - g["sched_time"][e]) < 1e-16
        assert abs(m.ExponentialDecayScheduler(0.1, 0.3).step(e) - g["sched_exp"][e]) < 1e-16
        assert abs(m.LinearDecayScheduler(0.1, 0.01, 4).step(e) - g["sched_lin"][e]) < 1e-16


def test_product_never_imports_oracle_or_torch():
    pkg = os.path.join(ROOT, "okrolearn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/okl_oracle.py states", ""), fn
            assert not re.search(r"^\s*(import|from)\s+(torch|cupy|triton)", src, flags=re.M), fn
