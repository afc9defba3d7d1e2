"""Conv3DLayer with >= 8 channels: BF16 tcgen05 kernels vs the FFMA path (OKL_CT_3D=0), forward + backward of one layer"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for mode in ("1", "0"):
    os.environ["OKL_CT_3D"] = mode
    import importlib
    import okrolearn_b200.okrolearn as okl
    import okrolearn_b200.layers as L
    L._CT_3D = mode == "1"
    okl.set_precision("bf16")
    ctx = okl.get_ctx()
    for (N, Cc, sp, Oc) in ((8, 64, (8, 32, 32), 64), (4, 32, (8, 56, 56), 64), (8, 128, (4, 16, 16), 128)):
        rng = np.random.default_rng(0)
        x = okl.Tensor(rng.standard_normal((N, Cc) + sp).astype(np.float32))
        lay = okl.Conv3DLayer(Cc, Oc, 3, 1, 1)
        y = lay.forward(x)
        g = okl.Tensor(rng.standard_normal(y.shape).astype(np.float32))
        g._b16(1) if getattr(lay, "_conv2", False) else g._dptr()
        lay.backward(g, 0.0)
        ctx.sync()
        ts = []
        for _ in range(5):
            ctx.timer_start()
            y = lay.forward(x)
            lay.backward(g, 0.0)
            ts.append(ctx.timer_stop() * 1e3)
        fl = 3 * 2.0 * N * np.prod(sp) * Oc * Cc * 27
        t = np.median(ts)
        print(f"OKL_CT_3D={mode} conv3d {Cc}->{Oc} on {N}x{sp}: fwd+bwd {t:8.1f} us = {fl / t / 1e6:7.1f} TFLOP/s (tensor-core path: {bool(getattr(lay, '_conv2', False))})", flush=True)
