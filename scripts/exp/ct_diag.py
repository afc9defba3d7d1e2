"""conv_tc_kernel: where does a tile's time go?  OKL_CT_DIAG bits: 1 = epilogue does nothing, 2 = no MMAs, 4 = activation slabs loaded once per stage"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import okrolearn_b200.okrolearn as okl
from okrolearn_b200 import _lib
from okrolearn_b200.tensor import _DevBuf
okl.set_precision("bf16")
ctx = okl.get_ctx(); lib, check = _lib.lib, _lib.check
def timeit(fn, reps=10):
    for _ in range(2): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): fn()
    return ctx.timer_stop() / reps * 1e3
N = 1024
for (Cc, Oc, S, k) in ((64, 64, 32, (3, 1)), (64, 64, 32, (3, 3)), (128, 128, 16, (3, 3)), (256, 256, 8, (3, 3))):
    d = _lib.ConvDesc(); d.nd, d.N, d.C, d.O = 2, N, Cc, Oc
    for i in range(2): d.inp[i], d.k[i], d.s[i], d.p[i] = S, k[i], 1, k[i] // 2
    d.act, d.x_layout, d.y_layout = 1, _lib.NXC, _lib.NXC
    x16, y16 = _DevBuf(ctx, N * Cc * S * S * 2), _DevBuf(ctx, N * Oc * S * S * 2)
    check(lib.okl_memset(ctx.h, x16.ptr, 0x3c, x16.nbytes))
    W = okl.Tensor((np.random.default_rng(0).standard_normal((Oc, Cc) + k) * 0.1).astype(np.float32))
    bias = okl.Tensor(np.zeros(Oc, dtype=np.float32))
    w2 = _DevBuf(ctx, W.size * 2)
    check(lib.okl_conv2_filter_banks(ctx.h, C.byref(d), W._dptr(), w2.ptr, None))
    out = []
    for pair in ("1", "0"):
        os.environ["OKL_CT_PAIR"] = pair
        for dg in (0, 1, 2, 7):
            os.environ["OKL_CT_DIAG"] = str(dg)
            t = timeit(lambda: check(lib.okl_conv2_fprop(ctx.h, C.byref(d), x16.ptr, w2.ptr, bias._dptr(), None, y16.ptr)))
            out.append(f"p{pair}d{dg} {t:6.1f}")
    os.environ["OKL_CT_DIAG"] = "0"; os.environ.pop("OKL_CT_PAIR", None)
    print(f"{Cc}->{Oc}@{S} k={k}: " + "  ".join(out), flush=True)
