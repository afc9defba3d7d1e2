"""weight gradient WITHOUT the bias gradient (the pooling backward produced it): activation-slab kernel vs first-generation 128 x 256 tiles"""
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
for (Cc, Oc, S) in ((64, 128, 16), (128, 128, 16), (128, 256, 8), (256, 256, 8)):
    d = _lib.ConvDesc(); d.nd, d.N, d.C, d.O = 2, N, Cc, Oc
    for i in range(2): d.inp[i], d.k[i], d.s[i], d.p[i] = S, 3, 1, 1
    d.act, d.x_layout, d.y_layout = 0, _lib.NXC, _lib.NXC
    x16, g16 = _DevBuf(ctx, N * Cc * S * S * 2), _DevBuf(ctx, N * Oc * S * S * 2)
    for b in (x16, g16): check(lib.okl_memset(ctx.h, b.ptr, 0x3c, b.nbytes))
    dW, db = _DevBuf(ctx, Oc * Cc * 9 * 4), _DevBuf(ctx, Oc * 4)
    out = []
    for name, env, fn in (("x-slab", {"OKL_WG_GS": "0", "OKL_WG_ALL": "1"}, "conv2"), ("g-slab", {"OKL_WG_ALL": "1"}, "conv2"), ("first gen", {}, "gen1")):
        for k in ("OKL_WG_GS", "OKL_WG_ALL"): os.environ.pop(k, None)
        os.environ.update(env)
        for with_db in (False, True):
            dbp = db.ptr if with_db else None
            if fn == "conv2":
                if not lib.okl_conv2_wgrad_supported(ctx.h, C.byref(d)): continue
                t = timeit(lambda: check(lib.okl_conv2_wgrad(ctx.h, C.byref(d), x16.ptr, g16.ptr, dW.ptr, dbp)))
            else:
                def f():
                    check(lib.okl_conv_wgrad_bf16(ctx.h, C.byref(d), x16.ptr, g16.ptr, None, dW.ptr, None))
                    if with_db: check(lib.okl_channel_sum_bf16(ctx.h, g16.ptr, db.ptr, N * S * S, Oc))
                t = timeit(f)
            out.append(f"{name}{'+db' if with_db else ''} {t:6.1f}")
    for k in ("OKL_WG_GS", "OKL_WG_ALL"): os.environ.pop(k, None)
    print(f"{Cc}->{Oc}@{S}: " + "  ".join(out), flush=True)
