import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import okrolearn_b200.okrolearn as okl
from okrolearn_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
okl.set_precision(prec)
ctx = okl.get_ctx(); lib, check = _lib.lib, _lib.check
lib.okl_debug_tc_timestamps.argtypes = [C.c_void_p, C.c_void_p]
N, Cc, H, O = 128, 64, 32, 64
rng = np.random.default_rng(0)
lay = okl.Conv2DLayer(Cc, O, 3, 1, 1)
x = okl.Tensor(rng.standard_normal((N, Cc, H, H)).astype(np.float32))
x._dptr(_lib.NXC)
for _ in range(3): y = lay.forward(x)
ctx.sync()
ctx.timer_start()
for _ in range(10): y = lay.forward(x)
print("fprop layer call avg us (incl. weight prep):", ctx.timer_stop() / 10 * 1e3)
dbg = okl.Tensor._empty(((148 * 8 + 64) * 2,), np.float32)
check(lib.okl_memset(ctx.h, dbg._ptr_raw(), 0, dbg.size * 4))
lib.okl_debug_tc_timestamps(ctx.h, dbg._ptr_raw())
y = lay.forward(x); ctx.sync()
lib.okl_debug_tc_timestamps(ctx.h, None)
raw = np.empty(148 * 8 + 64, dtype=np.uint64)
check(lib.okl_d2h_async_raw(ctx.h, raw.ctypes.data, dbg._ptr_raw(), raw.nbytes)); ctx.sync()
ts = raw[:148 * 8].reshape(148, 8)[:, :6].astype(np.int64)
t0 = ts[:, 0][ts[:, 0] > 0].min()
print("cta0 phases us:", [(int(v) - t0) / 1e3 for v in ts[0]])
tt = raw[148 * 8:148 * 8 + 48].astype(np.int64).reshape(24, 2)
print("per-tile (acc_ready, epi_done) us for CTA 0:")
for i, (a, b) in enumerate(tt):
    if a: print(i, round((a - t0) / 1e3, 2), round((b - t0) / 1e3, 2))
print("end max us", (ts[:, 5].max() - t0) / 1e3)
