import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import okrolearn_b200.okrolearn as okl
from okrolearn_b200 import _lib
ctx = okl.get_ctx(); lib, check = _lib.lib, _lib.check
lib.okl_debug_tc_timestamps.argtypes = [C.c_void_p, C.c_void_p]

def case(M, N, K, a, b, label):
    A = okl.Tensor._empty((K, M) if a else (M, K), np.float32); B = okl.Tensor._empty((K, N) if b else (N, K), np.float32)
    D = okl.Tensor._empty((M, N), np.float32)
    for t in (A, B): check(lib.okl_memset(ctx.h, t._ptr_raw(), 0x3c, t.size * 4))
    lda = M if a else K; ldb = N if b else K
    f = lambda: check(lib.okl_tc_gemm(ctx.h, 0, a, b, M, N, K, A._ptr_raw(), lda, B._ptr_raw(), ldb, D._ptr_raw(), N, None, 0))
    for _ in range(3): f()
    ctx.sync()
    # back-to-back
    ctx.timer_start()
    for _ in range(20): f()
    b2b = ctx.timer_stop() / 20 * 1e3
    # with an elementwise kernel in between (carveout switch)
    ctx.timer_start()
    for _ in range(20):
        check(lib.okl_scale(ctx.h, D._ptr_raw(), 1024, 1.0)); f()
    mixed = ctx.timer_stop() / 20 * 1e3
    ctx.timer_start()
    for _ in range(20): check(lib.okl_scale(ctx.h, D._ptr_raw(), 1024, 1.0))
    scale_only = ctx.timer_stop() / 20 * 1e3
    dbg = okl.Tensor._empty((148 * 8 * 2,), np.float32)
    check(lib.okl_memset(ctx.h, dbg._ptr_raw(), 0, dbg.size * 4))
    lib.okl_debug_tc_timestamps(ctx.h, dbg._ptr_raw())
    check(lib.okl_scale(ctx.h, D._ptr_raw(), 1024, 1.0)); f(); ctx.sync()
    lib.okl_debug_tc_timestamps(ctx.h, None)
    raw = np.empty(148 * 8, dtype=np.uint64)
    check(lib.okl_d2h_async_raw(ctx.h, raw.ctypes.data, dbg._ptr_raw(), raw.nbytes)); ctx.sync()
    ts = raw.reshape(148, 8)[:, :6].astype(np.int64)
    ts = ts[ts[:, 0] > 0]
    t0 = ts[:, 0].min()
    rel = (ts - t0) / 1e3
    print(f"{label}: b2b {b2b:.2f}us  mixed {mixed:.2f}us (scale alone {scale_only:.2f})  ctas {len(ts)}")
    names = ["start", "setup", "1st_full", "acc_rdy", "epi_done", "end"]
    print("   mean us:", {n: round(float(rel[:, i].mean()), 2) for i, n in enumerate(names)}, " max end", round(float(rel[:, 5].max()), 2))

case(256, 5408, 128, 0, 0, "dX 256x5408x128")
case(5408, 128, 256, 1, 1, "dW 5408x128x256")
case(256, 128, 5408, 0, 1, "fwd 256x128x5408")
case(1024, 1024, 1024, 0, 1, "1024^3")
