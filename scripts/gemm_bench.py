"""Time the tcgen05 GEMM (okl_tc_gemm) with CUDA events.  Usage: python scripts/gemm_bench.py [M N K kind a_mn b_mn]..."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import okrolearn_b200.okrolearn as okl
from okrolearn_b200 import _lib

ctx = okl.get_ctx()
lib, check = _lib.lib, _lib.check


def run(M, N, K, kind, a_mn, b_mn, reps=20, flush=False):
    A = okl.Tensor._empty((K, M) if a_mn else (M, K), np.float32)
    B = okl.Tensor._empty((K, N) if b_mn else (N, K), np.float32)
    D = okl.Tensor._empty((M, N), np.float32)
    check(lib.okl_memset(ctx.h, A._ptr_raw(), 0x3c, A.size * 4))
    check(lib.okl_memset(ctx.h, B._ptr_raw(), 0x3c, B.size * 4))
    lda = M if a_mn else K
    ldb = N if b_mn else K
    f = lambda: check(lib.okl_tc_gemm(ctx.h, kind, a_mn, b_mn, M, N, K, A._ptr_raw(), lda, B._ptr_raw(), ldb, D._ptr_raw(), N, None, 0))
    for _ in range(3):
        f()
    ctx.sync()
    ts = []
    if flush:
        for _ in range(reps):
            check(lib.okl_flush_l2(ctx.h))
            ctx.timer_start(); f(); ts.append(ctx.timer_stop())
        ms = float(np.median(ts))
    else:
        ctx.timer_start()
        for _ in range(reps):
            f()
        ms = ctx.timer_stop() / reps
    return ms


if __name__ == "__main__":
    cases = [(8192, 8192, 8192, 1, 0, 1), (8192, 8192, 8192, 1, 0, 0), (8192, 8192, 8192, 1, 1, 1), (8192, 8192, 8192, 0, 0, 1),
             (8192, 8192, 8192, 0, 0, 0), (4096, 4096, 4096, 1, 0, 0),
             (256, 128, 5408, 0, 0, 1), (5408, 128, 256, 0, 1, 1), (256, 5408, 128, 0, 0, 0)]
    out = []
    for (M, N, K, kind, a, b) in cases:
        for env in ([{}, {"OKL_TC_PAIR": "0"}] if M >= 4096 else [{}, {"OKL_TC_BN": "32"}, {"OKL_TC_BN": "64"}, {"OKL_TC_BN": "128"}]):
            for k, v in env.items():
                os.environ[k] = v
            try:
                ms = run(M, N, K, kind, a, b, reps=10 if M >= 4096 else 50, flush=M < 4096)
                r = {"M": M, "N": N, "K": K, "kind": "bf16" if kind else "tf32", "a_mn": a, "b_mn": b, "env": env, "ms": round(ms, 4),
                     "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
            except Exception as e:
                r = {"M": M, "N": N, "K": K, "kind": kind, "err": str(e)[:200]}
            for k in env:
                os.environ.pop(k)
            print(json.dumps(r), flush=True)
