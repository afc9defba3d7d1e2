"""Parity + timing probe of the second-generation tcgen05 convolution (okl_conv_tc.cu) through the C ABI.
   python scripts/conv2_probe.py [parity] [perf] [timeline]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import okrolearn_b200.okrolearn as okl
from okrolearn_b200 import _lib
from okrolearn_b200.tensor import _DevBuf
from oracle import okl_oracle as O

okl.set_precision("bf16")
ctx = okl.get_ctx()
lib, check = _lib.lib, _lib.check
what = sys.argv[1:] or ["parity", "perf"]


def desc(nd, N, Cc, Oc, sp, k, p, act=0):
    d = _lib.ConvDesc()
    d.nd, d.N, d.C, d.O = nd, N, Cc, Oc
    kt, pt = O._tup(k, nd), O._tup(p, nd)
    for i in range(nd):
        d.inp[i], d.k[i], d.s[i], d.p[i] = sp[i], kt[i], 1, pt[i]
    d.act, d.x_layout, d.y_layout = act, _lib.NXC, _lib.NXC
    return d


def bf16_round(a):
    """round-to-nearest-even fp32 -> bf16 -> fp32 on the host"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def fetch16(buf, shape_nxc):
    """bf16 device buffer (channels-last) -> fp32 ndarray in the reference (N, C, *sp) order"""
    n = int(np.prod(shape_nxc))
    raw = np.empty(n, dtype=np.uint16)
    check(lib.okl_d2h_async_raw(ctx.h, raw.ctypes.data, buf, n * 2))
    ctx.sync()
    f = (raw.astype(np.uint32) << 16).view(np.float32).reshape(shape_nxc)
    return np.moveaxis(f, -1, 1)


PAR = [  # nd, N, C, spatial, O, k, p
    (2, 8, 64, (32, 32), 64, 3, 1), (2, 4, 64, (16, 16), 128, 3, 1), (2, 6, 128, (8, 8), 256, 3, 1), (2, 5, 256, (8, 8), 256, 3, 1),
    (2, 3, 64, (13, 11), 64, 3, 0), (2, 2, 64, (9, 20), 128, (3, 5), (1, 2)), (2, 2, 64, (8, 8), 64, 1, 0), (1, 4, 64, (256,), 64, 3, 1),
    (2, 3, 128, (16, 16), 128, 3, 1), (2, 2, 64, (40, 36), 64, 3, 1), (2, 9, 64, (4, 4), 64, 3, 1), (2, 2, 64, (64, 64), 64, 5, 2),
]


def parity():
    bad = 0
    for (nd, N, Cc, sp, Oc, k, p) in PAR:
        for env in ({"OKL_CT_SLAB": "1"}, {"OKL_CT_SLAB": "0"}, {"OKL_CT_SLAB": "1", "OKL_CT_BN": "64"}, {"OKL_CT_SLAB": "1", "OKL_CT_BN": "256"},
                    {"OKL_CT_SLAB": "1", "OKL_CT_PAIR": "0"}, {"OKL_CT_SLAB": "0", "OKL_CT_PAIR": "0"}, {"OKL_CT_SLAB": "0", "OKL_CT_PAIR": "1", "OKL_CT_BN": "64"}):
            for k_ in ("OKL_CT_BN", "OKL_CT_CLUSTER", "OKL_CT_PAIR", "OKL_CT_SLAB"):
                os.environ.pop(k_, None)
            os.environ.update(env)
            rng = np.random.default_rng(N * 100 + Cc + Oc)
            kt = O._tup(k, nd)
            x = bf16_round(rng.standard_normal((N, Cc) + sp).astype(np.float32))
            W = (rng.standard_normal((Oc, Cc) + kt) * 0.1).astype(np.float32)
            b = (rng.standard_normal((Oc, 1)) * 0.1).astype(np.float32)
            d = desc(nd, N, Cc, Oc, sp, k, p, act=1)
            if not lib.okl_conv2_supported(ctx.h, C.byref(d)):
                print("unsupported", nd, N, Cc, sp, Oc, k, p)
                continue
            tx, tW, tb = okl.Tensor(x), okl.Tensor(W), okl.Tensor(b.reshape(-1))
            w2, wf = _DevBuf(ctx, W.size * 2), _DevBuf(ctx, W.size * 2)
            check(lib.okl_conv2_filter_banks(ctx.h, C.byref(d), tW._dptr(), w2.ptr, wf.ptr))
            Wr = bf16_round(W).astype(np.float64)
            yo = O.conv_forward_fast(x.astype(np.float64), Wr, b.astype(np.float64), 1, p, dtype=np.float64)
            yo_relu = np.maximum(yo, 0)
            oshape = yo.shape
            y = okl.Tensor._empty(oshape, np.float32, layout=_lib.NXC)
            y16 = _DevBuf(ctx, y.size * 2)
            check(lib.okl_conv2_fprop(ctx.h, C.byref(d), tx._b16(_lib.NXC), w2.ptr, tb._dptr(), y._ptr_raw(), y16.ptr))
            y._touch()
            e32 = rel(y.data, yo_relu)
            e16 = rel(fetch16(y16.ptr, (N,) + oshape[2:] + (Oc,)), yo_relu)
            # bf16-only output, no activation
            d0 = desc(nd, N, Cc, Oc, sp, k, p, act=0)
            check(lib.okl_memset(ctx.h, y16.ptr, 0, y.size * 2))
            check(lib.okl_conv2_fprop(ctx.h, C.byref(d0), tx._b16(_lib.NXC), w2.ptr, tb._dptr(), None, y16.ptr))
            e16b = rel(fetch16(y16.ptr, (N,) + oshape[2:] + (Oc,)), yo)
            # dgrad with the ReLU mask of the layer input
            g = bf16_round(rng.standard_normal(oshape).astype(np.float32))
            tg = okl.Tensor(g)
            _, _, dXo = O.conv_backward_fast(x.astype(np.float64), Wr, g.astype(np.float64), 1, p)
            dx = okl.Tensor._empty(x.shape, np.float32, layout=_lib.NXC)
            dx16 = _DevBuf(ctx, x.size * 2)
            check(lib.okl_conv2_dgrad(ctx.h, C.byref(d0), tg._b16(_lib.NXC), wf.ptr, dx._ptr_raw(), dx16.ptr, tx._b16(_lib.NXC)))
            dx._touch()
            ed = rel(dx.data, dXo * (x > 0))
            check(lib.okl_conv2_dgrad(ctx.h, C.byref(d0), tg._b16(_lib.NXC), wf.ptr, None, dx16.ptr, None))
            ed16 = rel(fetch16(dx16.ptr, (N,) + tuple(sp) + (Cc,)), dXo)
            ew = eb = -1.0
            if env == {"OKL_CT_SLAB": "1"} or env == {"OKL_CT_SLAB": "0"}:
                os.environ["OKL_WG_MT"] = "2" if env["OKL_CT_SLAB"] == "1" else "1"
                if lib.okl_conv2_wgrad_supported(ctx.h, C.byref(d0)):
                    dWo, dbo, _ = O.conv_backward_fast(x.astype(np.float64), Wr, g.astype(np.float64), 1, p)
                    dW = okl.Tensor._empty(W.shape, np.float32)
                    db = okl.Tensor._empty((Oc,), np.float32)
                    check(lib.okl_memset(ctx.h, dW._ptr_raw(), 0xff, W.size * 4))
                    check(lib.okl_conv2_wgrad(ctx.h, C.byref(d0), tx._b16(_lib.NXC), tg._b16(_lib.NXC), dW._ptr_raw(), db._ptr_raw()))
                    dW._touch(); db._touch()
                    ew, eb = rel(dW.data, dWo), rel(db.data, dbo.reshape(-1))
                else:
                    ew = eb = 0.0
                    print("   (wgrad unsupported for this shape)")
            ok = e32 < 2e-3 and e16 < 1e-2 and e16b < 1e-2 and ed < 2e-3 and ed16 < 1e-2 and ew < 2e-3 and eb < 2e-3
            bad += not ok
            print(("ok  " if ok else "BAD ") + f"{nd}d N={N} C={Cc} sp={sp} O={Oc} k={k} p={p} {env}: fprop f32 {e32:.2e} b16 {e16:.2e} b16only {e16b:.2e} "
                  f"dgrad f32 {ed:.2e} b16only {ed16:.2e} wgrad dW {ew:.2e} db {eb:.2e}", flush=True)
    for k_ in ("OKL_CT_BN", "OKL_CT_CLUSTER", "OKL_CT_PAIR", "OKL_CT_SLAB"):
        os.environ.pop(k_, None)
    print("PARITY", "FAILED %d" % bad if bad else "all ok")


LAYERS = [(64, 64, 32), (64, 128, 16), (128, 128, 16), (128, 256, 8), (256, 256, 8)]


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps * 1e3


def perf():
    N = int(os.environ.get("PROBE_BATCH", "1024"))
    rng = np.random.default_rng(0)
    for (Cc, Oc, S) in LAYERS:
        d = desc(2, N, Cc, Oc, (S, S), 3, 1, act=1)
        d0 = desc(2, N, Cc, Oc, (S, S), 3, 1, act=0)
        x16, y16, g16, dx16 = (_DevBuf(ctx, N * c * S * S * 2) for c in (Cc, Oc, Oc, Cc))
        y32, dx32 = _DevBuf(ctx, N * Oc * S * S * 4), _DevBuf(ctx, N * Cc * S * S * 4)
        for bfr in (x16, g16):
            check(lib.okl_memset(ctx.h, bfr.ptr, 0x3c, bfr.nbytes))
        W = okl.Tensor((rng.standard_normal((Oc, Cc, 3, 3)) * 0.1).astype(np.float32))
        bias = okl.Tensor(np.zeros(Oc, dtype=np.float32))
        w2, wf = _DevBuf(ctx, W.size * 2), _DevBuf(ctx, W.size * 2)
        check(lib.okl_conv2_filter_banks(ctx.h, C.byref(d), W._dptr(), w2.ptr, wf.ptr))
        fl = 2.0 * N * S * S * Oc * Cc * 9
        row = f"L {Cc:3d}->{Oc:3d}@{S:2d}: "
        for name, env in (("slab", {"OKL_CT_SLAB": "1"}), ("tap", {"OKL_CT_SLAB": "0"}), ("slab single", {"OKL_CT_SLAB": "1", "OKL_CT_PAIR": "0"}),
                          ("slab bn128", {"OKL_CT_SLAB": "1", "OKL_CT_BN": "128"}), ("slab bn128 single", {"OKL_CT_SLAB": "1", "OKL_CT_BN": "128", "OKL_CT_PAIR": "0"})):
            for k_ in ("OKL_CT_BN", "OKL_CT_CLUSTER", "OKL_CT_PAIR", "OKL_CT_SLAB"):
                os.environ.pop(k_, None)
            os.environ.update(env)
            t1 = timeit(lambda: check(lib.okl_conv2_fprop(ctx.h, C.byref(d), x16.ptr, w2.ptr, bias._dptr(), None, y16.ptr)))
            t2 = timeit(lambda: check(lib.okl_conv2_fprop(ctx.h, C.byref(d), x16.ptr, w2.ptr, bias._dptr(), y32.ptr, y16.ptr)))
            t3 = timeit(lambda: check(lib.okl_conv2_dgrad(ctx.h, C.byref(d0), g16.ptr, wf.ptr, None, dx16.ptr, x16.ptr)))
            t4 = timeit(lambda: check(lib.okl_conv2_dgrad(ctx.h, C.byref(d0), g16.ptr, wf.ptr, dx32.ptr, dx16.ptr, x16.ptr)))
            tw = -1.0
            os.environ["OKL_WG_ALL"] = "1"
            if name in ("slab", "tap") and lib.okl_conv2_wgrad_supported(ctx.h, C.byref(d0)):
                os.environ["OKL_WG_MT"] = "2" if name == "slab" else "1"
                dWb, dbb = _DevBuf(ctx, W.size * 4), _DevBuf(ctx, Oc * 4)
                tw = timeit(lambda: check(lib.okl_conv2_wgrad(ctx.h, C.byref(d0), x16.ptr, g16.ptr, dWb.ptr, dbb.ptr)))
                print(row + f"{name:10s} wgrad(MT={os.environ['OKL_WG_MT']}) {tw:7.1f}us {fl / tw / 1e6:6.0f} TF", flush=True)
            print(row + f"{name:10s} fprop b16 {t1:7.1f}us {fl / t1 / 1e6:6.0f} TF | +f32 {t2:7.1f}us {fl / t2 / 1e6:6.0f} TF | dgrad+mask b16 {t3:7.1f}us "
                  f"{fl / t3 / 1e6:6.0f} TF | +f32 {t4:7.1f}us {fl / t4 / 1e6:6.0f} TF", flush=True)
        os.environ.pop("OKL_CT_BN", None)
        os.environ["OKL_CT_SLAB"] = "1"


def timeline():
    lib.okl_debug_tc_timestamps.argtypes = [C.c_void_p, C.c_void_p]
    N, Cc, Oc, S = (int(v) for v in os.environ.get("TL_SHAPE", "1024,64,64,32").split(","))
    kk = tuple(int(v) for v in os.environ.get("TL_K", "3,3").split(","))
    d = desc(2, N, Cc, Oc, (S, S), kk, (kk[0] // 2, kk[1] // 2), act=1)
    x16, y16 = _DevBuf(ctx, N * Cc * S * S * 2), _DevBuf(ctx, N * Oc * S * S * 2)
    check(lib.okl_memset(ctx.h, x16.ptr, 0x3c, x16.nbytes))
    W = okl.Tensor((np.random.default_rng(0).standard_normal((Oc, Cc) + kk) * 0.1).astype(np.float32))
    bias = okl.Tensor(np.zeros(Oc, dtype=np.float32))
    w2 = _DevBuf(ctx, W.size * 2)
    check(lib.okl_conv2_filter_banks(ctx.h, C.byref(d), W._dptr(), w2.ptr, None))
    f = lambda: check(lib.okl_conv2_fprop(ctx.h, C.byref(d), x16.ptr, w2.ptr, bias._dptr(), None, y16.ptr))   # noqa: E731
    f(); ctx.sync()
    dbg = _DevBuf(ctx, (148 * 8 + 64) * 8)
    check(lib.okl_memset(ctx.h, dbg.ptr, 0, dbg.nbytes))
    lib.okl_debug_tc_timestamps(ctx.h, dbg.ptr)
    f(); ctx.sync()
    lib.okl_debug_tc_timestamps(ctx.h, None)
    raw = np.empty(148 * 8 + 64, dtype=np.uint64)
    check(lib.okl_d2h_async_raw(ctx.h, raw.ctypes.data, dbg.ptr, raw.nbytes)); ctx.sync()
    ts = raw[:148 * 8].reshape(148, 8).astype(np.int64)
    t0 = ts[:, 0][ts[:, 0] > 0].min()
    print("shape", N, Cc, Oc, S)
    print("cta0 start / first A landed / end (us):", [(int(v) - t0) / 1e3 for v in ts[0, [0, 1, 3]]], " last CTA end", (ts[:, 3].max() - t0) / 1e3)
    tt = raw[148 * 8:148 * 8 + 48].astype(np.int64).reshape(24, 2)
    print("per-tile (acc_ready, epi_done) us, CTA 0:", [(round((a - t0) / 1e3, 2), round((b - t0) / 1e3, 2)) for a, b in tt if a][:14])


def misc():
    """bf16 pooling / channel sums / bf16 -> fp32 against NumPy"""
    rng = np.random.default_rng(3)
    for (N, Cc, H, W) in ((4, 64, 32, 32), (3, 128, 16, 16), (5, 256, 8, 8), (2, 8, 6, 10)):
        x = bf16_round(rng.standard_normal((N, Cc, H, W)).astype(np.float32))
        x[0, :, 0:2, 0:2] = 0.5                                       # a tie: all four maxima receive the gradient
        g = bf16_round(rng.standard_normal((N, Cc, H // 2, W // 2)).astype(np.float32))
        tx, tg = okl.Tensor(x), okl.Tensor(g)
        y16, dx16 = _DevBuf(ctx, g.size * 2), _DevBuf(ctx, x.size * 2)
        db = okl.Tensor._empty((Cc,), np.float32)
        check(lib.okl_pool2_bf16_fwd(ctx.h, tx._b16(_lib.NXC), y16.ptr, N, Cc, H, W))
        yo = O.maxpool_forward(x.astype(np.float64), 2, 2)
        e1 = rel(fetch16(y16.ptr, (N, H // 2, W // 2, Cc)), yo)
        for relu in (0, 1):
            check(lib.okl_pool2_bf16_bwd(ctx.h, tx._b16(_lib.NXC), tg._b16(_lib.NXC), dx16.ptr, N, Cc, H, W, relu, db._ptr_raw(), 0))
            db._touch()
            gg = g.astype(np.float64) * ((yo > 0) if relu else 1.0)
            dxo = O.maxpool_backward(x.astype(np.float64), gg, 2, 2)
            e2 = rel(fetch16(dx16.ptr, (N, H, W, Cc)), dxo)
            e3 = rel(db.data, dxo.sum(axis=(0, 2, 3)))
            print(("ok  " if max(e1, e2) < 1e-6 and e3 < 1e-5 else "BAD ") + f"pool2 bf16 N={N} C={Cc} {H}x{W} relu={relu}: fwd {e1:.1e} bwd {e2:.1e} db {e3:.1e}")
        cs = okl.Tensor._empty((Cc,), np.float32)
        check(lib.okl_channel_sum_bf16(ctx.h, tx._b16(_lib.NXC), cs._ptr_raw(), N * H * W, Cc))
        cs._touch()
        e4 = rel(cs.data, x.astype(np.float64).sum(axis=(0, 2, 3)))
        f32 = okl.Tensor._empty((x.size,), np.float32)
        check(lib.okl_from_bf16(ctx.h, tx._b16(_lib.NXC), f32._ptr_raw(), x.size))
        f32._touch()
        e5 = rel(f32.data.reshape(N, H, W, Cc), np.moveaxis(x, 1, -1))
        print(("ok  " if e4 < 1e-5 and e5 == 0 else "BAD ") + f"channel_sum {e4:.1e} from_bf16 {e5:.1e}")


def first():
    cases = [((8, 3, 32, 32), 64, 3, 1), ((3, 3, 30, 27), 128, 3, 1), ((2, 1, 28, 28), 64, 5, 0), ((4, 3, 16, 16), 256, 3, 1),
             ((2, 3, 4, 10, 12), 64, 3, 1), ((1, 3, 5, 40, 36), 64, 3, 1), ((2, 2, 6, 9, 16), 128, (3, 3, 3), (0, 1, 1)), ((1, 3, 3, 112, 112), 64, 3, 1),
             ((2, 4, 20, 24), 64, 5, 2)]
    for (xs, Oc, k, pd) in cases:
        nd = len(xs) - 2
        N, Cc = xs[0], xs[1]
        rng = np.random.default_rng(N + Oc + nd)
        kt, pt = O._tup(k, nd), O._tup(pd, nd)
        x = rng.standard_normal(xs).astype(np.float32)
        Wt = (rng.standard_normal((Oc, Cc) + kt) * 0.2).astype(np.float32)
        b = (rng.standard_normal((Oc, 1)) * 0.1).astype(np.float32)
        d = _lib.ConvDesc()
        d.nd, d.N, d.C, d.O = nd, N, Cc, Oc
        for i in range(nd):
            d.inp[i], d.k[i], d.s[i], d.p[i] = xs[2 + i], kt[i], 1, pt[i]
        d.act, d.x_layout, d.y_layout = 1, _lib.NCX, _lib.NXC
        if not lib.okl_conv_first_supported(ctx.h, C.byref(d)):
            print("first-layer kernel unsupported for", xs, Oc, k, pd)
            continue
        tx, tW, tb = okl.Tensor(x), okl.Tensor(Wt), okl.Tensor(b.reshape(-1))
        xr, Wr, br = bf16_round(x).astype(np.float64), bf16_round(Wt).astype(np.float64), bf16_round(b).astype(np.float64)
        yo = np.maximum(O.conv_forward_fast(xr, Wr, br, 1, pd, dtype=np.float64), 0)
        y16 = _DevBuf(ctx, yo.size * 2)
        check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), tx._dptr(), None, tW._dptr(), tb._dptr(), y16.ptr))
        e1 = rel(fetch16(y16.ptr, (N,) + yo.shape[2:] + (Oc,)), yo)
        g = bf16_round(rng.standard_normal(yo.shape).astype(np.float32))
        tg = okl.Tensor(g)
        dWo, dbo, _ = O.conv_backward_fast(xr, Wr, g.astype(np.float64), 1, pd)
        dW, db = okl.Tensor._empty(Wt.shape, np.float32), okl.Tensor._empty((Oc,), np.float32)
        d.act = 0
        check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), tx._dptr(), None, tg._b16(_lib.NXC), dW._ptr_raw(), db._ptr_raw()))
        dW._touch(); db._touch()
        e2, e3 = rel(dW.data, dWo), rel(db.data, dbo.reshape(-1))
        print(("ok  " if e1 < 1e-2 and e2 < 2e-3 and e3 < 2e-3 else "BAD ") + f"first layer x={xs} O={Oc} k={k} p={pd}: fprop {e1:.2e} dW {e2:.2e} db {e3:.2e}", flush=True)
    # timing at the bench shape
    N = int(os.environ.get("PROBE_BATCH", "1024"))
    d = _lib.ConvDesc()
    d.nd, d.N, d.C, d.O = 2, N, 3, 64
    for i in range(2):
        d.inp[i], d.k[i], d.s[i], d.p[i] = 32, 3, 1, 1
    d.act, d.x_layout, d.y_layout = 1, _lib.NCX, _lib.NXC
    x = okl.Tensor(np.random.default_rng(0).standard_normal((N, 3, 32, 32), dtype=np.float32))
    Wt, bb = okl.Tensor(np.ones((64, 3, 3, 3), dtype=np.float32)), okl.Tensor(np.zeros(64, dtype=np.float32))
    y16, g16 = _DevBuf(ctx, N * 64 * 1024 * 2), _DevBuf(ctx, N * 64 * 1024 * 2)
    check(lib.okl_memset(ctx.h, g16.ptr, 0x3c, g16.nbytes))
    dW, db = _DevBuf(ctx, 64 * 27 * 4), _DevBuf(ctx, 256)
    t1 = timeit(lambda: check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), x._dptr(), None, Wt._dptr(), bb._dptr(), y16.ptr)))
    t2 = timeit(lambda: check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), x._dptr(), None, g16.ptr, dW.ptr, db.ptr)))
    print(f"first layer 3->64@32 batch {N}: fprop {t1:.1f} us ({N * 65536 * 2 / t1 / 1e3:.0f} GB/s written), wgrad+db {t2:.1f} us ({N * 65536 * 2 / t2 / 1e3:.0f} GB/s read)")
    # the video layer of BASELINE configs[3]: Conv3D(3, 64, 3, 1, 1) on 8 x 3 x 16 x 112 x 112
    d = _lib.ConvDesc()
    d.nd, d.N, d.C, d.O = 3, 8, 3, 64
    for i, e in enumerate((16, 112, 112)):
        d.inp[i], d.k[i], d.s[i], d.p[i] = e, 3, 1, 1
    d.act, d.x_layout, d.y_layout = 1, _lib.NCX, _lib.NXC
    x = okl.Tensor(np.random.default_rng(0).standard_normal((8, 3, 16, 112, 112), dtype=np.float32))
    Wt, bb = okl.Tensor(np.ones((64, 3, 3, 3, 3), dtype=np.float32)), okl.Tensor(np.zeros(64, dtype=np.float32))
    nout = 8 * 16 * 112 * 112 * 64
    y16, g16 = _DevBuf(ctx, nout * 2), _DevBuf(ctx, nout * 2)
    check(lib.okl_memset(ctx.h, g16.ptr, 0x3c, g16.nbytes))
    dW, db = _DevBuf(ctx, 64 * 81 * 4), _DevBuf(ctx, 256)
    t1 = timeit(lambda: check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), x._dptr(), None, Wt._dptr(), bb._dptr(), y16.ptr)))
    t2 = timeit(lambda: check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), x._dptr(), None, g16.ptr, dW.ptr, db.ptr)))
    print(f"video layer 3->64 k3 @16x112x112 batch 8: fprop {t1:.1f} us ({nout * 2 / t1 / 1e3:.0f} GB/s written), wgrad+db {t2:.1f} us ({nout * 2 / t2 / 1e3:.0f} GB/s read)")


if "first" in what:
    first()
def diag():
    """Where does the k loop's time go?  OKL_CT_DIAG / OKL_WG_DIAG: 1 = every MMA issued twice (time doubles only if the MMA / operand
    fetch is the limit), 2 / 4 = A / B operands loaded once per stage slot (time drops only if L2 -> shared memory is the limit)"""
    N = int(os.environ.get("PROBE_BATCH", "1024"))
    rng = np.random.default_rng(0)
    for (Cc, Oc, S) in LAYERS:
        d = desc(2, N, Cc, Oc, (S, S), 3, 1, act=1)
        d0 = desc(2, N, Cc, Oc, (S, S), 3, 1, act=0)
        x16, y16, g16, dx16 = (_DevBuf(ctx, N * c * S * S * 2) for c in (Cc, Oc, Oc, Cc))
        for bfr in (x16, g16):
            check(lib.okl_memset(ctx.h, bfr.ptr, 0x3c, bfr.nbytes))
        W = okl.Tensor((rng.standard_normal((Oc, Cc, 3, 3)) * 0.1).astype(np.float32))
        bias = okl.Tensor(np.zeros(Oc, dtype=np.float32))
        w2, wf = _DevBuf(ctx, W.size * 2), _DevBuf(ctx, W.size * 2)
        check(lib.okl_conv2_filter_banks(ctx.h, C.byref(d), W._dptr(), w2.ptr, wf.ptr))
        dWb, dbb = _DevBuf(ctx, W.size * 4), _DevBuf(ctx, Oc * 4)
        os.environ["OKL_WG_ALL"] = "1"
        row = f"L {Cc:3d}->{Oc:3d}@{S:2d}: "
        for pair in ("1", "0"):
            os.environ["OKL_CT_PAIR"] = pair
            out = []
            for dg in ((0, 1) if pair == "1" else (0, 1, 2, 4, 6, 7)):
                os.environ["OKL_CT_DIAG"] = str(dg)
                t1 = timeit(lambda: check(lib.okl_conv2_fprop(ctx.h, C.byref(d), x16.ptr, w2.ptr, bias._dptr(), None, y16.ptr)))
                out.append(f"diag{dg} {t1:6.1f}")
            os.environ["OKL_CT_DIAG"] = "0"
            print(row + f"fprop pair={pair}  " + "  ".join(out), flush=True)
        os.environ.pop("OKL_CT_PAIR", None)
        out = []
        for dg in (0, 1, 2, 3):
            os.environ["OKL_WG_DIAG"] = str(dg)
            tw = timeit(lambda: check(lib.okl_conv2_wgrad(ctx.h, C.byref(d0), x16.ptr, g16.ptr, dWb.ptr, dbb.ptr)))
            out.append(f"diag{dg} {tw:6.1f}")
        os.environ["OKL_WG_DIAG"] = "0"
        print(row + "wgrad (gen 2)  " + "  ".join(out), flush=True)


WG_PAR = PAR + [(2, 3, 128, (12, 12), 64, 3, 1), (2, 2, 192, (8, 8), 128, 3, 1), (2, 5, 64, (10, 14), 64, (2, 3), (0, 1)), (2, 3, 64, (16, 16), 192, (4, 3), (1, 1)),
                (2, 2, 64, (16, 16), 64, 3, 2), (2, 7, 256, (8, 8), 64, 3, 1), (2, 2, 64, (33, 30), 128, 3, 1), (2, 17, 128, (4, 8), 128, 3, 1)]


def wgrad():
    """parity of okl_conv2_wgrad (gradient-slab kernel at both k-block sizes, the first slab kernel) + timing at the CIFAR layer shapes"""
    bad = 0
    for (nd, N, Cc, sp, Oc, k, p) in ([] if os.environ.get('PROBE_SKIP_PARITY') else WG_PAR):
        rng = np.random.default_rng(N * 100 + Cc + Oc)
        kt = O._tup(k, nd)
        x = bf16_round(rng.standard_normal((N, Cc) + sp).astype(np.float32))
        Wr = np.zeros((Oc, Cc) + kt)
        d0 = desc(nd, N, Cc, Oc, sp, k, p, act=0)
        if not lib.okl_conv2_supported(ctx.h, C.byref(d0)):
            continue
        oshape = O.conv_forward_fast(x[:1].astype(np.float64), Wr, np.zeros((Oc, 1)), 1, p, dtype=np.float64).shape
        g = bf16_round(rng.standard_normal((N,) + oshape[1:]).astype(np.float32))
        dWo, dbo, _ = O.conv_backward_fast(x.astype(np.float64), Wr, g.astype(np.float64), 1, p)
        tx, tg = okl.Tensor(x), okl.Tensor(g)
        for env in ({}, {"OKL_WG_ALL": "1"}, {"OKL_WG_ALL": "1", "OKL_WG_PT": "64"}, {"OKL_WG_ALL": "1", "OKL_WG_PT": "128"}, {"OKL_WG_GS": "0", "OKL_WG_ALL": "1"}):
            for k_ in ("OKL_WG_PT", "OKL_WG_GS", "OKL_WG_ALL"):
                os.environ.pop(k_, None)
            os.environ.update(env)
            if not lib.okl_conv2_wgrad_supported(ctx.h, C.byref(d0)):
                print(f"    {nd}d N={N} C={Cc} sp={sp} O={Oc} k={k} p={p} {env}: unsupported")
                continue
            for want_db in (True, False):
                dW = okl.Tensor._empty(Wr.shape, np.float32)
                db = okl.Tensor._empty((Oc,), np.float32)
                check(lib.okl_memset(ctx.h, dW._ptr_raw(), 0xff, dW.size * 4))
                check(lib.okl_memset(ctx.h, db._ptr_raw(), 0xff, Oc * 4))
                check(lib.okl_conv2_wgrad(ctx.h, C.byref(d0), tx._b16(_lib.NXC), tg._b16(_lib.NXC), dW._ptr_raw(), db._ptr_raw() if want_db else None))
                dW._touch(); db._touch()
                ew = rel(dW.data, dWo)
                eb = rel(db.data, dbo.reshape(-1)) if want_db else 0.0
                ok = ew < 2e-3 and eb < 2e-3
                bad += not ok
                print(("ok  " if ok else "BAD ") + f"wgrad {nd}d N={N} C={Cc} sp={sp} O={Oc} k={k} p={p} {env} db={want_db}: dW {ew:.2e} db {eb:.2e}", flush=True)
    for k_ in ("OKL_WG_PT", "OKL_WG_GS", "OKL_WG_ALL"):
        os.environ.pop(k_, None)
    print("WGRAD PARITY", "FAILED %d" % bad if bad else "all ok")
    N = int(os.environ.get("PROBE_BATCH", "1024"))
    for (Cc, Oc, S) in LAYERS:
        d0 = desc(2, N, Cc, Oc, (S, S), 3, 1, act=0)
        x16, g16 = _DevBuf(ctx, N * Cc * S * S * 2), _DevBuf(ctx, N * Oc * S * S * 2)
        for bfr in (x16, g16):
            check(lib.okl_memset(ctx.h, bfr.ptr, 0x3c, bfr.nbytes))
        dWb, dbb = _DevBuf(ctx, Oc * Cc * 9 * 4), _DevBuf(ctx, Oc * 4)
        fl = 2.0 * N * S * S * Oc * Cc * 9
        out = []
        for name, env in (("default", {}), ("gs auto", {"OKL_WG_ALL": "1"}), ("gs pt64", {"OKL_WG_ALL": "1", "OKL_WG_PT": "64"}), ("gs pt128", {"OKL_WG_ALL": "1", "OKL_WG_PT": "128"}),
                          ("slab-x (old)", {"OKL_WG_GS": "0", "OKL_WG_ALL": "1"})):
            for k_ in ("OKL_WG_PT", "OKL_WG_GS", "OKL_WG_ALL"):
                os.environ.pop(k_, None)
            os.environ.update(env)
            if not lib.okl_conv2_wgrad_supported(ctx.h, C.byref(d0)):
                continue
            tw = timeit(lambda: check(lib.okl_conv2_wgrad(ctx.h, C.byref(d0), x16.ptr, g16.ptr, dWb.ptr, dbb.ptr)))
            out.append(f"{name} {tw:6.1f} us {fl / tw / 1e6:5.0f} TF")
        for k_ in ("OKL_WG_PT", "OKL_WG_GS", "OKL_WG_ALL"):
            os.environ.pop(k_, None)
        print(f"L {Cc:3d}->{Oc:3d}@{S:2d} wgrad+db: " + " | ".join(out), flush=True)
        out = []
        os.environ["OKL_WG_ALL"] = "1"
        for dg in (0, 4, 132):        # 1: MMAs issued twice, 2: operands loaded once per stage slot, 4: no loads / waits, 8: one accumulator, 32: N = 64, 64: A LBO = 0
            os.environ["OKL_WG_DIAG"] = str(dg)
            tw = timeit(lambda: check(lib.okl_conv2_wgrad(ctx.h, C.byref(d0), x16.ptr, g16.ptr, dWb.ptr, None)))
            out.append(f"diag{dg} {tw:6.1f}")
        os.environ.pop("OKL_WG_DIAG", None)
        os.environ.pop("OKL_WG_ALL", None)
        print(f"L {Cc:3d}->{Oc:3d}@{S:2d} wgrad, no db: " + "  ".join(out), flush=True)


if "wgrad" in what:
    wgrad()

def vexp():
    """timing of the expanded-taps first layer (CIFAR shape) piece by piece against the small-K kernels"""
    N = int(os.environ.get("PROBE_BATCH", "1024"))
    d = _lib.ConvDesc()
    d.nd, d.N, d.C, d.O = 2, N, 3, 64
    for i in range(2):
        d.inp[i], d.k[i], d.s[i], d.p[i] = 32, 3, 1, 1
    d.act, d.x_layout, d.y_layout = 1, _lib.NCX, _lib.NXC
    de = _lib.ConvDesc()
    check(lib.okl_conv_vexp_desc(ctx.h, C.byref(d), C.byref(de)))
    x = okl.Tensor(np.random.default_rng(0).standard_normal((N, 3, 32, 32), dtype=np.float32))
    Wt, bb = okl.Tensor(np.ones((64, 3, 3, 3), dtype=np.float32)), okl.Tensor(np.zeros(64, dtype=np.float32))
    xe, w2e = _DevBuf(ctx, N * 32 * 32 * 32), _DevBuf(ctx, 64 * 3 * 32)
    y16, g16 = _DevBuf(ctx, N * 64 * 1024 * 2), _DevBuf(ctx, N * 64 * 1024 * 2)
    check(lib.okl_memset(ctx.h, g16.ptr, 0x3c, g16.nbytes))
    dwe, dW, db = _DevBuf(ctx, 64 * 16 * 3 * 4), _DevBuf(ctx, 64 * 27 * 4), _DevBuf(ctx, 256)
    t0 = timeit(lambda: check(lib.okl_conv_vexp_expand(ctx.h, C.byref(d), x._dptr(), None, xe.ptr)))
    tb = timeit(lambda: check(lib.okl_conv_vexp_bank(ctx.h, C.byref(d), Wt._dptr(), w2e.ptr)))
    for env in ({}, {"OKL_CT_PAIR": "0"}, {"OKL_CT_SLAB": "0"}):
        for k_ in ("OKL_CT_PAIR", "OKL_CT_SLAB"):
            os.environ.pop(k_, None)
        os.environ.update(env)
        t1 = timeit(lambda: check(lib.okl_conv2_fprop(ctx.h, C.byref(de), xe.ptr, w2e.ptr, bb._dptr(), None, y16.ptr)))
        print(f"vexp fprop {env}: {t1:.1f} us")
    for k_ in ("OKL_CT_PAIR", "OKL_CT_SLAB"):
        os.environ.pop(k_, None)
    de.act = 0
    for env in ({}, {"OKL_WG_PT": "64"}, {"OKL_WG_PT": "128"}):
        os.environ.pop("OKL_WG_PT", None)
        os.environ.update(env)
        t2 = timeit(lambda: check(lib.okl_conv2_wgrad(ctx.h, C.byref(de), xe.ptr, g16.ptr, dwe.ptr, db.ptr)))
        print(f"vexp wgrad {env}: {t2:.1f} us")
    os.environ.pop("OKL_WG_PT", None)
    t3 = timeit(lambda: check(lib.okl_conv_vexp_dw(ctx.h, C.byref(d), dwe.ptr, dW.ptr)))
    f1 = timeit(lambda: check(lib.okl_conv_first_fprop(ctx.h, C.byref(d), x._dptr(), None, Wt._dptr(), bb._dptr(), y16.ptr)))
    f2 = timeit(lambda: check(lib.okl_conv_first_wgrad(ctx.h, C.byref(d), x._dptr(), None, g16.ptr, dW.ptr, db.ptr)))
    print(f"vexp: expand {t0:.1f} us, bank {tb:.1f} us, dW remap {t3:.1f} us | small-K kernels: fprop {f1:.1f} us, wgrad+db {f2:.1f} us")


if "vexp" in what:
    vexp()

if "diag" in what:
    diag()
if "parity" in what:
    parity()
if "misc" in what:
    misc()
if "perf" in what:
    perf()
if "timeline" in what:
    timeline()
