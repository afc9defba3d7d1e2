"""Where does the end-to-end (stream_inputs + sync_loss_every_step) train() call of the cifar6 bench spend its wall time?"""
import cProfile, logging, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from okrolearn_b200 import okrolearn as okl  # noqa: E402

name = "cifar6"
okl.set_precision("bf16")
ctx = okl.get_ctx()
w = bench.WORKLOADS[name]
B = w["batch"]
np.random.seed(1234)
net = bench.build_net(okl, name)
logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)
loss_fn = okl.CrossEntropyLoss()
sched, opt = okl.LearningRateScheduler(0.01), okl.SGDOptimizer(lr=0.01)
for K in (20, 100):
    X, T = bench.synth(name, B * K, 0)
    pX = okl.pinned_empty(X.shape, np.float32); pX[...] = X
    pT = okl.pinned_empty(T.shape, np.int32); pT[...] = T
    for stream, sync_loss in ((False, False), (True, False), (True, True)):
        net.stream_inputs, net.sync_loss_every_step = stream, sync_loss
        mk = (lambda: (okl.Tensor.wrap(pX), okl.Tensor.wrap(pT))) if stream else (lambda: (okl.Tensor(X), okl.Tensor(T)))
        hX, hT = mk()
        net.train(hX, hT, 1, sched, opt, B, loss_fn)
        ctx.sync()
        ts = []
        for rep in range(5):
            hX, hT = mk() if not stream else (hX, hT)
            if not stream:
                net.prepare(hX, hT, loss_fn)
            ctx.sync()
            t0 = time.perf_counter()
            net.train(hX, hT, 1, sched, opt, B, loss_fn)
            t1 = time.perf_counter()
            ctx.sync()
            t2 = time.perf_counter()
            ts.append((t2 - t0, t1 - t0))
        tot, host = np.median([a for a, _ in ts]), np.median([b for _, b in ts])
        print(f"K={K} stream={stream} sync_loss={sync_loss}: {tot * 1e3:.2f} ms per call = {tot / K * 1e3:.3f} ms/step (host returned after {host * 1e3:.2f} ms)", flush=True)
net.stream_inputs, net.sync_loss_every_step = True, True
K = 20
X, T = bench.synth(name, B * K, 0)
pX = okl.pinned_empty(X.shape, np.float32); pX[...] = X
pT = okl.pinned_empty(T.shape, np.int32); pT[...] = T
hX, hT = okl.Tensor.wrap(pX), okl.Tensor.wrap(pT)
net.train(hX, hT, 1, sched, opt, B, loss_fn); ctx.sync()
pr = cProfile.Profile(); pr.enable()
net.train(hX, hT, 1, sched, opt, B, loss_fn); ctx.sync()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
