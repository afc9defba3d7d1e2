"""Host-side profile of the end-to-end train loop (stream_inputs + sync_loss_every_step): where the Python/driver time goes."""
import cProfile
import logging
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from okrolearn_b200 import okrolearn as okl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mnist_cnn"
okl.set_precision("tf32")
ctx = okl.get_ctx()
w = bench.WORKLOADS[name]
B, K = w["batch"], 1000
np.random.seed(1234)
net = bench.build_net(okl, name)
logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)
loss_fn = okl.CrossEntropyLoss() if w["loss"] == "ce" else okl.MSELoss()
sched, opt = okl.LearningRateScheduler(0.01), okl.SGDOptimizer(lr=0.01)
X, T = bench.synth(name, B * K, 0)
pX = okl.pinned_empty(X.shape, np.float32)
pX[...] = X
net.stream_inputs, net.sync_loss_every_step = True, True
for mode in ("lagged loss", "no loss read"):
    net.sync_loss_every_step = mode == "lagged loss"
    hX, hT = okl.Tensor.wrap(pX), okl.Tensor(T)
    net.train(okl.Tensor.wrap(pX[:B * 8]), okl.Tensor(T[:B * 8]), 1, sched, opt, B, loss_fn)
    ctx.sync()
    t0 = time.perf_counter()
    net.train(hX, hT, 1, sched, opt, B, loss_fn)
    ctx.sync()
    print(f"{mode}: {(time.perf_counter() - t0) / K * 1e6:.1f} us/step")
net.sync_loss_every_step = True
hX, hT = okl.Tensor.wrap(pX), okl.Tensor(T)
pr = cProfile.Profile()
pr.enable()
net.train(hX, hT, 1, sched, opt, B, loss_fn)
ctx.sync()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
