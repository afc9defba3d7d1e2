"""Per-link cost of the train step's kernel chain (debug timeline: a %globaltimer stamp after every kernel, also inside the
captured graph).  python scripts/step_timeline.py [workload] [precision]"""
import ctypes as C
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from okrolearn_b200 import okrolearn as okl  # noqa: E402
from okrolearn_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mnist_cnn"
okl.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32")
from okrolearn_b200 import dist  # noqa: E402
rank, world = dist.init_data_parallel()
ctx = okl.get_ctx()
raw = C.CDLL(_lib.lib._name)
w = bench.WORKLOADS[name]
B = int(os.environ.get("TL_BATCH", w["batch"]))
np.random.seed(1234)
net = bench.build_net(okl, name)
logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)
loss_fn = okl.CrossEntropyLoss() if w["loss"] == "ce" else okl.MSELoss()
lr = 1e-9 if name == "mlp8192" else 0.01
sched, opt = okl.LearningRateScheduler(lr), okl.SGDOptimizer(lr=lr)
X, T = bench.synth(name, B * 6, rank)
tX, tT = okl.Tensor(X), okl.Tensor(T)
tX._dptr()
net.train(tX, tT, 1, sched, opt, B, loss_fn)            # allocations, workspace growth
net._graphs.clear()                                      # re-capture with the stamps inside the graph
ctx.sync()
raw.okl_debug_timeline_begin.argtypes = [C.c_void_p]
raw.okl_debug_timeline_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
assert raw.okl_debug_timeline_begin(ctx.h) == 0
net.train(tX, tT, 1, sched, opt, B, loss_fn)
out = (C.c_ulonglong * 4096)()
names = C.create_string_buffer(1 << 18)
n = C.c_int()
assert raw.okl_debug_timeline_read(ctx.h, out, 4096, names, 1 << 18, C.byref(n)) == 0
if rank != 0:
    sys.exit(0)
labels = names.value.decode().split("\n")[:n.value]
ts = [out[i] for i in range(n.value)]
cal = ts[1] - ts[0]
print(f"{name}: {n.value} stamps; back-to-back stamp distance {cal} ns (subtract from every link)")
# the captured kernels carry the times of the LAST replay; the eager gather of that step is the last eager label
gi = [i for i, l in enumerate(labels) if l.endswith("g")]
last_eager = max(i for i, l in enumerate(labels) if not l.endswith("g"))
chain = [last_eager - 1, last_eager] + gi if last_eager > gi[-1] else gi
prev = None
t_first = None
for i in chain:
    l, t = labels[i], ts[i]
    lane = l.split("@")[1][0]
    if t_first is None:
        t_first = t
    d = "" if prev is None else f"{t - prev:8d}"
    print(f"  {l:34s} lane {lane}  t={t - t_first:8d} ns   since previous stamp {d}")
    prev = t
