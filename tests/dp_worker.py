"""Worker for the data-parallel tests.  Launched by torch.distributed.run (env RANK / WORLD_SIZE / MASTER_*).

  mode "gloo": CPU only - the DP *scheme* (contiguous batch shards, loss gradient divided by the GLOBAL batch, sum-allreduce of
               each layer's [dW | db], replicated accumulated-gradient update) carried out with the oracle's math and a gloo
               allreduce must reproduce the single-process oracle on the full batch.
  mode "nccl": the product: one process per GPU, libokl_b200's NCCL exchange overlapped with backward, CUDA-graph replay;
               compared on rank 0 with the single-process oracle on the full batch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import okl_oracle as O  # noqa: E402

mode, out_path = sys.argv[1], sys.argv[2]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
rng = np.random.default_rng(77)
B, STEPS, LR = 16 * world, 3, 0.05                   # global batch
X = rng.standard_normal((STEPS, B, 1, 12, 12)).astype(np.float32)
T = rng.integers(0, 10, (STEPS, B))
Wc, W1, W2 = rng.standard_normal((4, 1, 3, 3)) * 0.3, rng.standard_normal((100, 16)) * 0.1, rng.standard_normal((16, 10)) * 0.1


def oracle_net():
    return O.OracleNetwork([O.OracleConv(Wc, np.zeros((4, 1)), fast=True), O.OracleReLU(), O.OracleMaxPool(2, 2), O.OracleFlatten(),
                            O.OracleDense(W1, np.zeros((1, 16))), O.OracleReLU(), O.OracleDense(W2, np.zeros((1, 10))),
                            O.OracleSoftmax()])


full = oracle_net()
ref_losses = [full.train_step(X[s].astype(np.float64), T[s], O.OracleCrossEntropy(), LR) for s in range(STEPS)]
sl = slice(rank * B // world, (rank + 1) * B // world)
res = {}

if mode == "gloo":
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = oracle_net()
    plist = [net.layers[i] for i in (0, 4, 6)]
    state = [[l.W.copy(), l.b.copy(), np.zeros_like(l.W), np.zeros_like(l.b)] for l in plist]      # W, b, gaccW, gaccb
    losses = []
    for s in range(STEPS):
        for l, st in zip(plist, state):
            l.W, l.b, l.gW, l.gb = st[0].copy(), st[1].copy(), None, None
        out = net.forward(X[s, sl].astype(np.float64))
        ce = O.OracleCrossEntropy()
        local_loss = ce.forward(out, T[s, sl])
        net.backward(ce.backward(out, T[s, sl], denom=B), 0.0)                 # lr 0: only collect this shard's dW / db
        lt = torch.tensor([local_loss / world])
        dist.all_reduce(lt)
        losses.append(float(lt))
        for l, st in zip(plist, state):
            bucket = torch.tensor(np.concatenate([l.dW.reshape(-1), l.db.reshape(-1)]))
            dist.all_reduce(bucket)                                             # the one exchange step of the path
            g = bucket.numpy()
            dW, db = g[:l.dW.size].reshape(l.dW.shape), g[l.dW.size:].reshape(st[1].shape)
            st[2] += dW
            st[3] += db
            st[0] -= LR * st[2]
            st[1] -= LR * st[3]
    res = {"loss_err": float(np.max(np.abs(np.array(losses) - np.array(ref_losses)))),
           "w_err": max(float(np.max(np.abs(st[0] - fl.W)) / np.max(np.abs(fl.W))) for st, fl in
                        zip(state, [full.layers[i] for i in (0, 4, 6)]))}
    dist.destroy_process_group()
else:
    os.environ.setdefault("OKL_DEVICE", os.environ.get("LOCAL_RANK", "0"))
    import logging
    import okrolearn_b200.okrolearn as okl
    from okrolearn_b200 import _lib
    from okrolearn_b200 import dist
    dist.init_data_parallel()
    ctx = okl.get_ctx()
    net = okl.NeuralNetwork()
    logging.getLogger("NeuralNetwork").setLevel(logging.CRITICAL)
    conv, d1, d2 = okl.Conv2DLayer(1, 4, 3), okl.DenseLayer(100, 16), okl.DenseLayer(16, 10)
    conv.filters.data, d1.weights.data, d2.weights.data = Wc.copy(), W1.copy(), W2.copy()
    for l in (conv, okl.ReLUActivationLayer(), okl.MaxPoolingLayer(2, 2), okl.FlattenLayer(), d1, okl.ReLUActivationLayer(), d2,
              okl.SoftmaxActivationLayer()):
        net.add(l)
    # train() shuffles with the host RNG: make the permutation the identity so every rank sees its shard in order
    np.random.shuffle = lambda a: None
    Xs, Ts = X[:, sl].reshape(-1, 1, 12, 12), T[:, sl].reshape(-1)
    losses = net.train(okl.Tensor(Xs), okl.Tensor(Ts), 1, okl.LearningRateScheduler(LR), okl.SGDOptimizer(), B // world,
                       okl.CrossEntropyLoss())
    errs = [float(np.max(np.abs(a - b)) / np.max(np.abs(b))) for a, b in ((conv.filters.data, full.layers[0].W),
                                                                           (d1.weights.data, full.layers[4].W),
                                                                           (d2.weights.data, full.layers[6].W),
                                                                           (d1.biases.data, full.layers[4].b))]
    res = {"w_err": max(errs), "local_mean_loss": float(losses[0]), "ref_mean_loss": float(np.mean(ref_losses)),
           "p2p": bool(ctx.p2p_ready)}
    ctx.sync()

if rank == 0:
    with open(out_path, "w") as f:
        json.dump(res, f)
