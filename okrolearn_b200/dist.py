"""Data-parallel bring-up for one node, one process per GPU (launched by torchrun / torch.distributed.run).

Rank 0 creates the NCCL unique id; every rank creates its symmetric peer-memory buffer; ids and cudaIpc handles are exchanged
through small files under /tmp keyed by MASTER_PORT / run id / launcher pid (single node, so a shared /tmp is enough and neither
torch nor a TCP store is needed on this path)."""
from __future__ import annotations

import ctypes as C
import os
import time

from . import _lib
from ._lib import lib, check, get_ctx


def _tag():
    return f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'x')}_{os.getppid()}"


def _publish(path, blob):
    with open(path + ".tmp", "wb") as f:
        f.write(blob)
    os.replace(path + ".tmp", path)


def _await(path, nbytes, timeout=120.0):
    t0 = time.time()
    while True:
        if os.path.exists(path):
            with open(path, "rb") as f:
                b = f.read()
            if len(b) == nbytes:
                return b
        if time.time() - t0 > timeout:
            raise RuntimeError(f"timed out waiting for {path}")
        time.sleep(0.005)


def exchange_unique_id(rank, world):
    path = f"/tmp/okl_nccl_id_{_tag()}"
    if rank == 0:
        uid = _lib.Context.comm_unique_id()
        _publish(path, uid)
        return uid
    return _await(path, 128)


def init_data_parallel(p2p_bytes=64 << 20):
    """Read RANK / WORLD_SIZE / LOCAL_RANK, create the context on the local GPU, join the NCCL communicator and map every peer's
    symmetric buffer.  Returns (rank, world).  No-op for a single process."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("OKL_DEVICE", os.environ.get("LOCAL_RANK", "0"))
    ctx = get_ctx()
    if world <= 1 or ctx.nranks == world:
        return rank, world
    ctx.comm_init(exchange_unique_id(rank, world), rank, world)
    if os.environ.get("OKL_P2P", "1") != "0" and world <= 8:
        try:
            check(lib.okl_p2p_alloc(ctx.h, p2p_bytes))
            h = C.create_string_buffer(64)
            check(lib.okl_p2p_export(ctx.h, h))
            base = f"/tmp/okl_p2p_{_tag()}_"
            _publish(base + str(rank), h.raw)
            for r in range(world):
                if r != rank:
                    peer = C.create_string_buffer(_await(base + str(r), 64), 64)
                    check(lib.okl_p2p_import(ctx.h, r, peer))
            check(lib.okl_barrier(ctx.h))          # nobody touches a peer before everyone has mapped everyone
            ctx.p2p_ready = True
        except (RuntimeError, ValueError, MemoryError) as e:      # no peer access on this box: NCCL handles every bucket
            ctx.p2p_ready = False
            print(f"[okl] peer-memory all-reduce unavailable ({e}); using NCCL for all buckets")
    return rank, world
