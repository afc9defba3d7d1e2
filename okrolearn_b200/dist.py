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


def _setup_multicast(ctx, rank, world, nbytes):
    """NVLS bring-up: one multicast object for the node, every rank's gradient region bound to it (okl_mc_* in csrc/okl_p2p.cu).
    Rank 0 creates the object and hands its POSIX file descriptor to the other ranks over a Unix-domain socket (SCM_RIGHTS).
    Every step is followed by an all-rank agreement, so a rank that cannot do it (no NVSwitch multicast, fabric manager down,
    fd passing refused) makes ALL ranks fall back to the peer-memory kernels together.  Returns True when the region is mapped."""
    import socket

    def agree(ok):
        v = C.c_double(0.0 if ok else 1.0)
        check(lib.okl_allreduce_max_host(ctx.h, C.byref(v)))
        return v.value == 0.0

    if not agree(bool(lib.okl_mc_supported(ctx.h))):
        if rank == 0:
            print("[okl] NVLS multicast not supported on this box; gradient exchange stays on the peer-memory kernels")
        return False
    path = f"/tmp/okl_mc_{_tag()}.sock"
    ok, srv, conns = True, None, []
    try:
        if rank == 0:
            fd = C.c_int(-1)
            ok = lib.okl_mc_create(ctx.h, nbytes, C.byref(fd)) == 0
            if os.path.exists(path):
                os.unlink(path)
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path)
            srv.listen(world)
            srv.settimeout(60.0)
            for _ in range(world - 1):                 # every peer connects even if creation failed (it then receives "no")
                c, _a = srv.accept()
                conns.append(c)
                if ok:
                    socket.send_fds(c, [b"y"], [fd.value])
                else:
                    c.sendall(b"n")
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 60:
                    raise RuntimeError("timed out waiting for the multicast handle socket")
                time.sleep(0.005)
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.settimeout(60.0)
            for _ in range(200):
                try:
                    c.connect(path)
                    break
                except (ConnectionRefusedError, FileNotFoundError):
                    time.sleep(0.01)
            msg, fds, _f, _a = socket.recv_fds(c, 1, 1)
            conns.append(c)
            ok = msg == b"y" and len(fds) == 1 and lib.okl_mc_import(ctx.h, fds[0], nbytes) == 0
    except (OSError, RuntimeError) as e:
        print(f"[okl] rank {rank}: multicast handle exchange failed: {e}")
        ok = False
    if not agree(ok):
        if rank == 0:
            print(f"[okl] NVLS multicast object could not be shared ({_lib.last_error()}); using the peer-memory kernels")
        return False
    if not agree(lib.okl_mc_add_device(ctx.h) == 0):       # the agreement is also the barrier "every device has been added"
        return False
    ok = lib.okl_mc_bind(ctx.h) == 0
    if not ok:
        print(f"[okl] rank {rank}: okl_mc_bind failed: {_lib.last_error()}")
    ready = agree(ok)
    for c in conns:
        c.close()
    if srv is not None:
        srv.close()
        try:
            os.unlink(path)
        except OSError:
            pass
    if rank == 0:
        print("[okl] NVLS multicast region mapped: gradient exchange = multimem.ld_reduce / multimem.st" if ready else
              "[okl] NVLS bind failed; using the peer-memory kernels", flush=True)
    return ready


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
            ctx.mc_ready = False
            if os.environ.get("OKL_NVLS", "1") != "0":
                ctx.mc_ready = _setup_multicast(ctx, rank, world, int(os.environ.get("OKL_NVLS_BYTES", str(64 << 20))))
        except (RuntimeError, ValueError, MemoryError) as e:      # no peer access on this box: NCCL handles every bucket
            ctx.p2p_ready = False
            print(f"[okl] peer-memory all-reduce unavailable ({e}); using NCCL for all buckets")
    return rank, world
