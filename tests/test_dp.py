"""Data-parallel path: world_size-2 gloo test of the sharding / global-denominator / sum-allreduce scheme on
CPU, and (GPU box with >= 2 devices) the product's NCCL path against the single-process oracle."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_workers(mode, world, port):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "res.json")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), mode, out]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.load(open(out))


def test_dp_scheme_gloo_world2():
    res = run_workers("gloo", 2, 29541)
    assert res["loss_err"] < 1e-12 and res["w_err"] < 1e-12


def test_rendezvous_file_roundtrip(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    class FakeCtx:
        @staticmethod
        def comm_unique_id():
            return bytes(range(128))

    class FakeLib:
        Context = FakeCtx
    monkeypatch.setenv("MASTER_PORT", "45999")
    monkeypatch.setenv("TORCHELASTIC_RUN_ID", f"t{os.getpid()}")
    from okrolearn_b200 import dist
    monkeypatch.setattr(dist._lib, "Context", FakeCtx)
    a = dist.exchange_unique_id(0, 2)
    b = dist.exchange_unique_id(1, 2)
    assert a == b == bytes(range(128))


def _gpu_count():
    try:
        return len([l for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if l.strip()])
    except OSError:
        return 0


@pytest.mark.gpu
def test_dp_nccl_world2_matches_single_process_oracle():
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = run_workers("nccl", 2, 29543)
    assert res["w_err"] < 2e-5
    os.environ["OKL_P2P"] = "0"                      # and once more with the NCCL exchange instead of the peer-memory kernel
    try:
        res2 = run_workers("nccl", 2, 29545)
    finally:
        del os.environ["OKL_P2P"]
    assert res2["w_err"] < 2e-5 and not res2["p2p"]
    os.environ["OKL_EARLY_EXCHANGE"] = "0"           # one all-reduce after the whole backward instead of the overlapped tail
    try:
        res3 = run_workers("nccl", 2, 29547)
    finally:
        del os.environ["OKL_EARLY_EXCHANGE"]
    assert res3["w_err"] < 2e-5
