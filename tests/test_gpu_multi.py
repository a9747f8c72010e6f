"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): torchrun with one process per GPU, NCCL
all-gather, device merge - bit-identical to the single-GPU search on the concatenated database."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_search_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MULTI_GPU_PARITY OK" in r.stdout


def test_sharded_search_inside_the_c_abi_single_process():
    """The sharded search as ONE C call, no torch.distributed: vrq_nccl_init_all (one process, all GPUs) +
    vrq_search3_sharded_group == the single-index search3 over the concatenated database.  Runs in a subprocess so that
    NCCL's per-process state does not leak into the other tests."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multi_gpu_cabi_worker.py"), str(world)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "CABI_SHARDED_PARITY OK" in r.stdout


def test_sharded_search_inside_the_c_abi_world1():
    """World size 1 needs no NCCL at all: vrq_index_search3_sharded == vrq_index_search3 (device pointers)."""
    import numpy as np
    import torch
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    from vectorragquantization_b200 import kernels as K
    n, nq, k, bo, io = 200_000, 20, 10, 10, 3
    ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_synthetic(21, 0, n, 100)
    qf = K.synth_f32(21, 5, nq) + K.synth_f32(22, 0, nq) * np.float32(0.5)
    qb = np.packbits(qf > 0, axis=1)
    want = ix.search3(qf, qb, k, bo, io)
    dev = torch.device("cuda", ix.ctx.device)
    qf_d, qb_d = torch.from_numpy(qf).to(dev), torch.from_numpy(qb).to(dev)
    out = [torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.int32, device=dev),
           torch.empty((nq, k), dtype=torch.float64, device=dev), torch.empty((nq, k), dtype=torch.float64, device=dev),
           torch.empty(nq, dtype=torch.int32, device=dev)]
    torch.cuda.synchronize()
    L.check(L.load().vrq_index_search3_sharded(ix._h, nq, L.ptr(qf_d), L.ptr(qb_d), k, bo, io, 0, n, *[L.ptr(t) for t in out]))
    ix.ctx.sync()
    for a, b in zip(want, out):
        assert np.array_equal(a, b.cpu().numpy())
