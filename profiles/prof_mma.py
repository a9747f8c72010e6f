"""Tensor-core (scan_mma.cu) vs integer-pipe (scan.cu) Hamming top-1000 over PROF_ROWS codes, 1024-query batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vectorragquantization_b200 as V  # noqa: E402
from vectorragquantization_b200 import _lib as L  # noqa: E402
from vectorragquantization_b200 import kernels as K  # noqa: E402

ctx = V.Context(0)
lib = L.load()
dev = torch.device("cuda", 0)
ctx.set_stream(0)
n = int(os.environ.get("PROF_ROWS", 32_000_000))
nqs = [int(x) for x in os.environ.get("PROF_NQ", "1024").split(",")]
modes = os.environ.get("PROF_MODES", "1,0").split(",")
k = int(os.environ.get("PROF_K", 1000))
ix = V.BinaryIndex(1024, ctx=ctx)
ix.reserve(n)
for off in range(0, n, 8_000_000):
    ix.add_synthetic(1, off, min(8_000_000, n - off), off)
qx = K.synth_f32(2, 0, max(nqs), ctx=ctx)
qb = torch.from_numpy(np.packbits(qx > 0, axis=1)).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for nq in nqs:
    for mode in modes:
        os.environ["VRQ_SCAN_MMA"] = mode
        dist = torch.empty((nq, k), dtype=torch.int32, device=dev)
        lab = torch.empty((nq, k), dtype=torch.int64, device=dev)
        best = 1e30
        for it in range(int(os.environ.get("PROF_ITERS", 3))):
            e0.record()
            L.check(lib.vrq_index_search(ix._h, nq, L.ptr(qb), k, L.ptr(dist), L.ptr(lab)))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[(nq, mode)] = (dist.cpu(), lab.cpu())
        print(f"nq={nq:5d} mma={mode}  {best:9.3f} ms  {n * nq / best / 1e6:8.1f} Gpair/s  {nq / best * 1e3 * n / 1e8:9.1f} QPS@100M", flush=True)
    if len(modes) > 1:
        a, b = res[(nq, modes[0])], res[(nq, modes[1])]
        print(f"    identical results across kernels: {bool((a[0] == b[0]).all() and (a[1] == b[1]).all())}", flush=True)
