"""Hamming top-1000 over PROF_ROWS codes at 1..64 queries per pass: where the scan leaves the HBM roofline."""
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
n = int(os.environ.get("PROF_ROWS", 100_000_000))
ix = V.BinaryIndex(1024, ctx=ctx)
ix.reserve(n)
for off in range(0, n, 8_000_000):
    ix.add_synthetic(1, off, min(8_000_000, n - off), off)
qx = K.synth_f32(2, 0, 64, ctx=ctx)
qb = torch.from_numpy(np.packbits(qx > 0, axis=1)).to(dev)
dist = torch.empty((64, 1000), dtype=torch.int32, device=dev)
lab = torch.empty((64, 1000), dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nq in (1, 2, 3, 4, 6, 8, 9, 16, 32, 64):
    for it in range(4):
        e0.record()
        L.check(lib.vrq_index_search(ix._h, nq, L.ptr(qb), 1000, L.ptr(dist), L.ptr(lab)))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"nq={nq:3d}  {ms:8.3f} ms  {n * 128 / ms / 1e6:7.0f} GB/s of codes  {n * nq / ms / 1e6:7.1f} Gpair/s  {nq / ms * 1e3:8.1f} QPS")
