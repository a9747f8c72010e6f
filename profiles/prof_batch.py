"""ncu driver for the batched Hamming scan: 256 queries per pass over PROF_ROWS codes (default 16M), top-1000."""
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
n, nq = int(os.environ.get("PROF_ROWS", 16_000_000)), int(os.environ.get("PROF_NQ", 256))
ix = V.BinaryIndex(1024, ctx=ctx)
ix.add_synthetic(1, 0, n, 0)
qx = K.synth_f32(2, 0, nq, ctx=ctx)
qb = torch.from_numpy(np.packbits(qx > 0, axis=1)).to(dev)
dist = torch.empty((nq, 1000), dtype=torch.int32, device=dev)
lab = torch.empty((nq, 1000), dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    e0.record()
    L.check(lib.vrq_index_search(ix._h, nq, L.ptr(qb), 1000, L.ptr(dist), L.ptr(lab)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print("ok", int(lab[0, 0]), int(dist[0, 0]), f"{ms:.2f} ms  {n * nq / ms / 1e6:.1f} Gpair/s")
