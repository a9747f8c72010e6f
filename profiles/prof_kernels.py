"""Small driver for ncu: one fused int8-global encode (1M rows) and Hamming scans at 1, 2 and 64 queries per pass over
32M codes.  Usage: python profiles/prof_kernels.py   (see profiles/README.md for the ncu command lines)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorragquantization_b200 as V  # noqa: E402
from vectorragquantization_b200 import _lib as L  # noqa: E402
from vectorragquantization_b200 import kernels as K  # noqa: E402

import torch  # noqa: E402

ctx = V.Context(0)
lib = L.load()
dev = torch.device("cuda", 0)
ctx.set_stream(0)
n_enc, n = 1_000_000, int(os.environ.get("PROF_ROWS", 32_000_000))
x = torch.empty((n_enc, 1024), dtype=torch.float32, device=dev)
L.check(lib.vrq_synth_f32(ctx.handle, 7, 0, n_enc, 1024, 1, L.ptr(x)))
q8 = torch.empty((n_enc, 1024), dtype=torch.int8, device=dev)
ub = torch.empty((n_enc, 128), dtype=torch.uint8, device=dev)
for _ in range(2):
    L.check(lib.vrq_quantize_int8_global(ctx.handle, L.ptr(x), n_enc, 1024, 0.3, L.ptr(q8), L.ptr(ub)))
ix = V.BinaryIndex(1024, ctx=ctx)
ix.add_synthetic(1, 0, n, 0)
qx = K.synth_f32(2, 0, 64, ctx=ctx)
qb = torch.from_numpy(np.packbits(qx > 0, axis=1)).to(dev)
dist = torch.empty((64, 1000), dtype=torch.int32, device=dev)
lab = torch.empty((64, 1000), dtype=torch.int64, device=dev)
for nq in (1, 1, 2, 64):
    L.check(lib.vrq_index_search(ix._h, nq, L.ptr(qb), 1000, L.ptr(dist), L.ptr(lab)))
torch.cuda.synchronize()
print("ok", int(lab[0, 0]), int(dist[0, 0]))
