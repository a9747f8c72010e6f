import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vectorragquantization_b200 as V
from vectorragquantization_b200 import _lib as L
from vectorragquantization_b200 import kernels as K
ctx = V.Context(0); lib = L.load(); dev = torch.device("cuda", 0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
n = int(os.environ.get("PROF_ROWS", 100_000_000))
ix = V.BinaryIndex(1024, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW)
ix.reserve(n)
for off in range(0, n, 8_000_000):
    ix.add_synthetic(1, off, min(8_000_000, n - off), off)
ctx.sync()
qs = []
for s in range(6):
    qx = K.synth_f32(2, s * 1024, 1024, ctx=ctx)
    qs.append((torch.from_numpy(qx).pin_memory(), torch.from_numpy(np.packbits(qx > 0, axis=1)).pin_memory()))
for i, (qf, qb) in enumerate(qs):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = ix.search3(qf.numpy(), qb.numpy(), 100, 10, 3)
    t1 = time.perf_counter()
    print(f"call {i}: {1e3*(t1-t0):.2f} ms  checksum {int(res[4].sum())}", flush=True)
