"""ncu driver for Phase III (BASELINE config 5): 4096 queries x 1000 gathered int8 candidates over PROF_ROWS rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vectorragquantization_b200 as V  # noqa: E402
from vectorragquantization_b200 import _lib as L  # noqa: E402

ctx = V.Context(0)
lib = L.load()
dev = torch.device("cuda", 0)
ctx.set_stream(0)
n = int(os.environ.get("PROF_ROWS", 32_000_000))
span = int(os.environ.get("PROF_SPAN", n))
nq, m = int(os.environ.get("PROF_NQ", 4096)), 1000
ix = V.BinaryIndex(1024, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW)
ix.reserve(n)
for off in range(0, n, 8_000_000):
    ix.add_synthetic(1, off, min(8_000_000, n - off), off)
codes_p, _, pay_p, _ = ix.device_ptrs()
g = torch.Generator(device=dev)
g.manual_seed(5)
pos = torch.randint(0, span, (nq, m), dtype=torch.int64, device=dev, generator=g)
qf = torch.empty((nq, 1024), dtype=torch.float32, device=dev)
L.check(lib.vrq_synth_f32(ctx.handle, 9, 0, nq, 1024, 0, L.ptr(qf)))
sc = torch.empty((nq, m), dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for which, fn, ptr, bpr in (("int8cos", lib.vrq_rescore_int8cos, pay_p, 1024), ("binary", lib.vrq_rescore_binary, codes_p, 128)):
    for it in range(3):
        e0.record()
        L.check(fn(ctx.handle, ptr, n, 1024, L.ptr(pos), nq, m, L.ptr(qf), L.ptr(sc)))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{which}: span {span} rows  {ms:.3f} ms  {nq * m / ms / 1e6:.2f} Gpair/s  {nq * m * bpr / ms / 1e6:.0f} GB/s gathered")
