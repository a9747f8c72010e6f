"""Phase III (vrq_rescore_int8cos) on the cfg5 shape: 4096 queries x 1000 random candidates out of PROF_ROWS int8 rows.
Prints ms per launch (CUDA events) and the largest relative deviation from a float64 torch reference on a few queries.
The kernel variant is chosen by the environment (VRQ_RESCORE_DP2A, VRQ_RESCORE_ASYNC), read once per process."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vectorragquantization_b200 as V  # noqa: E402
from vectorragquantization_b200 import _lib as L  # noqa: E402

D = 1024
ctx = V.Context(0)
lib = L.load()
dev = torch.device("cuda", 0)
ctx.set_stream(0)
n = int(os.environ.get("PROF_ROWS", 16_000_000))
nq, m = int(os.environ.get("PROF_NQ", 4096)), int(os.environ.get("PROF_M", 1000))
g = torch.Generator(device=dev)
g.manual_seed(5)
rows = torch.empty((n, D), dtype=torch.int8, device=dev)
for off in range(0, n, 2_000_000):
    rows[off:off + 2_000_000] = torch.randint(-128, 128, (min(2_000_000, n - off), D), dtype=torch.int8, device=dev, generator=g)
rows[12345] = 0  # a zero row: score must be -inf
pos = torch.randint(0, n, (nq, m), dtype=torch.int64, device=dev, generator=g)
pos[0, 0] = 12345
pos[1, 5] = -1   # a missing candidate
qf = torch.randn((nq, D), dtype=torch.float32, device=dev, generator=g) * 0.05
qf[2] = 0        # an all-zero query
qf[3, ::2] *= 1e-12  # a query with a wide exponent range
sc = torch.empty((nq, m), dtype=torch.float64, device=dev)
fn = lambda: L.check(lib.vrq_rescore_int8cos(ctx.handle, L.ptr(rows), n, D, L.ptr(pos), nq, m, L.ptr(qf), L.ptr(sc)))
fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get("PROF_ITERS", 5))
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
worst = 0.0
for qi in (0, 1, 2, 3, 7, nq - 1):
    p = pos[qi].clamp(min=0)
    r = rows[p].double()
    ref = (r @ qf[qi].double()) / r.pow(2).sum(1).sqrt()
    ref[r.pow(2).sum(1) == 0] = float("-inf")
    ref[pos[qi] < 0] = float("-inf")
    got = sc[qi]
    fin = torch.isfinite(ref)
    assert bool((got[~fin] == ref[~fin]).all()), qi
    scale = (r.abs() @ qf[qi].double().abs()) / r.pow(2).sum(1).sqrt().clamp(min=1)  # sum |q_i x_i| / |x|
    dev_rel = ((got[fin] - ref[fin]).abs() / scale[fin].clamp(min=1e-300)).max().item() if fin.any() else 0.0
    worst = max(worst, dev_rel)
print(f"variant dp2a={os.environ.get('VRQ_RESCORE_DP2A', '0')} async={os.environ.get('VRQ_RESCORE_ASYNC', '1')}: {ms:.3f} ms, "
      f"{nq * m * 1024 / ms / 1e6:.0f} GB/s gathered, worst |got - ref| / (sum|q x| / |x|) = {worst:.3e}", flush=True)
