"""Small Phase-I searches through every tensor-core kernel of scan_mma.cu, checked against the CPU oracle - a quick driver for
a debugger or (where the pool allows it; on this one compute-sanitizer is closed) `compute-sanitizer --tool memcheck`:

    python profiles/san_r02.py

3, 40 and 80 queries take the three forms of the swapped-operand kernel, 128 the single-CTA 128-query-tile kernel, 1024 the CTA
pairs with the lean epilogue; results are checked against the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import vectorragquantization_b200 as V  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402

n = int(os.environ.get("SAN_ROWS", 60000))
rng = np.random.default_rng(5)
codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
ix = V.BinaryIndex(1024)
ix.add_with_ids(codes, np.arange(n))
for nq in (3, 40, 80, 128, 1024):
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    q[0] = codes[n // 2]
    for k in (10, 300):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q[:16], k)
        assert np.array_equal(dist[:16], rd) and np.array_equal(labels[:16], rp), (nq, k)
    print("ok", nq, flush=True)
