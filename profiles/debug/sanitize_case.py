"""Tiny run of the three Phase-I kernels + rescoring for compute-sanitizer memcheck."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vectorragquantization_b200 as V
from vectorragquantization_b200 import _lib as L
rng = np.random.default_rng(3)
n = 3000
codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
pay = rng.integers(-128, 128, (n, 1024), dtype=np.int8)
ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
ix.add_with_ids(codes, np.arange(n), payload=pay)
for nq in (2, 8, 40, 256):
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    qf = rng.standard_normal((nq, 1024)).astype(np.float32)
    d, l = ix.search(q, 50)
    r = ix.search3(qf, q, 10)
    print(nq, int(d.sum()), int(r[4].sum()), flush=True)
print("done")
