"""Probe of the tensor-core scan: distance matrix and top-k vs NumPy on small inputs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorragquantization_b200 as V
from oracle import oracle_c as oc

os.environ.setdefault("VRQ_SCAN_MMA", "2")
rng = np.random.default_rng(1)
for n, nq, k in [(128, 128, 10), (1000, 5, 100), (5000, 200, 50), (300000, 300, 100)]:
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    if n <= 5000:
        d = ix.distances(q)
        ref = np.bitwise_count(q[:, None, :] ^ codes[None, :, :]).sum(-1).astype(np.int32)
        bad = int((d != ref).sum())
        print(f"n={n} nq={nq}: distance mismatches {bad} / {d.size}", flush=True)
        if bad:
            print("got", d[:4, :8]); print("ref", ref[:4, :8])
            print("diff rows", np.unique(np.nonzero(d != ref)[0])[:20], "cols", np.unique(np.nonzero(d != ref)[1])[:20])
    dist, labels = ix.search(q, k)
    rd, rp = oc.hamming_topk(codes, q, k)
    print(f"n={n} nq={nq} k={k}: topk dist ok {np.array_equal(dist, rd)} labels ok {np.array_equal(labels, rp)}", flush=True)
