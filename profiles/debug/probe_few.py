import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("VRQ_MMA_FEW", "1")
os.environ.setdefault("VRQ_SCAN_MMA", "2")
import vectorragquantization_b200 as V
from oracle import oracle_c as oc
rng = np.random.default_rng(1)
for n, nq, k in [(128, 8, 10), (1000, 6, 100), (5000, 33, 50), (40000, 32, 50), (300001, 17, 100), (3000000, 24, 1000)]:
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    if n <= 400000:
        d = ix.distances(q)
        ref = np.stack([np.bitwise_count(q[i][None, :] ^ codes).sum(-1).astype(np.int32) for i in range(nq)])
        bad = int((d != ref).sum())
        print(f"n={n} nq={nq}: distance mismatches {bad} / {ref.size}", flush=True)
        if bad:
            b = np.nonzero((d != ref).any(0))[0]
            print("  bad rows (first 20):", b[:20], " bad queries:", np.nonzero((d != ref).any(1))[0][:10])
            print("  got", d[0][:8], "ref", ref[0][:8])
    dist, labels = ix.search(q, k)
    rd, rp = oc.hamming_topk(codes, q, k)
    print(f"   k={k}: topk dist ok {np.array_equal(dist, rd)} labels ok {np.array_equal(labels, rp)}", flush=True)
