import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VRQ_SCAN_MMA", "2")
import vectorragquantization_b200 as V
from oracle import oracle_c as oc
rng = np.random.default_rng(1)
for n, nq in [(40000, 5), (300000, 130)]:
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    d = ix.distances(q)
    bad_rows = []
    nbad = 0
    for qi in range(nq):
        ref = np.bitwise_count(q[qi][None, :] ^ codes).sum(-1).astype(np.int32)
        b = np.nonzero(d[qi] != ref)[0]
        nbad += len(b)
        if len(b) and len(bad_rows) < 3:
            bad_rows.append((qi, b[:10], d[qi][b[:10]], ref[b[:10]]))
    print(f"n={n} nq={nq}: distance mismatches {nbad}", flush=True)
    for br in bad_rows: print(br)
    for k in (10, 100):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        okd, okl = np.array_equal(dist, rd), np.array_equal(labels, rp)
        print(f"  k={k}: topk dist ok {okd} labels ok {okl}", flush=True)
        if not okd:
            bq = np.nonzero((dist != rd).any(1))[0]
            print("  bad queries", bq[:20], "of", nq)
            qi = bq[0]
            print("  got", dist[qi][:12], labels[qi][:12]); print("  ref", rd[qi][:12], rp[qi][:12])
