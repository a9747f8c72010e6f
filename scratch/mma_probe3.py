import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VRQ_SCAN_MMA", "2")
import vectorragquantization_b200 as V
rng = np.random.default_rng(1)
n, nq = 300000, 130
codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
ix = V.BinaryIndex(1024)
ix.add_with_ids(codes, np.arange(n))
d = ix.distances(q)
for qi in (0, 129):
    ref = np.bitwise_count(q[qi][None, :] ^ codes).sum(-1).astype(np.int32)
    bad = d[qi] != ref
    b = np.nonzero(bad)[0]
    print("query", qi, "nbad", len(b))
    # per strip of 32 tiles (4096 rows): list bad (tile_in_strip, 32-col group)
    tiles = b // 128
    groups = (b % 128) // 32
    import collections
    c = collections.Counter(zip((tiles % 32).tolist(), groups.tolist()))
    print(sorted(c.items())[:80])
    # does the wrong value equal the distance of some other row nearby?
    for r in b[:5]:
        cand = np.nonzero(ref[max(0, r - 2048):r + 2048] == d[qi][r])[0] + max(0, r - 2048)
        print(r, d[qi][r], ref[r], "rows with that ref value nearby:", cand[:10] - r)
