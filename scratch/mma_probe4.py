import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VRQ_SCAN_MMA", "2")
import vectorragquantization_b200 as V
rng = np.random.default_rng(1)
n, nq = 300000, 128
codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
ix = V.BinaryIndex(1024)
ix.add_with_ids(codes, np.arange(n))
refs = np.stack([np.bitwise_count(q[qi][None, :] ^ codes).sum(-1).astype(np.int32) for qi in (0, 1, 37, 127)])
for cfg in [{}, {"VRQ_MMA_B_STAGES": "8"}, {"VRQ_MMA_RAW_STAGES": "4", "VRQ_MMA_B_STAGES": "8"}, {"VRQ_MMA_B_STAGES": "4"}]:
    for k_, v_ in cfg.items(): os.environ[k_] = v_
    d = ix.distances(q)[[0, 1, 37, 127]]
    bad = d != refs
    print(cfg, "bad per query", bad.sum(1), "rows bad in all 4:", int(bad.all(0).sum()), "rows bad in any:", int(bad.any(0).sum()), flush=True)
    b = np.nonzero(bad.any(0))[0]
    if len(b):
        # which K-block differs? recompute per-kblock distances for the first few bad rows
        for r in b[:6]:
            got = d[0][r]; ref = refs[0][r]
            per_kb = np.bitwise_count(q[0].reshape(8, 16) ^ codes[r].reshape(8, 16)).sum(1)
            # try: one kblock replaced by the same kblock of row r' = r - 128*j
            expl = []
            for back in range(1, 40):
                r2 = r - 128 * back
                if r2 < 0: break
                per2 = np.bitwise_count(q[0].reshape(8, 16) ^ codes[r2].reshape(8, 16)).sum(1)
                for kb in range(8):
                    if ref - per_kb[kb] + per2[kb] == got: expl.append((back, kb))
            print("   row", r, "tile", r // 128, "in-tile", r % 128, "got", got, "ref", ref, "explained by (tiles back, kb):", expl[:6])
    for k_ in cfg: os.environ.pop(k_)
