import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("VRQ_SCAN_MMA", "2")
import vectorragquantization_b200 as V
rng = np.random.default_rng(1)
n, nq = 300000, 128
codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
ix = V.BinaryIndex(1024)
ix.add_with_ids(codes, np.arange(n))
d = ix.distances(q)
ref0 = np.bitwise_count(q[0][None, :] ^ codes).sum(-1).astype(np.int32)
b = np.nonzero(d[0] != ref0)[0]
print("bad rows for q0:", len(b), "first", b[:20])
def perkb(r):  # [nq, 8] distances per K-block
    return np.bitwise_count(q.reshape(nq, 8, 16) ^ codes[r].reshape(1, 8, 16)).sum(-1).astype(np.int32)
for r in b[:12]:
    obs = d[:, r]
    base = perkb(r)
    found = []
    # hypothesis: each kblock kb independently comes from row r + 128*delta_kb; solve greedily per kb using residuals over queries
    # exhaustive over single-kb replacement and parity-class replacement
    for de in range(-6, 7):
        for do in range(-6, 7):
            ra, rb = r + 128 * de, r + 128 * do
            if not (0 <= ra < n and 0 <= rb < n): continue
            pa, pb = perkb(ra), perkb(rb)
            tot = pa[:, 0::2].sum(1) + pb[:, 1::2].sum(1)
            if np.array_equal(tot, obs): found.append(("parity", de, do))
    for kb in range(8):
        for dl in range(-6, 7):
            r2 = r + 128 * dl
            if dl == 0 or not (0 <= r2 < n): continue
            tot = base.sum(1) - base[:, kb] + perkb(r2)[:, kb]
            if np.array_equal(tot, obs): found.append(("kb", kb, dl))
    # partial sums (prefix of kblocks missing => dot partial): obs = pcq - sum_{kb in S} dot_kb ; dot_kb = pc(q_kb & c_kb) - pc(~q_kb & c_kb)
    qk = q.reshape(nq, 8, 16); ck = codes[r].reshape(1, 8, 16)
    dot = (np.bitwise_count(qk & ck).sum(-1).astype(np.int32) - np.bitwise_count(~qk & ck).sum(-1).astype(np.int32))
    pcq = np.bitwise_count(q).sum(1).astype(np.int32)
    for j in range(0, 9):
        if np.array_equal(pcq - dot[:, :j].sum(1), obs): found.append(("prefix", j))
        if np.array_equal(pcq - dot[:, j:].sum(1), obs): found.append(("suffix", j))
    print("row", r, "tile", r // 128, "in", r % 128, "->", found)
