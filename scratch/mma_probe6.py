import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorragquantization_b200 as V
from oracle import oracle_c as oc
from oracle import vrq_oracle as o
n, nq = 3_000_000, 300
codes, _ = oc.synth_codes_int8(61, 0, n, want_int8=False)
qx = oc.synth_f32(62, 0, nq)
q = o.synth_ubinary_from_f32(qx)
q[0] = codes[n - 5]
ix = V.BinaryIndex(1024)
ix.add_with_ids(codes, np.arange(n))
refs = {}
for k in (100, 1000):
    refs[k] = oc.hamming_topk(codes, q, k)
for name, env in [("default", {}), ("forced fallback", {"VRQ_MMA_SAMPLE_K": "1", "VRQ_MMA_SAFETY": "1"}), ("no sampling", {"VRQ_MMA_SAFETY": "0"}), ("integer pipes", {"VRQ_SCAN_MMA": "0"})]:
    for a, b in env.items(): os.environ[a] = b
    for k in (100, 1000):
        ix.search(q, k)
        t0 = time.time()
        dist, labels = ix.search(q, k)
        dt = time.time() - t0
        rd, rp = refs[k]
        print(f"{name:16s} k={k}: dist ok {np.array_equal(dist, rd)} labels ok {np.array_equal(labels, rp)}  {dt*1e3:.1f} ms", flush=True)
    for a in env: os.environ.pop(a)
