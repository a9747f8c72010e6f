"""cProfile of the cfg1 class-API search path (VectorDBInt8.search, one query per call, 10 k documents)."""
import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import numpy as np
import vectorragquantization_b200 as V
from oracle import oracle_c as oc
n, nq, D = 10000, 100, 1024
x = oc.synth_f32(1, 0, n, D, True)
qx = oc.synth_f32(2, 0, nq, D, True)
table = {f"doc {i}": i for i in range(n)}
table.update({f"query {i}": n + i for i in range(nq)})
allx = np.concatenate([x, qx])
docs = [f"doc {i}" for i in range(n)]
ctx = V.Context(0)
with tempfile.TemporaryDirectory() as tmp:
    db = V.VectorDBInt8(os.path.join(tmp, "a"), embedder=lambda texts: allx[[table[t] for t in texts]], ctx=ctx)
    db.add_documents(list(range(n)), docs, batch_size=64, save=False)
    for qi in range(5):
        db.search(f"query {qi}", k=10, binary_oversample=10)
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter()
    for qi in range(nq):
        db.search(f"query {qi}", k=10, binary_oversample=10)
    t = time.perf_counter() - t0
    pr.disable()
    print("search", nq / t, "queries/s", t / nq * 1e6, "us per query")
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
