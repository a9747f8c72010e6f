import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import numpy as np
import vectorragquantization_b200 as V
from oracle import oracle_c as oc
n=10000; D=1024
x = oc.synth_f32(1, 0, n, D, True)
table = {f"doc {i}": i for i in range(n)}
docs = [f"doc {i}" for i in range(n)]
ctx = V.Context(0)
with tempfile.TemporaryDirectory() as tmp:
    db = V.VectorDBInt8(os.path.join(tmp, "w"), embedder=lambda texts: x[[table[t] for t in texts]], ctx=ctx)
    db.add_documents(list(range(200)), docs[:200], batch_size=64, save=False)
    db = V.VectorDBInt8(os.path.join(tmp, "a"), embedder=lambda texts: x[[table[t] for t in texts]], ctx=ctx)
    pr = cProfile.Profile(); pr.enable()
    t0=time.perf_counter()
    db.add_documents(list(range(n)), docs, batch_size=64, save=False)
    ctx.sync()
    t=time.perf_counter()-t0
    pr.disable()
    print("add_documents", n/t, "docs/s")
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
