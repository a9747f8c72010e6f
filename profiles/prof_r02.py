"""One driver for every kernel that carries a roofline in bench.py, so that each can be captured by ncu on its own:

    python profiles/prof_r02.py dense      # Phase I, 1024 queries x PROF_ROWS codes (sample pass + dense pass + merges)
    python profiles/prof_r02.py stream     # Phase I at 1, 2, 3, 16, 64, 128 queries per pass
    python profiles/prof_r02.py encode     # the five fused encoders over PROF_ENC_ROWS float32 rows
    python profiles/prof_r02.py rescore    # cfg5: Phase II and Phase III on 4096 x 1000 random candidates of PROF_PAY_ROWS rows

Prints ms per call (CUDA events).  Under ncu use -k regex:<kernel> and --metrics dram__bytes_read.sum,dram__bytes_write.sum
(traffic.json) or --set full (summaries)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vectorragquantization_b200 as V  # noqa: E402
from vectorragquantization_b200 import _lib as L  # noqa: E402

D = 1024
what = sys.argv[1] if len(sys.argv) > 1 else "dense"
reps = int(os.environ.get("PROF_ITERS", 3))
ctx = V.Context(0)
lib = L.load()
dev = torch.device("cuda", 0)
ctx.set_stream(0)
h = ctx.handle


def timed(fn, n=reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def build(rows, payload):
    ix = V.BinaryIndex(D, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW if payload else L.PAYLOAD_NONE)
    ix.reserve(rows)
    for off in range(0, rows, 8_000_000):
        ix.add_synthetic(1, off, min(8_000_000, rows - off), off)
    ctx.sync()
    return ix


def queries(nq, seed=2):
    qf = torch.empty((nq, D), dtype=torch.float32, device=dev)
    qb = torch.empty((nq, D // 8), dtype=torch.uint8, device=dev)
    L.check(lib.vrq_synth_f32(h, seed, 0, nq, D, 0, L.ptr(qf)))
    L.check(lib.vrq_synth_codes_int8(h, seed, 0, nq, D, L.ptr(qb), None))
    return qf, qb


if what in ("dense", "stream"):
    n = int(os.environ.get("PROF_ROWS", 100_000_000))
    ix = build(n, False)
    kk = 1000
    nqs = [int(v) for v in os.environ["PROF_NQS"].split(",")] if os.environ.get("PROF_NQS") else None
    for nq in (nqs or ((1024,) if what == "dense" else (1, 2, 3, 16, 64, 128))):
        _, qb = queries(nq)
        dist = torch.empty((nq, kk), dtype=torch.int32, device=dev)
        lab = torch.empty((nq, kk), dtype=torch.int64, device=dev)
        ms = timed(lambda: L.check(lib.vrq_index_search(ix._h, nq, L.ptr(qb), kk, L.ptr(dist), L.ptr(lab))))
        print(f"phase I, {nq} queries x {n} codes, top-{kk}: {ms:.3f} ms  ({n * 128 / ms / 1e6:.0f} GB/s of codes per pass)", flush=True)
elif what == "encode":
    n = int(os.environ.get("PROF_ENC_ROWS", 10_000_000))
    x = torch.empty((n, D), dtype=torch.float32, device=dev)
    L.check(lib.vrq_synth_f32(h, 7, 0, n, D, 1, L.ptr(x)))
    ub = torch.empty((n, D // 8), dtype=torch.uint8, device=dev)
    q8 = torch.empty((n, D), dtype=torch.int8, device=dev)
    q16 = torch.empty((n, D), dtype=torch.int16, device=dev)
    lo = torch.empty((n,), dtype=torch.float64, device=dev)
    hi = torch.empty((n,), dtype=torch.float64, device=dev)
    cases = {
        "int8_global": (lambda: L.check(lib.vrq_quantize_int8_global(h, L.ptr(x), n, D, 0.3, L.ptr(q8), L.ptr(ub))), 5248),
        "int16_global": (lambda: L.check(lib.vrq_quantize_int16_global(h, L.ptr(x), n, D, 1.0, L.ptr(q16), L.ptr(ub))), 6272),
        "int4": (lambda: L.check(lib.vrq_quantize_int4(h, L.ptr(x), n, D, L.ptr(q8), L.ptr(lo), L.ptr(hi), L.ptr(ub))), 4752),
        "int8_perdoc": (lambda: L.check(lib.vrq_quantize_int8_perdoc(h, L.ptr(x), n, D, L.ptr(q8), L.ptr(lo), L.ptr(hi), L.ptr(ub))), 5256),
        "ubinary": (lambda: L.check(lib.vrq_to_binary_f32(h, L.ptr(x), n, D, 0, L.ptr(ub))), 4224),
    }
    for name, (fn, bpr) in cases.items():
        ms = timed(fn)
        print(f"encode {name}: {ms:.3f} ms, {n * bpr / ms / 1e6:.0f} GB/s", flush=True)
elif what == "rescore":
    n = int(os.environ.get("PROF_PAY_ROWS", 32_000_000))
    ix = build(n, True)
    codes_p, _, pay_p, _ = ix.device_ptrs()
    nq, m = 4096, 1000
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    pos = torch.randint(0, n, (nq, m), dtype=torch.int64, device=dev, generator=g)
    qf, _ = queries(nq, 9)
    sc = torch.empty((nq, m), dtype=torch.float64, device=dev)
    ms3 = timed(lambda: L.check(lib.vrq_rescore_int8cos(h, pay_p, n, D, L.ptr(pos), nq, m, L.ptr(qf), L.ptr(sc))))
    print(f"phase III cfg5 (IMMA={os.environ.get('VRQ_RESCORE_IMMA', 'default')}, shape={os.environ.get('VRQ_RESCORE_IMMA_SHAPE', 'default')}): "
          f"{ms3:.3f} ms, {nq * m * 1024 / ms3 / 1e6:.0f} GB/s gathered", flush=True)
    ms2 = timed(lambda: L.check(lib.vrq_rescore_binary(h, codes_p, n, D, L.ptr(pos), nq, m, L.ptr(qf), L.ptr(sc))))
    print(f"phase II cfg5 (BIN={os.environ.get('VRQ_RESCORE_BIN', 'default')}): {ms2:.3f} ms, {nq * m * 128 / ms2 / 1e6:.0f} GB/s gathered", flush=True)
