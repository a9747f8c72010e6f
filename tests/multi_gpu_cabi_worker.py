"""Worker for tests/test_gpu_multi.py: one process drives `world` GPUs through the C ABI alone (ctypes; torch only owns
the device buffers): vrq_nccl_init_all -> vrq_ctx_set_nccl -> vrq_search3_sharded_group, compared with the single-index
search over the concatenated database."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    from vectorragquantization_b200 import kernels as K
    from vectorragquantization_b200.sharded import shard_range
    world = int(sys.argv[1])
    lib = L.load()
    n, nq, k, bo, io = 3_000_001, 64, 10, 10, 3
    devs = (C.c_int * world)(*range(world))
    comms = (C.c_void_p * world)()
    L.check(lib.vrq_nccl_init_all(world, devs, comms))
    ctxs, ixs, bases = [], [], []
    for r in range(world):
        ctx = V.Context(r)
        L.check(lib.vrq_ctx_set_nccl(ctx.handle, C.c_void_p(comms[r]), r, world))
        a, b = shard_range(n, r, world)
        ix = V.BinaryIndex(1024, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW)
        ix.add_synthetic(11, a, b - a, a + 500)
        ctx.sync()
        ctxs.append(ctx)
        ixs.append(ix)
        bases.append(a)
    qf = K.synth_f32(11, 7, nq, ctx=ctxs[0]) + K.synth_f32(12, 0, nq, ctx=ctxs[0]) * np.float32(0.5)
    qb = np.packbits(qf > 0, axis=1)
    bufs = []
    for r in range(world):
        dev = torch.device("cuda", r)
        bufs.append({"qf": torch.from_numpy(qf).to(dev), "qb": torch.from_numpy(qb).to(dev),
                     "labels": torch.empty((nq, k), dtype=torch.int64, device=dev), "ham": torch.empty((nq, k), dtype=torch.int32, device=dev),
                     "sb": torch.empty((nq, k), dtype=torch.float64, device=dev), "sc": torch.empty((nq, k), dtype=torch.float64, device=dev),
                     "cnt": torch.empty(nq, dtype=torch.int32, device=dev)})
        torch.cuda.synchronize(dev)
    arr = lambda key: (C.c_void_p * world)(*[bufs[r][key].data_ptr() for r in range(world)])  # noqa: E731
    ix_arr = (C.c_void_p * world)(*[ix._h.value for ix in ixs])
    pos_base = (C.c_int64 * world)(*bases)
    for _ in range(3):  # more than one exchange through the same communicators
        L.check(lib.vrq_search3_sharded_group(world, ix_arr, nq, arr("qf"), arr("qb"), k, bo, io, pos_base, n, arr("labels"), arr("ham"),
                                              arr("sb"), arr("sc"), arr("cnt")))
    for c in ctxs:
        c.sync()
    full = V.BinaryIndex(1024, ctx=ctxs[0], payload_kind=L.PAYLOAD_INT8_RAW)
    full.add_synthetic(11, 0, n, 500)
    want = full.search3(qf, qb, k, bo, io)
    ok = True
    for r in range(world):
        got = [bufs[r][x].cpu().numpy() for x in ("labels", "ham", "sb", "sc", "cnt")]
        ok &= all(np.array_equal(x, y) for x, y in zip(want, got))
    print("CABI_SHARDED_PARITY", "OK" if ok else "MISMATCH", "world", world, flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
