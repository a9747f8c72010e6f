"""torchrun worker for tests/test_gpu_multi.py: row-sharded 3-phase search over NCCL vs the same search on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    from vectorragquantization_b200.sharded import CudaEngine, ShardedSearch3, shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, nq, k, bo, io = 3_000_001, 64, 10, 10, 3
    ctx = V.Context(local)
    a, b = shard_range(n, rank, world)
    ix = V.BinaryIndex(1024, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_synthetic(11, a, b - a, a + 500)
    from vectorragquantization_b200 import kernels as K
    qf = K.synth_f32(11, 7, nq, ctx=ctx) + K.synth_f32(12, 0, nq, ctx=ctx) * np.float32(0.5)
    qb = np.packbits(qf > 0, axis=1)
    s = ShardedSearch3(CudaEngine(ix, ctx), pos_base=a)
    assert s.ntotal == n
    out = s.search(qf, qb, k, bo, io)
    torch.cuda.synchronize()
    got = [out[x].cpu().numpy() for x in ("labels", "hamming", "score_binary", "score_cosine", "count")]
    ok = True
    if rank == 0:
        full = V.BinaryIndex(1024, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW)
        full.add_synthetic(11, 0, n, 500)
        want = full.search3(qf, qb, k, bo, io)
        ok = all(np.array_equal(x, y) for x, y in zip(want, got))
        print("MULTI_GPU_PARITY", "OK" if ok else "MISMATCH", "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
