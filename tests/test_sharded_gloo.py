"""world_size-2 (and 3) CPU test of the multi-GPU host logic in vectorragquantization_b200/sharded.py with the gloo
backend: shard ranges, global positions, the packed all-gather layout and the merge call.  The device steps are
played by an ORACLE-backed engine injected by this test (test infrastructure) - the product engine is CUDA only."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """CPU stand-in for CudaEngine: same packed layouts, oracle arithmetic."""

    def __init__(self, codes, ids, i8):
        self.codes, self.ids, self.i8 = codes, ids, i8
        self.device = torch.device("cpu")

    def local_ntotal(self):
        return self.codes.shape[0]

    def search3_local(self, qf, qb, nq, bk, pos_base, packed):
        from oracle import oracle_c as oc
        from oracle import vrq_oracle as o
        qf, qb = qf.numpy(), qb.numpy()
        dist_, pos = oc.hamming_topk(self.codes, qb, bk, pos_base=pos_base)
        keys = np.where(pos >= 0, (dist_.astype(np.int64) << 40) | pos, -1)  # -1 == ~0 as uint64
        labels = np.where(pos >= 0, self.ids[np.clip(pos - pos_base, 0, None)], -1)
        sb = np.full((nq, bk), -np.inf)
        sc = np.full((nq, bk), -np.inf)
        for qi in range(nq):
            ok = pos[qi] >= 0
            lp = pos[qi][ok] - pos_base
            sb[qi, ok] = o.rescore_binary(qf[qi], self.codes[lp])
            sc[qi, ok] = o.rescore_int8cos(qf[qi], self.i8[lp])
        packed[0] = torch.from_numpy(keys)
        packed[1] = torch.from_numpy(labels)
        packed[2] = torch.from_numpy(sb.view(np.int64))
        packed[3] = torch.from_numpy(sc.view(np.int64))

    def merge3(self, world, nq, bk, gathered, k, k2, out):
        from oracle import vrq_oracle as o
        g = gathered.numpy()
        for qi in range(nq):
            shards = []
            for w in range(world):
                keys = g[w, 0, qi]
                ok = keys != -1
                shards.append([{"score_hamming": int(kk >> 40), "pos": int(kk & ((1 << 40) - 1)), "doc_id": int(l),
                                "score_binary": float(b), "score_cosine": float(c)}
                               for kk, l, b, c in zip(keys[ok], g[w, 1, qi][ok], g[w, 2, qi].view(np.float64)[ok],
                                                      g[w, 3, qi].view(np.float64)[ok])])
            bo = 1
            merged = o.merge_shard_results(shards, k, bk // k if bk % k == 0 else bo, k2 // k)
            out["count"][qi] = len(merged)
            out["labels"][qi] = -1
            for i, h in enumerate(merged):
                out["labels"][qi, i] = h["doc_id"]
                out["hamming"][qi, i] = h["score_hamming"]
                out["score_binary"][qi, i] = h["score_binary"]
                out["score_cosine"][qi, i] = h["score_cosine"]


def _worker(rank, world, port, n, nq, k, bo, io, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_c as oc
        from oracle import vrq_oracle as o
        from vectorragquantization_b200.sharded import ShardedSearch3, shard_range
        a, b = shard_range(n, rank, world)
        codes, i8 = oc.synth_codes_int8(71, a, b - a)  # every rank regenerates ITS rows from the counter-based generator
        ids = np.arange(a, b, dtype=np.int64) * 5 + 2
        eng = OracleEngine(codes, ids, i8)
        s = ShardedSearch3(eng, pos_base=a)
        assert s.ntotal == n and s.world == world
        qf = (oc.synth_f32(71, 0, nq) + oc.synth_f32(72, 0, nq) * np.float32(0.6)).astype(np.float32)
        qb = o.synth_ubinary_from_f32(qf)
        out = s.search(qf, qb, k, bo, io)
        res = {kk: v.clone().numpy() for kk, v in out.items()}
        if rank == 0:
            q.put(res)
        else:
            q.put(None)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_matches_unsharded_oracle(world):
    from oracle import oracle_c as oc
    from oracle import vrq_oracle as o
    n, nq, k, bo, io = 9000, 5, 8, 10, 3
    port = 29500 + os.getpid() % 2000 + world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, nq, k, bo, io, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = next(r for r in results if r is not None)
    codes, i8 = oc.synth_codes_int8(71, 0, n)
    ids = np.arange(n, dtype=np.int64) * 5 + 2
    qf = (oc.synth_f32(71, 0, nq) + oc.synth_f32(72, 0, nq) * np.float32(0.6)).astype(np.float32)
    qb = o.synth_ubinary_from_f32(qf)
    for qi in range(nq):
        ref = o.search3(codes, ids, i8, qf[qi], qb[qi], k, bo, io)
        assert res["count"][qi] == len(ref)
        assert res["labels"][qi][: len(ref)].tolist() == [h["doc_id"] for h in ref]
        assert res["hamming"][qi][: len(ref)].tolist() == [h["score_hamming"] for h in ref]
        assert np.array_equal(res["score_cosine"][qi][: len(ref)], np.array([h["score_cosine"] for h in ref]))
