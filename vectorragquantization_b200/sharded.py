"""Row-sharded 3-phase search over the GPUs of one box: one process per GPU (torchrun), ``torch.distributed`` for the
plumbing, one all-gather per query batch, device-side merge.

The reference has no distributed path at all; this is the B200-native scaling of
``CohereEnhancedVectorDB.search``: rank r holds rows [r*N/W, (r+1)*N/W) of the codes and int8 vectors.  Per batch:

    every rank:   phase I top-binary_k of ITS shard  -> keys (hamming<<40 | GLOBAL position)
                  phase II and III scores for those candidates (pure functions of (query, document), so computing
                  them before the exchange is exact)                                     [vrq_index_search3_local]
    one all_gather of the packed [4, nq, binary_k] 8-byte records over NCCL / NVLink
    every rank:   global phase-I cut by (hamming, position) -> stable sort by score_binary -> cut ->
                  stable sort by score_cosine -> k                                        [vrq_merge3]

The result is bit-identical to the single-GPU ``search3`` on the concatenated database (tests/test_gpu_index.py::
test_merge3_equals_single_index and tests/test_sharded_gloo.py).

``engine`` is the object that runs the two device steps.  The product engine is ``CudaEngine`` (libvrq); the CPU
tests inject an oracle-backed engine to exercise this file's host logic with the gloo backend - the product never
does.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import _lib as L


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``: [rank*N/W, (rank+1)*N/W) (integer arithmetic, covers N exactly)."""
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


class CudaEngine:
    """libvrq-backed steps on torch CUDA tensors (torch = memory + streams only)."""

    def __init__(self, index, ctx=None):
        import torch
        self.torch = torch
        self.index = index
        self.ctx = ctx if ctx is not None else index.ctx
        self.device = torch.device("cuda", self.ctx.device)
        self._lib = L.load()

    def bind_stream(self):
        self.ctx.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)

    def local_ntotal(self) -> int:
        return self.index.ntotal

    def search3_local(self, qf, qb, nq, bk, pos_base, packed):
        # packed: int64[4, nq, bk] -> keys, labels, score_binary (f64 bits), score_cosine (f64 bits)
        self.index.search3_local_into(qf, qb, nq, bk, pos_base, packed[0], packed[1], packed[2], packed[3])

    def merge3(self, world, nq, bk, gathered, k, k2, out):
        # gathered: int64[world, 4, nq, bk]; the four arrays of rank w start at gathered[w, a]; rank stride = 4*nq*bk
        g = gathered
        L.check(self._lib.vrq_merge3(self.ctx.handle, world, nq, bk, 4 * nq * bk, L.ptr(g[0, 0]), L.ptr(g[0, 1]), L.ptr(g[0, 2]),
                                     L.ptr(g[0, 3]), k, k2, L.ptr(out["labels"]), L.ptr(out["hamming"]),
                                     L.ptr(out["score_binary"]), L.ptr(out["score_cosine"]), L.ptr(out["count"])))


class ShardedSearch3:
    """3-phase search over a row-sharded database.  ``dist`` must be initialised (nccl for CUDA, gloo in CPU tests)."""

    def __init__(self, engine, pos_base: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.pos_base = int(pos_base)
        self.device = engine.device
        n = torch.tensor([engine.local_ntotal()], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(n, group=group)
        self.ntotal = int(n.item())
        self._bufs = {}

    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self.torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def search(self, q_float, q_ubinary, k: int = 10, binary_oversample: int = 10, int8_oversample: int = 3):
        """q_float: float32[nq, D], q_ubinary: uint8[nq, D/8] - identical on every rank, already on ``self.device``
        (torch tensors) or NumPy arrays.  Returns a dict of device tensors
        {labels i64[nq,k], hamming i32, score_binary f64, score_cosine f64, count i32[nq]} on every rank."""
        torch = self.torch
        if isinstance(q_float, np.ndarray):
            q_float = torch.from_numpy(np.ascontiguousarray(q_float, np.float32)).to(self.device)
        if isinstance(q_ubinary, np.ndarray):
            q_ubinary = torch.from_numpy(np.ascontiguousarray(q_ubinary, np.uint8)).to(self.device)
        nq = q_float.shape[0]
        bk = min(k * binary_oversample, self.ntotal)  # binary_k = min(k*oversample, ntotal)  (:267), GLOBAL ntotal
        out = {"labels": self._buf("labels", (nq, k), torch.int64), "hamming": self._buf("hamming", (nq, k), torch.int32),
               "score_binary": self._buf("sb", (nq, k), torch.float64), "score_cosine": self._buf("sc", (nq, k), torch.float64),
               "count": self._buf("count", (nq,), torch.int32)}
        if bk == 0:
            out["count"].zero_()
            out["labels"].fill_(-1)
            return out
        if bk > L.MAX_K:
            raise L.VrqError(L.ERR_UNSUPPORTED, f"k * binary_oversample = {bk} exceeds the supported {L.MAX_K}")
        if hasattr(self.engine, "bind_stream"):
            self.engine.bind_stream()
        packed = self._buf("packed", (4, nq, bk), torch.int64)
        self.engine.search3_local(q_float, q_ubinary, nq, bk, self.pos_base, packed)
        if self.world > 1:
            gathered = self._buf("gathered", (self.world, 4, nq, bk), torch.int64)
            self.dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1), group=self.group)
        else:
            gathered = packed.view(1, 4, nq, bk)
        self.engine.merge3(self.world, nq, bk, gathered, k, k * int8_oversample, out)
        return out
