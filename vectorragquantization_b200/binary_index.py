"""faiss-shaped front end of the device-resident binary index (vrq_index_* in include/vrq.h).

The reference classes only ever touch this slice of faiss:

    faiss.IndexBinaryIDMap2(faiss.IndexBinaryFlat(d))      CohereEnhancedVectorDB.py:126, VectorDBInt8.py:69
    index.add_with_ids(u8[n, d/8], i64[n])                 :217 / :175
    index.search(u8[nq, d/8], k) -> (i32[nq,k], i64[nq,k]) :268 / :218
    index.reconstruct(id) -> u8[d/8]                       :286
    index.remove_ids(i64[m]) -> n_removed                  :334
    index.ntotal                                           :247, :267
    faiss.read_index_binary / write_index_binary           :123 / :346

so ``BinaryIndex`` answers exactly those, with the same argument meaning, and ``read_index_binary`` /
``write_index_binary`` read and write the same bytes faiss does ("IBM2" wrapping "IBxF").
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib as L


class IndexBinaryFlat:
    """Constructor token mirroring ``faiss.IndexBinaryFlat(d)``; only meaningful wrapped in IndexBinaryIDMap2."""

    def __init__(self, d: int):
        self.d = int(d)


class BinaryIndex:
    def __init__(self, d, ctx: Optional[L.Context] = None, payload_kind: int = L.PAYLOAD_NONE, global_limit: float = 0.0,
                 _handle=None):
        if isinstance(d, IndexBinaryFlat):
            d = d.d
        self.ctx = ctx if ctx is not None else L.default_context()
        self._lib = L.load()
        if _handle is None:
            h = C.c_void_p()
            L.check(self._lib.vrq_index_create(self.ctx.handle, int(d), C.byref(h)))
            self._h = h
            if payload_kind != L.PAYLOAD_NONE:
                L.check(self._lib.vrq_index_set_payload(self._h, int(payload_kind), float(global_limit)))
        else:
            self._h = _handle
        self.d = int(self._lib.vrq_index_d(self._h))
        self.code_size = self.d // 8

    # ---- faiss surface ---------------------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(self._lib.vrq_index_ntotal(self._h))

    @property
    def payload_kind(self) -> int:
        return int(self._lib.vrq_index_payload_kind(self._h))

    def add_with_ids(self, codes, ids, payload=None, aux=None) -> None:
        codes = np.ascontiguousarray(codes, np.uint8)
        if codes.ndim == 1:
            codes = codes[None]
        ids = np.ascontiguousarray(ids, np.int64).reshape(-1)
        if codes.shape[1] != self.code_size or codes.shape[0] != ids.shape[0]:
            raise ValueError("codes must be uint8[n, d/8] and ids int64[n]")
        if payload is not None:
            payload = np.ascontiguousarray(payload)
        if aux is not None:
            aux = np.ascontiguousarray(aux)
        L.check(self._lib.vrq_index_add_with_ids(self._h, codes.shape[0], L.ptr(codes), L.ptr(ids), L.ptr(payload), L.ptr(aux)))

    def search(self, q, k: int) -> Tuple[np.ndarray, np.ndarray]:
        q = np.ascontiguousarray(q, np.uint8)
        if q.ndim == 1:
            q = q[None]
        if q.shape[1] != self.code_size:
            raise ValueError("query codes must be uint8[nq, d/8]")
        nq = q.shape[0]
        dist = np.empty((nq, k), np.int32)
        labels = np.empty((nq, k), np.int64)
        L.check(self._lib.vrq_index_search(self._h, nq, L.ptr(q), int(k), L.ptr(dist), L.ptr(labels)))
        return dist, labels

    def distances(self, q) -> np.ndarray:
        """Every Hamming distance int32[nq, ntotal] from the tensor-core scan's accumulators (d == 1024 only)."""
        q = np.ascontiguousarray(q, np.uint8)
        if q.ndim == 1:
            q = q[None]
        if q.shape[1] != self.code_size:
            raise ValueError("query codes must be uint8[nq, d/8]")
        dist = np.empty((q.shape[0], self.ntotal), np.int32)
        L.check(self._lib.vrq_index_distances(self._h, q.shape[0], L.ptr(q), L.ptr(dist)))
        return dist

    def reconstruct(self, doc_id: int) -> np.ndarray:
        out = np.empty(self.code_size, np.uint8)
        L.check(self._lib.vrq_index_reconstruct(self._h, int(doc_id), L.ptr(out)))
        return out

    def remove_ids(self, ids) -> int:
        ids = np.ascontiguousarray(ids, np.int64).reshape(-1)
        r = int(self._lib.vrq_index_remove_ids(self._h, ids.shape[0], L.ptr(ids)))
        if r < 0:
            raise L.VrqError(r, L.last_error())
        return r

    # ---- beyond faiss: what replaces the Python rescoring loops ---------------------------------------------
    def reserve(self, n: int) -> None:
        L.check(self._lib.vrq_index_reserve(self._h, int(n)))

    def device_ptrs(self):
        """(codes, ids, payload, aux) device addresses (ints; 0 where absent) - for the stand-alone kernels."""
        ptrs = [C.c_void_p() for _ in range(4)]
        L.check(self._lib.vrq_index_device_ptrs(self._h, *[C.byref(p) for p in ptrs]))
        return tuple(p.value or 0 for p in ptrs)

    def position_of(self, doc_id: int) -> int:
        return int(self._lib.vrq_index_position_of(self._h, int(doc_id)))

    def get_payload(self, positions, row_dtype, row_len: int, aux_dtype=None):
        positions = np.ascontiguousarray(positions, np.int64).reshape(-1)
        m = positions.shape[0]
        pay = np.empty((m, row_len), row_dtype)
        aux = np.empty((m, 2), aux_dtype) if aux_dtype is not None else None
        L.check(self._lib.vrq_index_get_payload(self._h, m, L.ptr(positions), L.ptr(pay), L.ptr(aux)))
        return pay, aux

    def add_synthetic(self, seed: int, row0: int, nrows: int, id0: int) -> None:
        L.check(self._lib.vrq_index_add_synthetic(self._h, C.c_uint64(seed), int(row0), int(nrows), int(id0)))

    # ---- resident rows and the payload sidecar file ------------------------------------------------------------------
    def payload_layout(self):
        """(row dtype, row length, aux dtype or None) of this index's payload kind."""
        d = self.d
        return {L.PAYLOAD_INT8_RAW: (np.int8, d, None), L.PAYLOAD_INT8_PERDOC: (np.int8, d, np.float32),
                L.PAYLOAD_INT8_GLOBAL: (np.int8, d, None), L.PAYLOAD_INT16_GLOBAL: (np.int16, d, None),
                L.PAYLOAD_INT4_PERDOC: (np.int8, d // 2, np.float64), L.PAYLOAD_INT4_GLOBAL: (np.int8, d // 2, None),
                L.PAYLOAD_F32: (np.float32, d, None)}[self.payload_kind]

    def read_rows(self, which: int, offset: int, count: int) -> np.ndarray:
        """``count`` consecutive rows of the resident codes / ids / payload / aux arrays (L.ROWS_*), copied to the host."""
        if which == L.ROWS_CODES:
            out = np.empty((count, self.code_size), np.uint8)
        elif which == L.ROWS_IDS:
            out = np.empty(count, np.int64)
        else:
            dt, ln, adt = self.payload_layout()
            out = np.empty((count, ln), dt) if which == L.ROWS_PAYLOAD else np.empty((count, 2), adt)
        L.check(self._lib.vrq_index_read_rows(self._h, int(which), int(offset), int(count), L.ptr(out)))
        return out

    def write_rows(self, which: int, offset: int, rows) -> None:
        rows = np.ascontiguousarray(rows)
        L.check(self._lib.vrq_index_write_rows(self._h, int(which), int(offset), int(rows.shape[0]), L.ptr(rows)))

    def attach_payload(self, kind: int, global_limit: float = 0.0) -> None:
        L.check(self._lib.vrq_index_attach_payload(self._h, int(kind), float(global_limit)))

    def write_payload(self, path: str) -> None:
        """Stream the payload (+ aux) rows to a sidecar file next to index.bin (vrq_index_write_payload)."""
        L.check(self._lib.vrq_index_write_payload(self._h, str(path).encode()))

    def read_payload(self, path: str) -> None:
        L.check(self._lib.vrq_index_read_payload(self._h, str(path).encode()))

    def set_synthetic_payload(self, seed: int, row0: int) -> None:
        """Benchmark-only: INT8_RAW rows are regenerated on demand instead of stored (vrq_index_set_synthetic_payload)."""
        L.check(self._lib.vrq_index_set_synthetic_payload(self._h, C.c_uint64(seed), int(row0)))

    def search3(self, q_float, q_ubinary, k: int, binary_oversample: int = 10, int8_oversample: int = 3):
        """Phases I-III of CohereEnhancedVectorDB.search (:267-322) for a batch of queries.
        Returns (labels i64[nq,k], hamming i32, score_binary f64, score_cosine f64, count i32[nq])."""
        qf = np.ascontiguousarray(q_float, np.float32)
        qb = np.ascontiguousarray(q_ubinary, np.uint8)
        if qf.ndim == 1:
            qf, qb = qf[None], qb.reshape(1, -1)
        nq = qf.shape[0]
        if qf.shape[1] != self.d or qb.shape != (nq, self.code_size):
            raise ValueError("q_float must be float32[nq, d] and q_ubinary uint8[nq, d/8]")
        labels = np.empty((nq, k), np.int64)
        ham = np.empty((nq, k), np.int32)
        sb = np.empty((nq, k), np.float64)
        sc = np.empty((nq, k), np.float64)
        cnt = np.empty(nq, np.int32)
        L.check(self._lib.vrq_index_search3(self._h, nq, L.ptr(qf), L.ptr(qb), int(k), int(binary_oversample),
                                            int(int8_oversample), L.ptr(labels), L.ptr(ham), L.ptr(sb), L.ptr(sc), L.ptr(cnt)))
        return labels, ham, sb, sc, cnt

    def search2(self, q_float, q_ubinary, k: int, binary_oversample: int = 10):
        """The VectorDB* 2-phase search (VectorDBInt8.py:213-242) for a batch of queries.
        ``q_ubinary`` None: the library derives it from ``q_float`` on the device (packbits(q > mean(q))).
        Returns (labels i64[nq,k], score f32[nq,k], count i32[nq])."""
        qf = np.ascontiguousarray(q_float, np.float32)
        qb = None if q_ubinary is None else np.ascontiguousarray(q_ubinary, np.uint8)
        if qf.ndim == 1:
            qf = qf[None]
            qb = None if qb is None else qb.reshape(1, -1)
        nq = qf.shape[0]
        labels = np.empty((nq, k), np.int64)
        score = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.int32)
        L.check(self._lib.vrq_index_search2(self._h, nq, L.ptr(qf), None if qb is None else L.ptr(qb), int(k), int(binary_oversample), L.ptr(labels),
                                            L.ptr(score), L.ptr(cnt)))
        return labels, score, cnt

    def search_ip(self, q_float, k: int):
        """``faiss.IndexFlatIP.search`` on the float32 payload rows (CohereVectorDBFloat.py:156): (scores f32[nq,k] descending,
        labels i64[nq,k]); padded with (-inf, -1)."""
        qf = np.ascontiguousarray(q_float, np.float32)
        if qf.ndim == 1:
            qf = qf[None]
        if qf.shape[1] != self.d:
            raise ValueError("q_float must be float32[nq, d]")
        scores = np.empty((qf.shape[0], k), np.float32)
        labels = np.empty((qf.shape[0], k), np.int64)
        L.check(self._lib.vrq_index_search_ip(self._h, qf.shape[0], L.ptr(qf), int(k), L.ptr(scores), L.ptr(labels)))
        return scores, labels

    def add_float_rows(self, x, ids) -> None:
        """``IndexIDMap(IndexFlatIP).add_with_ids`` (CohereVectorDBFloat.py:133): float32 rows only, no binary codes."""
        x = np.ascontiguousarray(x, np.float32)
        ids = np.ascontiguousarray(ids, np.int64).reshape(-1)
        if x.ndim != 2 or x.shape != (ids.shape[0], self.d):
            raise ValueError("x must be float32[n, d] and ids int64[n]")
        L.check(self._lib.vrq_index_add_with_ids(self._h, x.shape[0], None, L.ptr(ids), L.ptr(x), None))

    def search3_local_into(self, q_float_dev, q_ubin_dev, nq: int, binary_k: int, pos_base: int, keys, labels, sbin, scos):
        """Device-pointer, asynchronous: per-shard candidates for the multi-GPU path (sharded.py)."""
        L.check(self._lib.vrq_index_search3_local(self._h, int(nq), L.ptr(q_float_dev), L.ptr(q_ubin_dev), int(binary_k),
                                                  int(pos_base), L.ptr(keys), L.ptr(labels), L.ptr(sbin), L.ptr(scos)))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.vrq_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def IndexBinaryIDMap2(inner, **kw) -> BinaryIndex:
    """``faiss.IndexBinaryIDMap2(faiss.IndexBinaryFlat(d))`` (CohereEnhancedVectorDB.py:126)."""
    return BinaryIndex(inner, **kw)


def write_index_binary(index: BinaryIndex, path: str) -> None:
    """``faiss.write_index_binary`` (CohereEnhancedVectorDB.py:346): byte-compatible file."""
    L.check(index._lib.vrq_index_write(index._h, str(path).encode()))


def write_index_float(index: BinaryIndex, path: str) -> None:
    """``faiss.write_index`` of ``IndexIDMap(IndexFlatIP)`` (CohereVectorDBFloat.py:184): byte-compatible "IxMp"/"IxFI" file."""
    L.check(index._lib.vrq_index_write_float(index._h, str(path).encode()))


def read_index_float(path: str, ctx: Optional[L.Context] = None) -> BinaryIndex:
    """``faiss.read_index`` of that file (CohereVectorDBFloat.py:58)."""
    ctx = ctx if ctx is not None else L.default_context()
    h = C.c_void_p()
    L.check(L.load().vrq_index_read_float(ctx.handle, str(path).encode(), C.byref(h)))
    return BinaryIndex(0, ctx=ctx, _handle=h)


def read_index_binary(path: str, ctx: Optional[L.Context] = None) -> BinaryIndex:
    """``faiss.read_index_binary`` (CohereEnhancedVectorDB.py:123)."""
    ctx = ctx if ctx is not None else L.default_context()
    h = C.c_void_p()
    L.check(L.load().vrq_index_read(ctx.handle, str(path).encode(), C.byref(h)))
    return BinaryIndex(0, ctx=ctx, _handle=h)
