"""NumPy-in / NumPy-out front end of the libvrq encoders and decoders.

These are the functions the reference's private static methods turn into (``VectorDBInt8._quantize_to_int8`` etc.,
see the class modules); they accept a single vector ``[D]`` like the reference or a batch ``[n, D]``.  Arrays live
in host memory here: the call goes through the C ABI's host-buffer path (chunked H2D -> kernel -> D2H).  Callers that
hold device memory (e.g. torch tensors) call the same C entry points with device pointers (``_lib.ptr(tensor)``), which
only enqueue work on the context's stream - see ``sharded.py`` and ``bench.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib as L


def _ctx(ctx: Optional[L.Context]) -> L.Context:
    return ctx if ctx is not None else L.default_context()


def _rows(x: np.ndarray, dtype) -> Tuple[np.ndarray, bool]:
    a = np.ascontiguousarray(x, dtype=dtype)
    if a.ndim == 1:
        return a[None, :], True
    if a.ndim != 2:
        raise ValueError("expected a vector [D] or a batch [n, D]")
    return a, False


def quantize_int8_perdoc(x, want_binary: bool = False, ctx=None):
    """VectorDBInt8._quantize_to_int8 (VectorDBInt8.py:114-126) -> (q int8, min f32, max f32[, ubinary])."""
    a, single = _rows(x, np.float32)
    n, d = a.shape
    q = np.empty((n, d), np.int8)
    lo = np.empty(n, np.float32)
    hi = np.empty(n, np.float32)
    ub = np.empty((n, d // 8), np.uint8) if want_binary else None
    lib, c = L.load(), _ctx(ctx)
    L.check(lib.vrq_quantize_int8_perdoc(c.handle, L.ptr(a), n, d, L.ptr(q), L.ptr(lo), L.ptr(hi), L.ptr(ub)))
    if single:
        return (q[0], lo[0], hi[0]) + ((ub[0],) if want_binary else ())
    return (q, lo, hi) + ((ub,) if want_binary else ())


def quantize_int8_global(x, limit: float, want_binary: bool = False, ctx=None):
    """VectorDBInt8Global._quantize_to_int8 (VectorDBInt8Global.py:130-142)."""
    a, single = _rows(x, np.float32)
    n, d = a.shape
    q = np.empty((n, d), np.int8)
    ub = np.empty((n, d // 8), np.uint8) if want_binary else None
    L.check(L.load().vrq_quantize_int8_global(_ctx(ctx).handle, L.ptr(a), n, d, float(limit), L.ptr(q), L.ptr(ub)))
    if want_binary:
        return (q[0], ub[0]) if single else (q, ub)
    return q[0] if single else q


def quantize_int16_global(x, limit: float, want_binary: bool = False, ctx=None):
    """VectorDBInt16Global._quantize_to_int16 (VectorDBInt16Global.py:130-142)."""
    a, single = _rows(x, np.float32)
    n, d = a.shape
    q = np.empty((n, d), np.int16)
    ub = np.empty((n, d // 8), np.uint8) if want_binary else None
    L.check(L.load().vrq_quantize_int16_global(_ctx(ctx).handle, L.ptr(a), n, d, float(limit), L.ptr(q), L.ptr(ub)))
    if want_binary:
        return (q[0], ub[0]) if single else (q, ub)
    return q[0] if single else q


def quantize_int4(x, want_binary: bool = False, ctx=None):
    """VectorDBInt4._quantize_to_int4 (VectorDBInt4.py:116-154) -> (packed int8[D/2], min f64, max f64[, ubinary])."""
    a, single = _rows(x, np.float32)
    n, d = a.shape
    if d % 2:
        raise L.VrqError(L.ERR_ARG, "embedding_dim must be a multiple of 8")
    q = np.empty((n, d // 2), np.int8)
    lo = np.empty(n, np.float64)
    hi = np.empty(n, np.float64)
    ub = np.empty((n, d // 8), np.uint8) if want_binary else None
    L.check(L.load().vrq_quantize_int4(_ctx(ctx).handle, L.ptr(a), n, d, L.ptr(q), L.ptr(lo), L.ptr(hi), L.ptr(ub)))
    if single:
        return (q[0], float(lo[0]), float(hi[0])) + ((ub[0],) if want_binary else ())
    return (q, lo, hi) + ((ub,) if want_binary else ())


def to_binary(x, ge: bool = False, ctx=None) -> np.ndarray:
    """``_to_binary`` for float32 / int8 / int16 input (VectorDBInt8.py:140-146, CohereVectorDBInt8.py:130-135,
    VectorDBInt16.py:148-157); ``ge`` = CohereVectorDBBinary's ``>=`` threshold."""
    x = np.asarray(x)
    if x.dtype == np.int8:
        a, single = _rows(x, np.int8)
        fn = L.load().vrq_to_binary_i8
    elif x.dtype == np.int16:
        a, single = _rows(x, np.int16)
        fn = L.load().vrq_to_binary_i16
    else:
        a, single = _rows(x, np.float32)
        fn = L.load().vrq_to_binary_f32
    n, d = a.shape
    ub = np.empty((n, d // 8), np.uint8)
    L.check(fn(_ctx(ctx).handle, L.ptr(a), n, d, int(ge), L.ptr(ub)))
    return ub[0] if single else ub


def dequantize_int8_perdoc(q, lo, hi, ctx=None) -> np.ndarray:
    """VectorDBInt8._dequantize_int8 (VectorDBInt8.py:128-138)."""
    a, single = _rows(q, np.int8)
    n, d = a.shape
    lo = np.ascontiguousarray(np.atleast_1d(lo), np.float32)
    hi = np.ascontiguousarray(np.atleast_1d(hi), np.float32)
    out = np.empty((n, d), np.float32)
    L.check(L.load().vrq_dequantize_int8_perdoc(_ctx(ctx).handle, L.ptr(a), n, d, L.ptr(lo), L.ptr(hi), L.ptr(out)))
    return out[0] if single else out


def dequantize_int8_global(q, limit: float, ctx=None) -> np.ndarray:
    """VectorDBInt8Global._dequantize_int8 (VectorDBInt8Global.py:144-152)."""
    a, single = _rows(q, np.int8)
    n, d = a.shape
    out = np.empty((n, d), np.float32)
    L.check(L.load().vrq_dequantize_int8_global(_ctx(ctx).handle, L.ptr(a), n, d, float(limit), L.ptr(out)))
    return out[0] if single else out


def dequantize_int16_global(q, limit: float, ctx=None) -> np.ndarray:
    """VectorDBInt16Global._dequantize_int16 (VectorDBInt16Global.py:144-152)."""
    a, single = _rows(q, np.int16)
    n, d = a.shape
    out = np.empty((n, d), np.float32)
    L.check(L.load().vrq_dequantize_int16_global(_ctx(ctx).handle, L.ptr(a), n, d, float(limit), L.ptr(out)))
    return out[0] if single else out


def dequantize_int4_perdoc(packed, length: int, lo, hi, ctx=None) -> np.ndarray:
    """VectorDBInt4._dequantize_int4 (VectorDBInt4.py:156-184)."""
    a, single = _rows(packed, np.int8)
    n = a.shape[0]
    if a.shape[1] * 2 != length:
        raise L.VrqError(L.ERR_ARG, "length must equal 2 * packed width")
    lo = np.ascontiguousarray(np.atleast_1d(lo), np.float64)
    hi = np.ascontiguousarray(np.atleast_1d(hi), np.float64)
    out = np.empty((n, length), np.float32)
    L.check(L.load().vrq_dequantize_int4_perdoc(_ctx(ctx).handle, L.ptr(a), n, length, L.ptr(lo), L.ptr(hi), L.ptr(out)))
    return out[0] if single else out


def dequantize_int4_global(packed, length: int, limit: float, ctx=None) -> np.ndarray:
    """VectorDBInt4Global._dequantize_int4 (VectorDBInt4Global.py:166-188)."""
    a, single = _rows(packed, np.int8)
    n = a.shape[0]
    if a.shape[1] * 2 != length:
        raise L.VrqError(L.ERR_ARG, "length must equal 2 * packed width")
    out = np.empty((n, length), np.float32)
    L.check(L.load().vrq_dequantize_int4_global(_ctx(ctx).handle, L.ptr(a), n, length, float(limit), L.ptr(out)))
    return out[0] if single else out


def rescore_binary(codes, pos, q_float, ctx=None) -> np.ndarray:
    """Phase II (CohereEnhancedVectorDB.py:283-293) for candidates ``pos[nq, m]`` of ``codes[n, D/8]`` -> f64[nq, m]."""
    codes = np.ascontiguousarray(codes, np.uint8)
    pos = np.ascontiguousarray(pos, np.int64)
    qf = np.ascontiguousarray(q_float, np.float32)
    if qf.ndim == 1:
        qf, pos = qf[None], pos.reshape(1, -1)
    nq, m = pos.shape
    out = np.empty((nq, m), np.float64)
    L.check(L.load().vrq_rescore_binary(_ctx(ctx).handle, L.ptr(codes), codes.shape[0], codes.shape[1] * 8, L.ptr(pos), nq, m,
                                        L.ptr(qf), L.ptr(out)))
    return out


def rescore_int8cos(rows, pos, q_float, ctx=None) -> np.ndarray:
    """Phase III (CohereEnhancedVectorDB.py:302-318) for candidates ``pos[nq, m]`` of ``rows[n, D]`` int8 -> f64[nq, m]."""
    rows = np.ascontiguousarray(rows, np.int8)
    pos = np.ascontiguousarray(pos, np.int64)
    qf = np.ascontiguousarray(q_float, np.float32)
    if qf.ndim == 1:
        qf, pos = qf[None], pos.reshape(1, -1)
    nq, m = pos.shape
    out = np.empty((nq, m), np.float64)
    L.check(L.load().vrq_rescore_int8cos(_ctx(ctx).handle, L.ptr(rows), rows.shape[0], rows.shape[1], L.ptr(pos), nq, m,
                                         L.ptr(qf), L.ptr(out)))
    return out


def synth_f32(seed: int, row0: int, nrows: int, d: int = 1024, row_scale: bool = False, ctx=None) -> np.ndarray:
    """Synthetic float32 embeddings (stand-in for the Ollama / Cohere services)."""
    out = np.empty((nrows, d), np.float32)
    L.check(L.load().vrq_synth_f32(_ctx(ctx).handle, C.c_uint64(seed), row0, nrows, d, int(row_scale), L.ptr(out)))
    return out


def synth_codes_int8(seed: int, row0: int, nrows: int, d: int = 1024, ctx=None):
    """Cohere-like (ubinary, int8) pair of the synthetic rows."""
    codes = np.empty((nrows, d // 8), np.uint8)
    i8 = np.empty((nrows, d), np.int8)
    L.check(L.load().vrq_synth_codes_int8(_ctx(ctx).handle, C.c_uint64(seed), row0, nrows, d, L.ptr(codes), L.ptr(i8)))
    return codes, i8
