"""ctypes binding of oracle/liboracle.so (the C half of the CPU oracle).  TEST INFRASTRUCTURE ONLY.

Build with ``make -C oracle`` (``__graft_entry__.build()`` does it).  See oracle_c.c for the reference
file:line each function restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return os.path.join(_HERE, "liboracle.so")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.vrqo_pairwise_sum_f32.restype = C.c_float
        _LIB.vrqo_num_threads.restype = C.c_int
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().vrqo_num_threads())


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs want every core this process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().vrqo_set_num_threads(C.c_int(n))
    return num_threads()


def pairwise_sum_f32(a: np.ndarray) -> np.float32:
    a = np.ascontiguousarray(a, np.float32)
    return np.float32(lib().vrqo_pairwise_sum_f32(_p(a), C.c_int64(a.shape[0])))


def to_binary_f32(x: np.ndarray, ge: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    out = np.empty((n, (d + 7) // 8), np.uint8)
    lib().vrqo_to_binary_f32(_p(x), C.c_int64(n), C.c_int(d), C.c_int(int(ge)), _p(out))
    return out


def to_binary_int(x: np.ndarray, ge: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x)
    n, d = x.shape
    out = np.empty((n, (d + 7) // 8), np.uint8)
    fn = {np.dtype(np.int8): lib().vrqo_to_binary_i8, np.dtype(np.int16): lib().vrqo_to_binary_i16}[x.dtype]
    fn(_p(x), C.c_int64(n), C.c_int(d), C.c_int(int(ge)), _p(out))
    return out


def quantize_int8_perdoc(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    q = np.empty((n, d), np.int8)
    lo = np.empty(n, np.float32)
    hi = np.empty(n, np.float32)
    lib().vrqo_quantize_int8_perdoc(_p(x), C.c_int64(n), C.c_int(d), _p(q), _p(lo), _p(hi))
    return q, lo, hi


def quantize_int8_global(x: np.ndarray, limit: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    q = np.empty((n, d), np.int8)
    lib().vrqo_quantize_int8_global(_p(x), C.c_int64(n), C.c_int(d), C.c_double(limit), _p(q))
    return q


def quantize_int16_global(x: np.ndarray, limit: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    q = np.empty((n, d), np.int16)
    lib().vrqo_quantize_int16_global(_p(x), C.c_int64(n), C.c_int(d), C.c_double(limit), _p(q))
    return q


def quantize_int4(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    q = np.empty((n, (d + 1) // 2), np.int8)
    lo = np.empty(n, np.float64)
    hi = np.empty(n, np.float64)
    lib().vrqo_quantize_int4(_p(x), C.c_int64(n), C.c_int(d), _p(q), _p(lo), _p(hi))
    return q, lo, hi


def hamming_topk(codes: np.ndarray, q: np.ndarray, k: int, pos_base: int = 0, nthreads: int = 0):
    codes = np.ascontiguousarray(codes, np.uint8)
    q = np.ascontiguousarray(q, np.uint8)
    if q.ndim == 1:
        q = q[None]
    n, cb = codes.shape if codes.ndim == 2 else (0, q.shape[1])
    nq = q.shape[0]
    dist = np.empty((nq, k), np.int32)
    pos = np.empty((nq, k), np.int64)
    lib().vrqo_hamming_topk(_p(codes), C.c_int64(n), C.c_int(cb), _p(q), C.c_int(nq), C.c_int(k),
                            C.c_int64(pos_base), _p(dist), _p(pos), C.c_int(nthreads))
    return dist, pos


def rescore_binary(q_float: np.ndarray, cand_codes: np.ndarray) -> np.ndarray:
    q_float = np.ascontiguousarray(q_float, np.float32)
    cand_codes = np.ascontiguousarray(cand_codes, np.uint8)
    m = cand_codes.shape[0]
    out = np.empty(m, np.float64)
    lib().vrqo_rescore_binary(_p(q_float), C.c_int(q_float.shape[0]), _p(cand_codes), C.c_int64(m), _p(out))
    return out


def int8_sumsq(rows: np.ndarray) -> np.ndarray:
    rows = np.ascontiguousarray(rows, np.int8)
    m, d = rows.shape
    out = np.empty(m, np.int64)
    lib().vrqo_int8_sumsq(_p(rows), C.c_int64(m), C.c_int(d), _p(out))
    return out


def synth_f32(seed: int, row0: int, nrows: int, d: int = 1024, row_scale: bool = False) -> np.ndarray:
    out = np.empty((nrows, d), np.float32)
    lib().vrqo_synth_f32(C.c_uint64(seed), C.c_int64(row0), C.c_int64(nrows), C.c_int(d),
                         C.c_int(int(row_scale)), _p(out))
    return out


def synth_codes_int8(seed: int, row0: int, nrows: int, d: int = 1024, want_codes: bool = True,
                     want_int8: bool = True):
    codes = np.empty((nrows, d // 8), np.uint8) if want_codes else None
    i8 = np.empty((nrows, d), np.int8) if want_int8 else None
    lib().vrqo_synth_codes_int8(C.c_uint64(seed), C.c_int64(row0), C.c_int64(nrows), C.c_int(d),
                                _p(codes) if want_codes else None, _p(i8) if want_int8 else None)
    return codes, i8
