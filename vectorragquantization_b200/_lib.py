"""ctypes binding of libvrq.so (the C ABI declared in include/vrq.h).

There is deliberately NO fallback: if the shared library is missing, was built for another architecture, or no
CUDA device is present, importing works (so CPU-only tooling can introspect the package) but the first call
raises ``VrqError`` - nothing here ever routes to NumPy or to the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VRQ_LIBVRQ") or os.path.join(_HERE, "libvrq.so")  # VRQ_LIBVRQ: A/B runs against another build

ERR_ARG, ERR_UNSUPPORTED, ERR_IO, ERR_STATE, ERR_NOMEM = -1, -2, -3, -4, -5

PAYLOAD_NONE, PAYLOAD_INT8_RAW, PAYLOAD_INT8_PERDOC, PAYLOAD_INT8_GLOBAL = 0, 1, 2, 3
PAYLOAD_INT16_GLOBAL, PAYLOAD_INT4_PERDOC, PAYLOAD_INT4_GLOBAL, PAYLOAD_F32 = 4, 5, 6, 7
PAYLOAD_CODES_PM1 = 8

ROWS_CODES, ROWS_IDS, ROWS_PAYLOAD, ROWS_AUX = 0, 1, 2, 3

KEY_POS_BITS = 40
KEY_NONE = 0xFFFFFFFFFFFFFFFF
MAX_K = 16384


class VrqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvrq error {code}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None
_lock = threading.Lock()

_vp, _i64, _i32, _dbl, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64

# name -> (restype, argtypes); kept in one table so tests can check it against include/vrq.h
SIGNATURES = {
    "vrq_version": (_i32, []),
    "vrq_last_error": (C.c_char_p, []),
    "vrq_ctx_create": (_i32, [_i32, C.POINTER(_vp)]),
    "vrq_ctx_destroy": (_i32, [_vp]),
    "vrq_ctx_set_stream": (_i32, [_vp, _vp]),
    "vrq_ctx_reset_stream": (_i32, [_vp]),
    "vrq_ctx_sync": (_i32, [_vp]),
    "vrq_ctx_launch_count": (_i64, [_vp]),
    "vrq_ctx_device": (_i32, [_vp]),
    "vrq_ctx_enable_timing": (_i32, [_vp, _i32]),
    "vrq_ctx_timing_ms": (_dbl, [_vp, C.c_char_p, C.POINTER(_i64)]),
    "vrq_quantize_int8_perdoc": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "vrq_quantize_int8_global": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp, _vp]),
    "vrq_quantize_int16_global": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp, _vp]),
    "vrq_quantize_int4": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "vrq_to_binary_f32": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "vrq_to_binary_i8": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "vrq_to_binary_i16": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "vrq_dequantize_int8_perdoc": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "vrq_dequantize_int8_global": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp]),
    "vrq_dequantize_int16_global": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp]),
    "vrq_dequantize_int4_perdoc": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "vrq_dequantize_int4_global": (_i32, [_vp, _vp, _i64, _i32, _dbl, _vp]),
    "vrq_index_create": (_i32, [_vp, _i32, C.POINTER(_vp)]),
    "vrq_index_free": (_i32, [_vp]),
    "vrq_index_ntotal": (_i64, [_vp]),
    "vrq_index_d": (_i32, [_vp]),
    "vrq_index_reserve": (_i32, [_vp, _i64]),
    "vrq_index_device_ptrs": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "vrq_index_set_payload": (_i32, [_vp, _i32, _dbl]),
    "vrq_index_payload_kind": (_i32, [_vp]),
    "vrq_index_add_with_ids": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "vrq_index_search": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "vrq_index_distances": (_i32, [_vp, _i64, _vp, _vp]),
    "vrq_index_reconstruct": (_i32, [_vp, _i64, _vp]),
    "vrq_index_remove_ids": (_i64, [_vp, _i64, _vp]),
    "vrq_index_write": (_i32, [_vp, C.c_char_p]),
    "vrq_index_read": (_i32, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "vrq_index_write_payload": (_i32, [_vp, C.c_char_p]),
    "vrq_index_read_payload": (_i32, [_vp, C.c_char_p]),
    "vrq_index_attach_payload": (_i32, [_vp, _i32, _dbl]),
    "vrq_index_read_rows": (_i32, [_vp, _i32, _i64, _i64, _vp]),
    "vrq_index_write_rows": (_i32, [_vp, _i32, _i64, _i64, _vp]),
    "vrq_index_get_payload": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "vrq_index_position_of": (_i64, [_vp, _i64]),
    "vrq_index_search_ip": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "vrq_index_write_float": (_i32, [_vp, C.c_char_p]),
    "vrq_index_read_float": (_i32, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "vrq_index_search3": (_i32, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vrq_index_search2": (_i32, [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vrq_index_search3_local": (_i32, [_vp, _i64, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "vrq_merge3": (_i32, [_vp, _i32, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vrq_nccl_unique_id": (_i32, [_vp]),
    "vrq_nccl_init_rank": (_i32, [_i32, _i32, _vp, _i32, C.POINTER(_vp)]),
    "vrq_nccl_init_all": (_i32, [_i32, _vp, _vp]),
    "vrq_nccl_destroy": (_i32, [_vp]),
    "vrq_ctx_set_nccl": (_i32, [_vp, _vp, _i32, _i32]),
    "vrq_index_search3_sharded": (_i32, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "vrq_search3_sharded_group": (_i32, [_i32, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "vrq_rescore_binary": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp]),
    "vrq_rescore_int8cos": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp]),
    "vrq_synth_f32": (_i32, [_vp, _u64, _i64, _i64, _i32, _i32, _vp]),
    "vrq_synth_codes_int8": (_i32, [_vp, _u64, _i64, _i64, _i32, _vp, _vp]),
    "vrq_index_add_synthetic": (_i32, [_vp, _u64, _i64, _i64, _i64]),
    "vrq_index_set_synthetic_payload": (_i32, [_vp, _u64, _i64]),
}


def load() -> C.CDLL:
    """dlopen libvrq.so and attach the prototypes.  Raises VrqError if the library has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VrqError(ERR_STATE, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                                      f"g.build()'` or `make -C vectorragquantization_b200/csrc` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().vrq_last_error().decode("utf-8", "replace")


def check(rc: int) -> int:
    if rc != 0:
        raise VrqError(rc, last_error())
    return rc


def ptr(a) -> Optional[int]:
    """Address of a NumPy array / torch tensor / raw int address (None stays None)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host memory): plumbing only
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    return int(a)


class Context:
    """One per GPU (vrq_ctx)."""

    def __init__(self, device: int = 0):
        lib = load()
        h = _vp()
        check(lib.vrq_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @property
    def handle(self):
        if self._h is None:
            raise VrqError(ERR_STATE, "context was destroyed")
        return self._h

    def set_stream(self, cuda_stream: Optional[int]):
        """Enqueue on this cudaStream_t (0 / None = the legacy default stream, i.e. torch's default stream)."""
        check(load().vrq_ctx_set_stream(self.handle, _vp(cuda_stream or 0)))

    def reset_stream(self):
        check(load().vrq_ctx_reset_stream(self.handle))

    def sync(self):
        check(load().vrq_ctx_sync(self.handle))

    def launch_count(self) -> int:
        return int(load().vrq_ctx_launch_count(self.handle))

    def enable_timing(self, on: bool = True):
        check(load().vrq_ctx_enable_timing(self.handle, int(on)))

    def timing_ms(self, which: str):
        n = _i64(0)
        ms = load().vrq_ctx_timing_ms(self.handle, which.encode(), C.byref(n))
        return float(ms), int(n.value)

    def close(self):
        if getattr(self, "_h", None) is not None:
            load().vrq_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: Optional[int] = None) -> Context:
    """Process-wide context for ``device`` (default: LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _lock:
        pass
    c = _default_ctx.get(device)
    if c is None:
        c = Context(device)
        _default_ctx[device] = c
    return c
