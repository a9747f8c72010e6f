"""Fused encoders (encode.cu) over PROF_ROWS x 1024 float32 rows resident in HBM: ms and GB/s per codec (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vectorragquantization_b200 as V  # noqa: E402
from vectorragquantization_b200 import _lib as L  # noqa: E402

D = 1024
ctx = V.Context(0)
lib = L.load()
dev = torch.device("cuda", 0)
ctx.set_stream(0)
n = int(os.environ.get("PROF_ROWS", 4_000_000))
reps = int(os.environ.get("PROF_ITERS", 5))
only = os.environ.get("PROF_CODECS", "")
x = torch.empty((n, D), dtype=torch.float32, device=dev)
L.check(lib.vrq_synth_f32(ctx.handle, 7, 0, n, D, 1, L.ptr(x)))
ub = torch.empty((n, D // 8), dtype=torch.uint8, device=dev)
q8 = torch.empty((n, D), dtype=torch.int8, device=dev)
q16 = torch.empty((n, D), dtype=torch.int16, device=dev)
lo = torch.empty((n,), dtype=torch.float64, device=dev)
hi = torch.empty((n,), dtype=torch.float64, device=dev)
h = ctx.handle
cases = {
    "int8_global+ubinary": (lambda: L.check(lib.vrq_quantize_int8_global(h, L.ptr(x), n, D, 0.3, L.ptr(q8), L.ptr(ub))), 4096 + 1024 + 128),
    "int16_global+ubinary": (lambda: L.check(lib.vrq_quantize_int16_global(h, L.ptr(x), n, D, 1.0, L.ptr(q16), L.ptr(ub))), 4096 + 2048 + 128),
    "int4+ubinary": (lambda: L.check(lib.vrq_quantize_int4(h, L.ptr(x), n, D, L.ptr(q8), L.ptr(lo), L.ptr(hi), L.ptr(ub))), 4096 + 512 + 16 + 128),
    "int8_perdoc+ubinary": (lambda: L.check(lib.vrq_quantize_int8_perdoc(h, L.ptr(x), n, D, L.ptr(q8), L.ptr(lo), L.ptr(hi), L.ptr(ub))), 4096 + 1024 + 8 + 128),
    "ubinary_only": (lambda: L.check(lib.vrq_to_binary_f32(h, L.ptr(x), n, D, 0, L.ptr(ub))), 4096 + 128),
}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# PROF_CONFIGS="ring,warps,stages;..." sweeps the launch shape in one process (the library reads the knobs per launch)
configs = [c.split(",") for c in os.environ.get("PROF_CONFIGS", "").split(";") if c] or [None]
for cfg in configs:
    if cfg and cfg[0] == "default":
        for key in ("VRQ_ENCODE_RING", "VRQ_ENCODE_WARPS", "VRQ_ENCODE_STAGES"):
            os.environ.pop(key, None)
        print("--- library defaults", flush=True)
    elif cfg:
        os.environ["VRQ_ENCODE_RING"], os.environ["VRQ_ENCODE_WARPS"], os.environ["VRQ_ENCODE_STAGES"] = cfg
        print(f"--- ring={cfg[0]} warps/block={cfg[1]} stages={cfg[2]}", flush=True)
    for name, (fn, bpr) in cases.items():
        if only and name not in only.split(","):
            continue
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name:24s} {ms:8.3f} ms  {n * bpr / ms / 1e6:8.1f} GB/s  ({n} rows)", flush=True)
