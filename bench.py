#!/usr/bin/env python
"""bench.py - headline benchmark of the quantise + 3-phase search hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one its `metric` is quoted on): CohereEnhancedVectorDB 3-phase search -
ubinary Hamming top-1000 over 100 M x 1024-bit codes, float.(+-1 bits) rescoring, int8 "cosine" rescoring - for a
batch of 1024 queries; k=100, binary_oversample=10, int8_oversample=3.  One step = one 1024-query batch.  Synthetic
Cohere-like data from the counter-based generator (no network).  With N GPUs every rank holds its own 100 M-row
shard (weak scaling: the database is N x 100 M rows), each query is answered exactly over the whole database
(per-shard candidates -> one NCCL all-gather -> device merge), and `value` counts queries x (database rows / 100 M)
per second, i.e. it is plain QPS@100M at N=1 and total pair throughput in the same unit for N>1.

The JSON line carries, besides the base contract: `roofline` (dominant kernel = the dense pass of the batched Hamming
scan, a tcgen05 e2m1 contraction bound by the tensor pipe, with its HBM figures alongside), `roofline_scan_stream` /
`roofline_encode` (the two HBM-bound kernels the metric names: scan at <=2 queries per pass, fused int8 encode),
`cpu_baseline`, `e2e`, `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 1024
NQ = 1024
K, BO, IO = 100, 10, 3
N_PER_GPU = 100_000_000
DB_SEED, Q_SEED = 1, 2
POPC_PER_CLK_PER_SM = 16.0  # measured: profiles/microbench/popc_bench_r01.txt
LOP3_PER_CLK_PER_SM = 64.0  # measured: same file
LOP3_PER_PAIR = 64.0        # 32 XOR + 16 carry-save adders x 2 LOP3 (scan.cu: hamming128_csa); + 16 POPC on the XU pipe
METRIC = "3-phase search QPS @100Mx1024-d"
UNIT = "queries/s per 100M codes"


INT8_DENSE_NOMINAL_TOPS = 4500.0  # B200 dense int8 / fp8 tensor rate (B200_PROFILING.md, nominal table)
FP4_DENSE_NOMINAL_TOPS = 9000.0   # B200 dense fp4 tensor rate (same table)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bf16_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["bf16_tflops"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_search3_sample(n_s: int, nq_s: int, step: int = 0):
    """The reference's CPU path on a bounded sample of the workload: faiss-equivalent Hamming scan (C/OpenMP over
    queries, the way faiss parallelises IndexBinaryFlat.search) on the first n_s rows, then the reference's literal
    per-candidate NumPy loops for phases II and III (CohereEnhancedVectorDB.py:283-322).  Returns seconds."""
    from oracle import oracle_c as oc
    from oracle import vrq_oracle as o
    if not hasattr(cpu_search3_sample, "cache") or cpu_search3_sample.cache[0] != n_s:
        codes, _ = oc.synth_codes_int8(DB_SEED, 0, n_s, want_int8=False)
        cpu_search3_sample.cache = (n_s, codes)
    codes = cpu_search3_sample.cache[1]
    qf = oc.synth_f32(Q_SEED, step * NQ, nq_s)
    qb = o.synth_ubinary_from_f32(qf)
    t0 = time.perf_counter()
    bk = K * BO
    dist, pos = oc.hamming_topk(codes, qb, bk)
    for qi in range(nq_s):
        hits = [(int(d), int(p)) for d, p in zip(dist[qi], pos[qi]) if p != -1]
        hits.sort(key=lambda h: h[0])
        sb = []
        for _, p in hits:  # Phase II, reference style
            u = np.unpackbits(codes[p], axis=-1).astype(np.int32)
            u = 2 * u - 1
            sb.append(float(qf[qi].dot(u)))
        order = sorted(range(len(hits)), key=lambda i: sb[i], reverse=True)[:K * IO]
        sc = []
        for i in order:  # Phase III, reference style (the int8 row is regenerated instead of a RocksDB get)
            d8 = oc.synth_codes_int8(DB_SEED, hits[i][1], 1, want_codes=False)[1][0]
            nrm = np.linalg.norm(d8)
            sc.append(-np.inf if nrm == 0 else float(qf[qi].dot(d8)) / nrm)
        sorted(range(len(order)), key=lambda i: sc[i], reverse=True)[:K]
    return time.perf_counter() - t0


def cpu_baseline(target_s: float = 12.0):
    from oracle import oracle_c as oc
    cores = oc.use_all_cores()
    n_s, nq_s = 2_000_000, max(cores, 8)
    t = cpu_search3_sample(n_s, nq_s)
    while t < target_s / 3 and nq_s < 512:
        nq_s *= 2
        t = cpu_search3_sample(n_s, nq_s)
    value = nq_s / t * (n_s / N_PER_GPU)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nq_s} queries x first {n_s} of the 100M synthetic codes (C/OpenMP scan over queries + the reference's "
                      f"per-candidate NumPy loops for phases II/III), {t:.2f} s; scaled by {n_s}/{N_PER_GPU} to the 100M workload"}, (n_s, nq_s)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c as oc
    cores = oc.use_all_cores()
    cb, (n_s, nq_s) = cpu_baseline(6.0)
    for w in range(args.warmup):
        cpu_search3_sample(n_s, nq_s, w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_search3_sample(n_s, nq_s, args.warmup + s)
    el = time.perf_counter() - t0
    value = nq_s * args.steps / el * (n_s / N_PER_GPU)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
            "config": {"workload": "cfg3: CohereEnhancedVectorDB 3-phase search, 100M x 1024-bit codes, 1024-query batch, k=100, "
                                   "binary_oversample=10, int8_oversample=3 (CPU: bounded sample per step)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": cb["sample"].split(",")[0] + f" per step, {args.steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    from vectorragquantization_b200.sharded import CudaEngine, ShardedSearch3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_local = args.rows
    ctx = V.Context(local)
    lib = L.load()
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    # ---- database shard: codes + int8 rows generated in place on the device -------------------------------------
    index = V.BinaryIndex(D, ctx=ctx, payload_kind=L.PAYLOAD_INT8_RAW)
    index.reserve(n_local)
    base = rank * n_local
    t0 = time.time()
    CH = 8_000_000
    for off in range(0, n_local, CH):
        index.add_synthetic(DB_SEED, base + off, min(CH, n_local - off), base + off)
    ctx.sync()
    t_build = time.time() - t0
    searcher = ShardedSearch3(CudaEngine(index, ctx), pos_base=base)

    # ---- queries: a different batch every step, resident on the device (value) and in pinned host memory (e2e) ----
    total_steps = args.warmup_actual + args.steps
    n_e2e = max(1, min(5, args.steps))
    qf_d = torch.empty((total_steps + n_e2e + 2, NQ, D), dtype=torch.float32, device=dev)
    qb_d = torch.empty((total_steps + n_e2e + 2, NQ, D // 8), dtype=torch.uint8, device=dev)
    for s in range(qf_d.shape[0]):
        L.check(lib.vrq_synth_f32(ctx.handle, Q_SEED, s * NQ, NQ, D, 0, L.ptr(qf_d[s])))
        L.check(lib.vrq_synth_codes_int8(ctx.handle, Q_SEED, s * NQ, NQ, D, L.ptr(qb_d[s]), None))
    ctx.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for s in range(args.warmup_actual):
        searcher.search(qf_d[s], qb_d[s], K, BO, IO)
    barrier()
    ctx.enable_timing(True)
    ctx.timing_ms("scan")
    ctx.timing_ms("scan_dense")
    ctx.timing_ms("rescore")
    ctx.timing_ms("merge")
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for s in range(args.steps):
        out = searcher.search(qf_d[args.warmup_actual + s], qb_d[args.warmup_actual + s], K, BO, IO)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    scan_ms, scan_n = ctx.timing_ms("scan")
    dense_ms, dense_n = ctx.timing_ms("scan_dense")
    resc_ms, _ = ctx.timing_ms("rescore")
    merge_ms, _ = ctx.timing_ms("merge")
    ctx.enable_timing(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    check = int(out["count"].sum().item())
    ms_per_step = ms / args.steps
    scale = n_local / N_PER_GPU
    value = NQ * world * scale / (ms_per_step / 1e3)

    # ---- e2e: host buffers in, host results out, through the public API / C ABI -----------------------------------
    qf_h = [torch.empty((NQ, D), dtype=torch.float32).pin_memory() for _ in range(n_e2e)]
    qb_h = [torch.empty((NQ, D // 8), dtype=torch.uint8).pin_memory() for _ in range(n_e2e)]
    for i in range(n_e2e):
        qf_h[i].copy_(qf_d[total_steps + i])
        qb_h[i].copy_(qb_d[total_steps + i])
    barrier()
    h2d = NQ * D * 4 + NQ * D // 8
    d2h = NQ * K * (8 + 4 + 8 + 8) + NQ * 4
    def e2e_call(i):
        if world == 1:
            return index.search3(qf_h[i].numpy(), qb_h[i].numpy(), K, BO, IO)  # C ABI, host pointers: H2D + kernels + D2H inside
        o_ = searcher.search(qf_h[i].to(dev, non_blocking=True), qb_h[i].to(dev, non_blocking=True), K, BO, IO)
        return [o_[k_].cpu() for k_ in ("labels", "hamming", "score_binary", "score_cosine", "count")]

    e2e_call(0)  # warm-up of the host-buffer path (its staging buffers are allocated on first use)
    barrier()
    e2e_calls_ms = []
    t0 = time.perf_counter()
    for i in range(n_e2e):
        tc = time.perf_counter()
        res = e2e_call(i)
        e2e_calls_ms.append((time.perf_counter() - tc) * 1e3)
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = NQ * world * scale / e2e_s

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    pairs_per_step = NQ * n_local
    scan_s = scan_ms / 1e3 / max(scan_n, 1)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    alu_peak = LOP3_PER_CLK_PER_SM * 148 * sm_mhz * 1e6 / LOP3_PER_PAIR  # (query, code) pairs per second at the measured clock
    popc_ceiling = POPC_PER_CLK_PER_SM * 148 * sm_mhz * 1e6 / 32.0
    mma_path = os.environ.get("VRQ_SCAN_MMA", "1") != "0"
    f4 = os.environ.get("VRQ_MMA_KIND", "4") != "8"
    if mma_path and dense_n > 0:
        dense_s = dense_ms / 1e3 / dense_n
        ops = 2.0 * D * pairs_per_step  # one multiply-add per (query bit, code bit) = 2 ops
        bf16_burst, bf16_sust, bf16_src = bf16_peaks()
        nominal = FP4_DENSE_NOMINAL_TOPS if f4 else INT8_DENSE_NOMINAL_TOPS
        ratio = 4 if f4 else 2  # nominal rate of the operand kind relative to bf16
        pair = f4 and os.environ.get("VRQ_MMA_PAIR", "1") != "0"
        kind = (f"tcgen05.mma.cta_group::{2 if pair else 1}.kind::mxf4.block_scale (packed e2m1 operands, unit UE8M0 scales, f32 "
                "accumulate - exact: every product is +-1 and |sum| <= 1024" + ("; CTA pairs: M=256 queries, each CTA expands half of "
                "every 128-row tile)" if pair else ")") if f4 else "tcgen05.mma.cta_group::1.kind::i8 (int8 operands, s32 accumulate)")
        roofline = {
            "kernel": f"hamming_scan_mma_kernel<{'e2m1' if f4 else 'int8'}>, dense pass: {kind}; M=128 queries resident in TMEM x N=128 "
                      "codes expanded from bits in shared memory, K=1024; 1024-query batch",
            "bound": "tensor", "unit": "TFLOP/s", "ops": "multiply-add of a query bit and a code bit = 2 ops",
            "achieved": ops / dense_s / 1e12, "peak": nominal, "frac": ops / dense_s / 1e12 / nominal,
            "peak_source": f"nominal dense {'fp4' if f4 else 'int8'} tensor rate (B200_PROFILING.md table; MEASURED_PEAKS.json holds only a "
                           f"cuBLAS bf16 figure). {ratio} x the {bf16_src} bf16 numbers would be {ratio * bf16_burst:.0f} (burst) / "
                           f"{ratio * bf16_sust:.0f} (sustained) TFLOP/s - see frac_of_scaled_measured_bf16. The kernel is bound by the "
                           "shared-memory bandwidth that feeds the B operand (expanded in place, never in HBM) before the tensor pipe",
            "frac_of_scaled_measured_bf16": {"burst": ops / dense_s / 1e12 / (ratio * bf16_burst),
                                             "sustained": ops / dense_s / 1e12 / (ratio * bf16_sust)},
            "kernel_ms": dense_s * 1e3, "pairs_per_s": pairs_per_step / dense_s,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch at 100 M rows x 1024 queries from one
            # `ncu --set full` capture (profiles/r01/ncu_scan_mma_final_raw.csv): 25.54 GB + 0.15 GB, against 12.8 GB of codes
            # requested 8 times (once per 128-query tile) = 102 GB: most re-reads are served by L2
            "traffic": 25.70e9 if (f4 and pair and n_local == N_PER_GPU) else None, "traffic_unit": "bytes per launch (DRAM)",
            "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": n_local * 128 * 8 / dense_s / 1e9, "peak": hbm_peak,
                    "frac": n_local * 128 * 8 / dense_s / 1e9 / hbm_peak, "peak_source": peak_src,
                    "note": "algorithmic bytes = 128 B per code per 128-query tile (8 tiles per 1024-query batch, 7 of them served by "
                            "L2: ncu dram__bytes_read ~ 13-15 GB per 102 GB requested); not the binding resource for a query batch"},
            "integer_pipe_kernel": {"note": "scan.cu (XOR + carry-save POPC) handles <= 3 queries per pass and VRQ_SCAN_MMA=0; "
                                            "its 1024-query rate measured earlier in round 1 was 216 Gpair/s (profiles/r01)",
                                    "alu_peak_Gpair_s": alu_peak / 1e9},
            "scan_ms_per_step": scan_ms / args.steps, "rescore_ms_per_step": resc_ms / args.steps,
            "merge_ms_per_step": merge_ms / args.steps,
        }
    else:
        roofline = {
            "kernel": "hamming_scan_kernel<TMA, 16 consumer warps, 16-CSA> (batched: 1024 queries per pass)",
            "bound": "alu", "unit": "Gpair/s",
            "achieved": pairs_per_step / scan_s / 1e9, "peak": alu_peak / 1e9, "frac": pairs_per_step / scan_s / alu_peak,
            "peak_source": f"ALU-pipe (LOP3) issue rate measured at 64 lanes/clk/SM (profiles/microbench) x 148 SMs x {sm_mhz:.0f} MHz "
                           "sampled during the run / 64 LOP3 per 1024-bit pair (32 XOR + 16 carry-save adders)",
            "plain_popc_ceiling": popc_ceiling / 1e9, "traffic": None,
            "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": n_local * 128 / scan_s / 1e9, "peak": hbm_peak,
                    "frac": n_local * 128 / scan_s / 1e9 / hbm_peak, "peak_source": peak_src},
            "scan_ms_per_step": scan_ms / args.steps, "rescore_ms_per_step": resc_ms / args.steps,
            "merge_ms_per_step": merge_ms / args.steps,
        }

    extras = {}
    if world == 1 and not args.no_extras:
        extras = hbm_bound_kernels(torch, ctx, lib, L, index, qb_d, dev, stream, hbm_peak, peak_src, n_local)
    cb = None
    if world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "e2m1/int8 tensor-core contraction of u8 bit codes, exact (Phase I) / f64 (Phases II-III)",
            "data": "synthetic",
            "config": {"workload": "cfg3: CohereEnhancedVectorDB 3-phase search, Hamming top-1000 over 100M x 1024-bit codes per GPU, "
                                   "1024-query batch, k=100, binary_oversample=10, int8_oversample=3",
                       "rows_per_gpu": n_local, "rows_total": n_local * world, "queries_per_step": NQ,
                       "parallelism": f"row-shard x{world}, per-shard top-k + 1 NCCL all-gather + device merge" if world > 1 else "1 GPU",
                       "cache": "inputs larger than L2: 12.8 GB of codes streamed per step, a new query batch every step",
                       "value_definition": "queries/s x (database rows / 100M); equals plain QPS@100M at 1 GPU",
                       "warmup_steps_run": args.warmup_actual, "db_build_s": round(t_build, 2), "result_checksum": check},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "calls_ms": [round(x, 3) for x in e2e_calls_ms], "steps": n_e2e},
            }
    line.update(extras)
    if cb is not None:
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def hbm_bound_kernels(torch, ctx, lib, L, index, qb_d, dev, stream, hbm_peak, peak_src, n_local):
    h = ctx.handle
    """The two HBM-bound kernels BASELINE.json's metric names, measured with CUDA events on the launching stream:
    Hamming scan with <= 2 queries per pass (128 B per code) and the fused encoders (4096 B in + codes out per row)."""
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps / 1e3

    kk = K * BO
    dist_d = torch.empty((16, kk), dtype=torch.int32, device=dev)
    lab_d = torch.empty((16, kk), dtype=torch.int64, device=dev)
    for nq in (1, 2, 16):
        q = qb_d[0][:nq].contiguous()
        s = timed(lambda: L.check(lib.vrq_index_search(index._h, nq, L.ptr(q), kk, L.ptr(dist_d), L.ptr(lab_d))), 5)
        gbs = n_local * 128 / s / 1e9
        kern = ("hamming_scan_kernel<true> (XOR + POPC)" if nq < 4 else
                "hamming_scan_mma_few_kernel (tcgen05, database rows = M expanded into TMEM) incl. its sample pass")
        out[f"roofline_scan_stream_nq{nq}"] = {"kernel": f"{kern} + merge, {nq} query/pass, top-{kk}", "bound": "hbm",
                                                "unit": "GB/s", "achieved": gbs, "peak": hbm_peak, "frac": gbs / hbm_peak,
                                                "peak_source": peak_src, "ms": s * 1e3, "traffic": None}
    # BASELINE config 5: Phase III micro - 4096 queries x 1000 gathered int8 candidates each (HBM-gather-bound)
    codes_p, _, pay_p, _ = index.device_ptrs()
    if pay_p:
        nq5, m5 = 4096, 1000
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        pos5 = torch.randint(0, n_local, (nq5, m5), dtype=torch.int64, device=dev, generator=g)
        qf5 = torch.empty((nq5, D), dtype=torch.float32, device=dev)
        L.check(lib.vrq_synth_f32(ctx.handle, 9, 0, nq5, D, 0, L.ptr(qf5)))
        sc5 = torch.empty((nq5, m5), dtype=torch.float64, device=dev)
        s3 = timed(lambda: L.check(lib.vrq_rescore_int8cos(h, pay_p, n_local, D, L.ptr(pos5), nq5, m5, L.ptr(qf5), L.ptr(sc5))), 5)
        s2 = timed(lambda: L.check(lib.vrq_rescore_binary(h, codes_p, n_local, D, L.ptr(pos5), nq5, m5, L.ptr(qf5), L.ptr(sc5))), 5)
        gb3 = nq5 * m5 * 1024 / s3 / 1e9
        out["roofline_rescore_int8cos"] = {"kernel": "rescore_int8cos_async_kernel<d=1024> (CUDA-core path, float64 accumulation, cp.async shared-memory ring: 128 KB in flight per SM)",
                                           "workload": "cfg5: 4096 queries x 1000 gathered int8 candidates, random positions over the resident rows",
                                           "bound": "hbm", "unit": "GB/s", "achieved": gb3, "peak": hbm_peak, "frac": gb3 / hbm_peak,
                                           "peak_source": peak_src, "ms": s3 * 1e3, "traffic": None,
                                           "pairs_per_s": nq5 * m5 / s3,
                                           "gather_roofline": "random 1 KB rows stream at 6.8 TB/s with >= 128 KB in flight per SM (profiles/microbench/gather_bench_r01.txt)",
                                           "imma_path": "not built: 1 KB gathered per (query, candidate) with no operand reuse, so the kernel is "
                                                        "bound by the gather; tensor cores have nothing to amortise (DESIGN.md 3.3)"}
        out["rescore_binary_cfg5"] = {"ms": s2 * 1e3, "pairs_per_s": nq5 * m5 / s2, "GB/s": nq5 * m5 * 128 / s2 / 1e9}
    # BASELINE config 2: global-limit encode of 10 M x 1024 float32 rows resident in HBM (41 GB in); the search index is
    # released first so that input + every output fit
    index.close()
    torch.cuda.empty_cache()
    n_enc = 10_000_000
    x = torch.empty((n_enc, D), dtype=torch.float32, device=dev)
    L.check(lib.vrq_synth_f32(ctx.handle, 7, 0, n_enc, D, 1, L.ptr(x)))
    ub = torch.empty((n_enc, D // 8), dtype=torch.uint8, device=dev)
    q8 = torch.empty((n_enc, D), dtype=torch.int8, device=dev)
    q16 = torch.empty((n_enc, D), dtype=torch.int16, device=dev)
    lo = torch.empty((n_enc,), dtype=torch.float64, device=dev)
    hi = torch.empty((n_enc,), dtype=torch.float64, device=dev)
    h = ctx.handle
    cases = {
        "int8_global+ubinary": (lambda: L.check(lib.vrq_quantize_int8_global(h, L.ptr(x), n_enc, D, 0.3, L.ptr(q8), L.ptr(ub))), 4096 + 1024 + 128),
        "int16_global+ubinary": (lambda: L.check(lib.vrq_quantize_int16_global(h, L.ptr(x), n_enc, D, 1.0, L.ptr(q16), L.ptr(ub))), 4096 + 2048 + 128),
        "int4+ubinary": (lambda: L.check(lib.vrq_quantize_int4(h, L.ptr(x), n_enc, D, L.ptr(q8), L.ptr(lo), L.ptr(hi), L.ptr(ub))), 4096 + 512 + 16 + 128),
        "int8_perdoc+ubinary": (lambda: L.check(lib.vrq_quantize_int8_perdoc(h, L.ptr(x), n_enc, D, L.ptr(q8), L.ptr(lo), L.ptr(hi), L.ptr(ub))), 4096 + 1024 + 8 + 128),
        "ubinary_only": (lambda: L.check(lib.vrq_to_binary_f32(h, L.ptr(x), n_enc, D, 0, L.ptr(ub))), 4096 + 128),
    }
    enc = {}
    for name, (fn, bpr) in cases.items():
        s = timed(fn, 5)
        gbs = n_enc * bpr / s / 1e9
        enc[name] = {"GB/s": gbs, "frac": gbs / hbm_peak, "ms": s * 1e3, "bytes_per_row": bpr, "rows": n_enc}
    # the same encoder through the C ABI with HOST buffers (pinned): H2D of x and D2H of codes inside the call
    n_h = 1_000_000
    xh = torch.empty((n_h, D), dtype=torch.float32).pin_memory()
    xh.copy_(x[:n_h])
    q8h = torch.empty((n_h, D), dtype=torch.int8).pin_memory()
    ubh = torch.empty((n_h, D // 8), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize(dev)
    th = 1e30
    for it in range(3):  # the first call allocates the library's staging buffers; report the best warm call
        t0 = time.perf_counter()
        L.check(lib.vrq_quantize_int8_global(h, L.ptr(xh), n_h, D, 0.3, L.ptr(q8h), L.ptr(ubh)))
        if it:
            th = min(th, time.perf_counter() - t0)
    enc["int8_global+ubinary_host_buffers"] = {"GB/s": n_h * 5248 / th / 1e9, "ms": th * 1e3, "rows": n_h,
                                               "note": "end to end with pinned host input/output: PCIe-bound (4096 B in + 1152 B out per row)"}
    out["roofline_encode"] = {"kernel": "encode1024_ring_kernel<INT8_GLOBAL, ubinary fused> (cp.async ring of rows per warp, magic-number rounding)", "bound": "hbm", "unit": "GB/s",
                              "achieved": enc["int8_global+ubinary"]["GB/s"], "peak": hbm_peak,
                              "frac": enc["int8_global+ubinary"]["frac"], "peak_source": peak_src, "rows": n_enc, "traffic": None,
                              "workload": "cfg2: 10M x 1024 float32 rows resident in HBM", "all_codecs": enc}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_PER_GPU, help="rows per GPU (default: the 100M of the headline config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the encode / stream-scan roofline legs")
    args = ap.parse_args()
    args.warmup_actual = max(3, args.warmup) if args.impl == "b200" else args.warmup  # timing rule: >= 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
