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

The JSON line carries, besides the base contract:
  roofline                 dominant kernel = dense pass of the batched Hamming scan (tcgen05 e2m1 contraction), against the
                           MEASURED mxf4 issue ceiling (profiles/microbench/mxf4_peak.cu -> profiles/r02/mxf4_peak.json)
  roofline_scan_stream_*   the HBM-bound scan regimes (1 / 2 / 16 queries per pass), roofline_encode (cfg2),
  roofline_rescore_int8cos cfg5: Phase III, CUDA-core path vs tensor-core (IMMA) path; rescore_binary_cfg5 (Phase II)
  cfg1                     VectorDBInt8 class flow on 10 k documents, literal CPU loop vs the GPU class
  cfg4                     the same 3-phase search over 1 BILLION codes row-sharded over the N GPUs of the run (strong scaling)
  adversarial              a step whose sampled thresholds fail for 64 queries (cost of the exact fallback pass)
  parity                   a small sharded problem checked against the single-index search and the CPU oracle before timing
  cpu_baseline, e2e, clocks, gpu_launches
`traffic` values are DRAM bytes per launch from `ncu --set full` captures, read from profiles/r02/traffic.json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
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
N_CFG4 = 1_000_000_000
DB_SEED, Q_SEED = 1, 2
POPC_PER_CLK_PER_SM = 16.0  # measured: profiles/microbench/popc_bench_r01.txt
LOP3_PER_CLK_PER_SM = 64.0  # measured: same file
LOP3_PER_PAIR = 64.0        # 32 XOR + 16 carry-save adders x 2 LOP3 (scan.cu: hamming128_csa); + 16 POPC on the XU pipe
METRIC = "3-phase search QPS @100Mx1024-d"
UNIT = "queries/s per 100M codes"
WORKLOAD = ("cfg3: CohereEnhancedVectorDB 3-phase search, Hamming top-1000 over 100M x 1024-bit codes per GPU, "
            "1024-query batch, k=100, binary_oversample=10, int8_oversample=3")

INT8_DENSE_NOMINAL_TOPS = 4500.0  # B200 dense int8 / fp8 tensor rate (B200_PROFILING.md, nominal table)
FP4_DENSE_NOMINAL_TOPS = 9000.0   # B200 dense fp4 tensor rate (same table)
GATHER_1KB_GBS = 6800.0           # random 1 KB rows, >= 128 KB in flight per SM (profiles/microbench/gather_bench_r01.txt)


def _load_json(rel, default=None):
    p = os.path.join(ROOT, rel)
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return default


def peaks():
    j = _load_json("MEASURED_PEAKS.json")
    if j:
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bf16_peaks():
    j = _load_json("MEASURED_PEAKS.json")
    if j:
        return float(j["bf16_tflops"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def traffic(key):
    """DRAM bytes (read + write) per launch of a kernel from one `ncu --set full` capture (profiles/r02/traffic.json)."""
    j = _load_json("profiles/r02/traffic.json", {})
    v = j.get(key)
    return (float(v["dram_bytes"]) if isinstance(v, dict) else None), (v.get("source") if isinstance(v, dict) else None)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_search3_sample(n_s: int, nq_s: int, step: int = 0):
    """The reference's CPU path on a bounded sample of the workload: faiss-equivalent Hamming scan (C/OpenMP over
    queries, the way faiss parallelises IndexBinaryFlat.search) over the first n_s rows, then the reference's literal
    per-candidate NumPy loops for phases II and III (CohereEnhancedVectorDB.py:283-322).
    Returns (scan seconds, rescoring seconds): the scan is proportional to the database size, the rescoring loops cost
    k*binary_oversample + k*int8_oversample candidates per query whatever the database size."""
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
    t1 = time.perf_counter()
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
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_model(nq_s, n_s, scan_s, resc_s):
    """queries/s at 100 M rows from the two measured terms: only the scan scales with the database."""
    return nq_s / (scan_s * (N_PER_GPU / n_s) + resc_s)


def cpu_baseline(target_s: float = 12.0):
    from oracle import oracle_c as oc
    cores = oc.use_all_cores()
    n_s, nq_s = 2_000_000, max(cores, 8)
    ts, tr = cpu_search3_sample(n_s, nq_s)
    while ts + tr < target_s / 3 and nq_s < 512:
        nq_s *= 2
        ts, tr = cpu_search3_sample(n_s, nq_s)
    value = cpu_model(nq_s, n_s, ts, tr)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "scan_s": ts, "rescore_s": tr, "sample_rows": n_s, "sample_queries": nq_s,
            "model": "value = queries / (scan_s * 100M / sample_rows + rescore_s): the Hamming scan is proportional to the database, "
                     "phases II / III cost 1000 + 300 candidates per query at any size",
            "sample": f"{nq_s} queries x first {n_s} of the 100M synthetic codes: C/OpenMP scan over queries {ts:.2f} s (x{N_PER_GPU // n_s} "
                      f"to 100M rows) + the reference's per-candidate NumPy loops for phases II/III {tr:.2f} s (not scaled)"}, (n_s, nq_s)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c as oc
    cores = oc.use_all_cores()
    cb, (n_s, nq_s) = cpu_baseline(6.0)
    for w in range(args.warmup):
        cpu_search3_sample(n_s, nq_s, w)
    ts = tr = 0.0
    t0 = time.perf_counter()
    for s in range(args.steps):
        a, b = cpu_search3_sample(n_s, nq_s, args.warmup + s)
        ts += a
        tr += b
    el = time.perf_counter() - t0
    value = cpu_model(nq_s * args.steps, n_s, ts, tr)
    step_model_ms = (ts * (N_PER_GPU / n_s) + tr) / args.steps * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_model_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": N_PER_GPU, "queries_per_step": nq_s,
                       "note": "CPU arm: every step is a bounded sample (sample_queries x sample_rows); ms_per_step is the modelled time "
                               "of the sampled queries over 100M rows, measured_ms_per_step the wall time of the sample itself"},
            "measured_ms_per_step": el / args.steps * 1e3,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "scan_s": ts, "rescore_s": tr,
                             "sample_rows": n_s, "sample_queries": nq_s * args.steps, "model": cb["model"],
                             "sample": f"{nq_s} queries x first {n_s} of the 100M synthetic codes per step, {args.steps} steps: scan {ts:.2f} s "
                                       f"(scaled x{N_PER_GPU // n_s}), phases II/III loops {tr:.2f} s (not scaled)"},
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
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
class Env:
    """What every leg needs: torch, the library, the device, the process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import vectorragquantization_b200 as V
        from vectorragquantization_b200 import _lib as L
        self.torch, self.dist, self.V, self.L = torch, dist, V, L
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = V.Context(self.local)
        self.lib = L.load()
        self.stream = torch.cuda.current_stream(self.dev)
        self.ctx.set_stream(self.stream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def queries(self, nbatches, seed=Q_SEED, first=0):
        torch, L = self.torch, self.L
        qf = torch.empty((nbatches, NQ, D), dtype=torch.float32, device=self.dev)
        qb = torch.empty((nbatches, NQ, D // 8), dtype=torch.uint8, device=self.dev)
        for s in range(nbatches):
            L.check(self.lib.vrq_synth_f32(self.ctx.handle, seed, (first + s) * NQ, NQ, D, 0, L.ptr(qf[s])))
            L.check(self.lib.vrq_synth_codes_int8(self.ctx.handle, seed, (first + s) * NQ, NQ, D, L.ptr(qb[s]), None))
        self.ctx.sync()
        return qf, qb

    def build_index(self, rows, base, seed=DB_SEED, resident_payload=True):
        V, L = self.V, self.L
        index = V.BinaryIndex(D, ctx=self.ctx, payload_kind=L.PAYLOAD_INT8_RAW)
        if not resident_payload:
            index.set_synthetic_payload(seed, base)
        index.reserve(rows)
        t0 = time.time()
        CH = 8_000_000
        for off in range(0, rows, CH):
            index.add_synthetic(seed, base + off, min(CH, rows - off), base + off)
        self.ctx.sync()
        return index, time.time() - t0


def parity_probe(env):
    """Before anything is timed: a small sharded problem (3 000 001 rows over the ranks of this run, 64 queries) answered by
    the row-sharded path (per-shard search3_local -> all-gather -> merge3), compared on rank 0 with the single-index search3
    over the whole database and with the CPU oracle (phase I: C scan of every query over the whole database; phases II / III:
    the reference's literal loops for 8 of the queries).  The oracle is only the checker here."""
    torch, V, L = env.torch, env.V, env.L
    from vectorragquantization_b200.sharded import CudaEngine, ShardedSearch3, shard_range
    n_total, nq, seed, qseed = 3_000_001, 64, 3, 4
    a, b = shard_range(n_total, env.rank, env.world)
    shard, _ = env.build_index(b - a, a, seed)
    qf = torch.empty((nq, D), dtype=torch.float32, device=env.dev)
    qb = torch.empty((nq, D // 8), dtype=torch.uint8, device=env.dev)
    L.check(env.lib.vrq_synth_f32(env.ctx.handle, qseed, 0, nq, D, 0, L.ptr(qf)))
    L.check(env.lib.vrq_synth_codes_int8(env.ctx.handle, qseed, 0, nq, D, L.ptr(qb), None))
    out = ShardedSearch3(CudaEngine(shard, env.ctx), pos_base=a).search(qf, qb, K, BO, IO)
    env.barrier()
    got = {k_: v.cpu().numpy().copy() for k_, v in out.items()}
    shard.close()
    res = None
    if env.rank == 0:
        res = {"n_ranks": env.world, "rows": n_total, "queries": nq, "k": K, "binary_oversample": BO, "int8_oversample": IO}
        single, _ = env.build_index(n_total, 0, seed)
        want = single.search3(qf.cpu().numpy(), qb.cpu().numpy(), K, BO, IO)
        single.close()
        names = ("labels", "hamming", "score_binary", "score_cosine", "count")
        res["equals_single_index"] = bool(all(np.array_equal(got[n_], w) for n_, w in zip(names, want)))
        try:
            from oracle import oracle_c as oc
            from oracle import vrq_oracle as o
            oc.use_all_cores()
            qf_h, qb_h = qf.cpu().numpy(), qb.cpu().numpy()
            codes, _ = oc.synth_codes_int8(seed, 0, n_total, want_int8=False)
            dist, pos = oc.hamming_topk(codes, qb_h, K * BO)
            nq_o, ids_ok, ham_ok, eb, ec = 8, True, True, 0.0, 0.0
            for qi in range(nq_o):
                ref = o.search3_after_phase1(dist[qi], pos[qi], lambda p: p, lambda p: codes[p],
                                             lambda p: np.stack([oc.synth_codes_int8(seed, int(r), 1, want_codes=False)[1][0] for r in p]),
                                             qf_h[qi], K, BO, IO)
                ids_ok &= [h["doc_id"] for h in ref] == got["labels"][qi].tolist()
                ham_ok &= [h["score_hamming"] for h in ref] == got["hamming"][qi].tolist()
                rb = np.array([h["score_binary"] for h in ref])
                rc = np.array([h["score_cosine"] for h in ref])
                eb = max(eb, float(np.max(np.abs(got["score_binary"][qi] - rb) / np.maximum(np.abs(rb), 1e-300))))
                rows = np.stack([oc.synth_codes_int8(seed, int(h["pos"]), 1, want_codes=False)[1][0] for h in ref])
                fl = o.rescore_int8cos_absfloor(qf_h[qi], rows)
                ec = max(ec, float(np.max(np.maximum(np.abs(got["score_cosine"][qi] - rc) - fl, 0.0) / np.maximum(np.abs(rc), 1e-300))))
            # phase I of ALL queries: the final hamming column must be consistent with the oracle's top-1000 sets
            p1_ok = all(set(got["labels"][qi].tolist()) <= set(pos[qi].tolist()) for qi in range(nq))
            res.update({"oracle_queries": nq_o, "ids_equal_oracle": bool(ids_ok), "hamming_equal_oracle": bool(ham_ok),
                        "winners_within_oracle_phase1_sets_all_queries": bool(p1_ok), "max_rel_err_score_binary": eb,
                        "max_rel_err_score_cosine_beyond_sdot_floor": ec, "tolerance": 1e-5})
            res["ok"] = bool(res["equals_single_index"] and ids_ok and ham_ok and p1_ok and eb <= 1e-5 and ec <= 1e-5)
        except Exception as e:  # the oracle is test infrastructure: its absence must not take the bench down
            res["oracle_error"] = repr(e)
            res["ok"] = bool(res["equals_single_index"])
    env.barrier()
    return res


def timed_search_loop(env, searcher, qf_d, qb_d, warmup, steps):
    torch = env.torch
    for s in range(warmup):
        searcher.search(qf_d[s], qb_d[s], K, BO, IO)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record(env.stream)
    out = None
    for s in range(steps):
        out = searcher.search(qf_d[warmup + s], qb_d[warmup + s], K, BO, IO)
    e1.record(env.stream)
    env.barrier()
    return env.max_over_ranks(e0.elapsed_time(e1)), out


def cfg4_leg(env, args, resident):
    """BASELINE.json configs[3]: the same 3-phase search over 1 BILLION codes row-sharded over the ranks of this run
    (strong scaling: N_CFG4 / world rows per GPU).  The int8 store is 1 TB: it is resident only when this rank's share
    fits in HBM (8 GPUs: 125 M rows = 16 GB codes + 128 GB int8 rows); otherwise Phase III regenerates the candidate rows
    from the counter-based generator (the "materialise candidate rows on demand" plan) - codes, Phase I,
    Phase II, the exchange and the merge are the real thing either way."""
    from vectorragquantization_b200.sharded import CudaEngine, ShardedSearch3
    rows = N_CFG4 // env.world
    base = env.rank * rows
    index, t_build = env.build_index(rows, base, DB_SEED, resident_payload=resident)
    searcher = ShardedSearch3(CudaEngine(index, env.ctx), pos_base=base)
    w, k_ = 2, max(2, min(5, args.steps))
    qf_d, qb_d = env.queries(w + k_, first=1000)
    sampler = ClockSampler(env.local)
    sampler.start()
    ms, out = timed_search_loop(env, searcher, qf_d, qb_d, w, k_)
    clocks = sampler.stop()
    check = int(out["count"].sum().item())
    index.close()
    env.torch.cuda.empty_cache()
    step = ms / k_
    return {"workload": "cfg4: 3-phase search over 1B x 1024-bit codes row-sharded across the GPUs of this run, 1024-query batch, k=100, "
                        "binary_oversample=10, int8_oversample=3",
            "rows_total": rows * env.world, "rows_per_gpu": rows, "n_gpus": env.world, "scaling": "strong", "steps": k_, "warmup": w,
            "ms_per_step": step, "queries_per_s": NQ / (step / 1e3), "value_in_headline_unit": NQ * (rows * env.world / N_PER_GPU) / (step / 1e3),
            "int8_payload": "resident in HBM (gathered by the Phase III kernel)" if resident else
                            "not resident (1 TB does not fit on this many GPUs): Phase III regenerates candidate rows from the counter-based generator (SURVEY H6)",
            "db_build_s": round(t_build, 2), "result_checksum": check, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}


def run_gpu(args):
    env = Env()
    torch, dist, V, L = env.torch, env.dist, env.V, env.L
    from vectorragquantization_b200.sharded import CudaEngine, ShardedSearch3
    rank, world, dev, ctx, lib, stream = env.rank, env.world, env.dev, env.ctx, env.lib, env.stream

    parity = None if args.no_parity else parity_probe(env)

    # ---- database shard: codes + int8 rows generated in place on the device -------------------------------------
    n_local = args.rows
    base = rank * n_local
    index, t_build = env.build_index(n_local, base)
    searcher = ShardedSearch3(CudaEngine(index, ctx), pos_base=base)

    # ---- queries: a different batch every step, resident on the device (value) and in pinned host memory (e2e) ----
    total_steps = args.warmup_actual + args.steps
    n_e2e = max(1, min(5, args.steps))
    qf_d, qb_d = env.queries(total_steps + n_e2e + 2)

    for s in range(args.warmup_actual):
        searcher.search(qf_d[s], qb_d[s], K, BO, IO)
    env.barrier()
    ctx.enable_timing(True)
    for cat in ("scan", "scan_dense", "rescore", "merge"):
        ctx.timing_ms(cat)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(env.local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record(stream)
    for s in range(args.steps):
        out = searcher.search(qf_d[args.warmup_actual + s], qb_d[args.warmup_actual + s], K, BO, IO)
    e1.record(stream)
    env.barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    scan_ms, scan_n = ctx.timing_ms("scan")
    dense_ms, dense_n = ctx.timing_ms("scan_dense")
    resc_ms, _ = ctx.timing_ms("rescore")
    merge_ms, _ = ctx.timing_ms("merge")
    ctx.enable_timing(False)
    ms = env.max_over_ranks(ms)
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
    env.barrier()
    h2d = NQ * D * 4 + NQ * D // 8
    d2h = NQ * K * (8 + 4 + 8 + 8) + NQ * 4

    def e2e_call(i):
        if world == 1:
            return index.search3(qf_h[i].numpy(), qb_h[i].numpy(), K, BO, IO)  # C ABI, host pointers: H2D + kernels + D2H inside
        o_ = searcher.search(qf_h[i].to(dev, non_blocking=True), qb_h[i].to(dev, non_blocking=True), K, BO, IO)
        return [o_[k_].cpu() for k_ in ("labels", "hamming", "score_binary", "score_cosine", "count")]

    e2e_call(0)  # warm-up of the host-buffer path (its staging buffers are allocated on first use)
    env.barrier()
    e2e_calls_ms = []
    t0 = time.perf_counter()
    for i in range(n_e2e):
        tc = time.perf_counter()
        e2e_call(i)
        e2e_calls_ms.append((time.perf_counter() - tc) * 1e3)
    env.barrier()
    e2e_s = env.max_over_ranks((time.perf_counter() - t0) / n_e2e)
    e2e_value = NQ * world * scale / e2e_s

    hbm_peak, peak_src = peaks()
    roofline, extras = None, {}
    if rank == 0:
        roofline = headline_roofline(args, clocks, n_local, scan_ms, scan_n, dense_ms, dense_n, resc_ms, merge_ms, hbm_peak, peak_src)
        if world == 1 and not args.no_extras:
            extras = hbm_bound_kernels(env, index, qf_d, qb_d, hbm_peak, peak_src, n_local, ms_per_step)
    index.close()
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_extras:
        extras.update(encode_legs(env, hbm_peak, peak_src))
        extras["cfg1"] = cfg1_leg(env)
    torch.cuda.empty_cache()

    # ---- cfg4: 1 billion codes over the GPUs of this run ------------------------------------------------------------
    cfg4 = {}
    if not args.no_cfg4:
        rows4 = N_CFG4 // world
        need_resident = rows4 * (128 + 1024) + (8 << 30)
        fits = torch.cuda.mem_get_info(dev)[0] > need_resident
        if world > 1:
            t = torch.tensor([1 if fits else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            fits = bool(t.item())
        cfg4["cfg4"] = cfg4_leg(env, args, resident=fits)
        if fits:  # the same code path as the smaller runs, for a like-for-like 1 -> N comparison
            cfg4["cfg4_regenerated_payload"] = cfg4_leg(env, args, resident=False)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cb = None
    if world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "e2m1/int8 tensor-core contraction of u8 bit codes, exact (Phase I) / f64 + s8 IMMA on fixed-point digits (Phases II-III)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "rows_per_gpu": n_local, "rows_total": n_local * world, "queries_per_step": NQ,
                       "parallelism": f"row-shard x{world}, per-shard top-k + 1 NCCL all-gather + device merge" if world > 1 else "1 GPU",
                       "cache": "inputs larger than L2: 12.8 GB of codes streamed per step, a new query batch every step",
                       "value_definition": "queries/s x (database rows / 100M); equals plain QPS@100M at 1 GPU",
                       "warmup_steps_run": args.warmup_actual, "db_build_s": round(t_build, 2), "result_checksum": check},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "calls_ms": [round(x, 3) for x in e2e_calls_ms], "steps": n_e2e},
            }
    if parity is not None:
        line["parity"] = parity
    line.update(extras)
    line.update(cfg4)
    if cb is not None:
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def headline_roofline(args, clocks, n_local, scan_ms, scan_n, dense_ms, dense_n, resc_ms, merge_ms, hbm_peak, peak_src):
    pairs_per_step = NQ * n_local
    scan_s = scan_ms / 1e3 / max(scan_n, 1)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    alu_peak = LOP3_PER_CLK_PER_SM * 148 * sm_mhz * 1e6 / LOP3_PER_PAIR  # (query, code) pairs per second at the measured clock
    popc_ceiling = POPC_PER_CLK_PER_SM * 148 * sm_mhz * 1e6 / 32.0
    mma_path = os.environ.get("VRQ_SCAN_MMA", "1") != "0"
    f4 = os.environ.get("VRQ_MMA_KIND", "4") != "8"
    if not (mma_path and dense_n > 0):
        return {
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
    dense_s = dense_ms / 1e3 / dense_n
    ops = 2.0 * D * pairs_per_step  # one multiply-add per (query bit, code bit) = 2 ops
    achieved = ops / dense_s / 1e12
    bf16_burst, bf16_sust, bf16_src = bf16_peaks()
    nominal = FP4_DENSE_NOMINAL_TOPS if f4 else INT8_DENSE_NOMINAL_TOPS
    ratio = 4 if f4 else 2  # nominal rate of the operand kind relative to bf16
    pair = f4 and os.environ.get("VRQ_MMA_PAIR", "1") != "0"
    # measured ceiling of the instruction itself: bare issue loop, operands resident, no epilogue (mxf4_peak.cu)
    mp = _load_json("profiles/r02/mxf4_peak.json") if f4 else None
    mkey = "cta_group2_n128_scan" if pair else "cta_group1_n128_scan"
    measured = float(mp[mkey]["tflops"]) if mp and mkey in mp else None
    peak = measured if measured else nominal
    kind = (f"tcgen05.mma.cta_group::{2 if pair else 1}.kind::mxf4.block_scale (packed e2m1 operands, unit UE8M0 scales, f32 "
            "accumulate - exact: every product is +-1 and |sum| <= 1024" + ("; CTA pairs: M=256 queries, each CTA expands half of "
            "every 128-row tile)" if pair else ")") if f4 else "tcgen05.mma.cta_group::1.kind::i8 (int8 operands, s32 accumulate)")
    tr, tr_src = traffic("scan_dense_100M_1024q") if (f4 and pair and n_local == N_PER_GPU) else (None, None)
    return {
        "kernel": f"hamming_scan_mma_kernel<{'e2m1' if f4 else 'int8'}>, dense pass: {kind}; M=128 queries resident in TMEM x N=128 "
                  "codes expanded from bits in shared memory, K=1024; 1024-query batch; epilogue warps at 128 registers (setmaxnreg): "
                  "the accumulator goes back to the issuer before it is examined",
        "bound": "tensor", "unit": "TFLOP/s", "ops": "multiply-add of a query bit and a code bit = 2 ops",
        "achieved": achieved, "peak": peak, "frac": achieved / peak,
        "peak_source": (f"MEASURED: bare {mkey} issue loop of profiles/microbench/mxf4_peak.cu on this pool's B200 "
                        f"(profiles/r02/mxf4_peak.json: operands resident, no TMA / expansion / epilogue, the scan's own one-hot nibble data, "
                        f"{mp[mkey].get('sm_mhz', '?')} MHz); nominal dense fp4 = {nominal:.0f} TFLOP/s" if measured else
                        f"nominal dense {'fp4' if f4 else 'int8'} tensor rate (B200_PROFILING.md table; no measured mxf4 figure found)"),
        "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
        "frac_of_scaled_measured_bf16": {"burst": achieved / (ratio * bf16_burst), "sustained": achieved / (ratio * bf16_sust), "source": bf16_src},
        "kernel_ms": dense_s * 1e3, "pairs_per_s": pairs_per_step / dense_s,
        "traffic": tr, "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": tr_src,
        "algorithmic_bytes": n_local * 128,
        "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": n_local * 128 * 8 / dense_s / 1e9, "peak": hbm_peak,
                "frac": n_local * 128 * 8 / dense_s / 1e9 / hbm_peak, "peak_source": peak_src,
                "note": "requested bytes = 128 B per code per 128-query tile (8 tiles per 1024-query batch; the re-reads are served by "
                        "L2 when the CTA pairs of a strip stay in lockstep); not the binding resource for a query batch"},
        "integer_pipe_kernel": {"note": "scan.cu (XOR + carry-save POPC) handles <= 2 queries per pass and VRQ_SCAN_MMA=0; "
                                        "its 1024-query rate measured in round 1 was 216 Gpair/s (profiles/r01)",
                                "alu_peak_Gpair_s": alu_peak / 1e9},
        "scan_ms_per_step": scan_ms / args.steps, "rescore_ms_per_step": resc_ms / args.steps,
        "merge_ms_per_step": merge_ms / args.steps,
    }


def _timed(env, fn, reps):
    torch = env.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize(env.dev)
    e0.record(env.stream)
    for _ in range(reps):
        fn()
    e1.record(env.stream)
    torch.cuda.synchronize(env.dev)
    return e0.elapsed_time(e1) / reps / 1e3


def hbm_bound_kernels(env, index, qf_d, qb_d, hbm_peak, peak_src, n_local, ms_per_step):
    """Legs that need the resident 100 M-row index: the HBM-bound scan regimes, cfg5 (Phase III / Phase II on gathered
    candidates, CUDA-core vs tensor-core path) and the adversarial step."""
    torch, L, lib, ctx, dev = env.torch, env.L, env.lib, env.ctx, env.dev
    h = ctx.handle
    out = {}
    kk = K * BO
    dist_d = torch.empty((128, kk), dtype=torch.int32, device=dev)
    lab_d = torch.empty((128, kk), dtype=torch.int64, device=dev)
    for nq in (1, 2, 3, 16, 64, 96, 128):
        q = qb_d[0][:nq].contiguous()
        s = _timed(env, lambda: L.check(lib.vrq_index_search(index._h, nq, L.ptr(q), kk, L.ptr(dist_d), L.ptr(lab_d))), 5)
        gbs = n_local * 128 / s / 1e9
        kern = ("hamming_scan_kernel<true> (XOR + POPC)" if nq < 3 else
                ("hamming_scan_mma_wide_kernel (tcgen05, database rows = M expanded into TMEM, queries = N, bias column) incl. the "
                 "sample pass" if nq <= 96 else "hamming_scan_mma_kernel, one 128-query tile per CTA, incl. the sample pass"))
        tr, tr_src = traffic(f"scan_stream_nq{nq}")
        out[f"roofline_scan_stream_nq{nq}"] = {"kernel": f"{kern} + merge, {nq} query/pass, top-{kk}", "bound": "hbm",
                                                "unit": "GB/s", "achieved": gbs, "peak": hbm_peak, "frac": gbs / hbm_peak,
                                                "peak_source": peak_src, "ms": s * 1e3, "traffic": tr, "traffic_source": tr_src,
                                                "algorithmic_bytes": n_local * 128}
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
        sc5b = torch.empty((nq5, m5), dtype=torch.float64, device=dev)
        variants = {}
        saved = {k_: os.environ.get(k_) for k_ in ("VRQ_RESCORE_IMMA", "VRQ_RESCORE_IMMA_SHAPE", "VRQ_RESCORE_BIN")}

        def run3(dst):
            return lambda: L.check(lib.vrq_rescore_int8cos(h, pay_p, n_local, D, L.ptr(pos5), nq5, m5, L.ptr(qf5), L.ptr(dst)))

        os.environ["VRQ_RESCORE_IMMA"] = "0"
        variants["cuda_core_f64_cp_async_ring"] = _timed(env, run3(sc5), 5)
        os.environ["VRQ_RESCORE_IMMA"] = "1"
        for bulk, w_, s_ in ((1, 6, 2), (0, 6, 2), (1, 4, 3), (0, 4, 3), (1, 2, 3), (1, 3, 4), (0, 3, 4), (1, 12, 1), (0, 12, 1), (1, 8, 1), (0, 8, 1)):
            os.environ["VRQ_RESCORE_IMMA_SHAPE"] = str(100 * bulk + 10 * w_ + s_)
            variants[f"imma_s8_{'bulk1KB' if bulk else 'cpasync16B'}_w{w_}_s{s_}"] = _timed(env, run3(sc5b), 5)
        for k_, v in saved.items():
            if v is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v
        fin = torch.isfinite(sc5) & torch.isfinite(sc5b)
        dev_rel = float(((sc5 - sc5b).abs()[fin] / sc5.abs()[fin].clamp_min(1e-30)).max().item())
        s3 = _timed(env, run3(sc5), 5)  # the default path, whichever it is
        def run2():
            L.check(lib.vrq_rescore_binary(h, codes_p, n_local, D, L.ptr(pos5), nq5, m5, L.ptr(qf5), L.ptr(sc5)))

        bin_ms = {}
        for mode, name in (("0", "cuda_core_register_kernel"), ("1", "cuda_core_nibble_table"), ("2", "imma_s8_code_bits")):
            os.environ["VRQ_RESCORE_BIN"] = mode
            bin_ms[name] = _timed(env, run2, 5) * 1e3
        os.environ.pop("VRQ_RESCORE_BIN", None)
        if saved["VRQ_RESCORE_BIN"] is not None:
            os.environ["VRQ_RESCORE_BIN"] = saved["VRQ_RESCORE_BIN"]
        s2 = _timed(env, run2, 5)  # the default path
        gb3 = nq5 * m5 * 1024 / s3 / 1e9
        tr, tr_src = traffic("rescore_int8cos_cfg5")
        best = min(variants, key=variants.get)
        out["roofline_rescore_int8cos"] = {
            "kernel": "vrq_rescore_int8cos, default path (d=1024)",
            "workload": "cfg5: 4096 queries x 1000 gathered int8 candidates, random positions over the resident rows",
            "bound": "hbm", "unit": "GB/s", "achieved": gb3, "peak": GATHER_1KB_GBS, "frac": gb3 / GATHER_1KB_GBS,
            "peak_source": "measured gather roofline: random 1 KB rows stream at 6.8 TB/s with >= 128 KB in flight per SM "
                           "(profiles/microbench/gather_bench_r01.txt)",
            "frac_of_hbm_copy_peak": gb3 / hbm_peak, "hbm_copy_peak": hbm_peak, "ms": s3 * 1e3, "traffic": tr, "traffic_source": tr_src,
            "algorithmic_bytes": nq5 * m5 * 1024, "pairs_per_s": nq5 * m5 / s3,
            "imma_vs_cuda_core": {"ms": {k_: v * 1e3 for k_, v in variants.items()}, "fastest": best,
                                  "GB/s": {k_: nq5 * m5 * 1024 / v / 1e9 for k_, v in variants.items()},
                                  "max_rel_deviation_between_paths": dev_rel,
                                  "paths": "cuda_core: float64 FMA of exact products, cp.async ring (rescore.cu); imma: mma.sync.m16n8k32.s8 over eight "
                                           "base-256 digits of the 64-bit fixed-point query, 1 KB bulk copies into a per-warp ring (rescore_mma.cu), "
                                           "w = warps per block, s = ring stages"}}
        out["rescore_binary_cfg5"] = {"kernel": "vrq_rescore_binary, default path (d=1024): mma.sync s8, code bits x digit planes of the fixed-point query",
                                      "ms": s2 * 1e3, "pairs_per_s": nq5 * m5 / s2, "GB/s": nq5 * m5 * 128 / s2 / 1e9,
                                      "variants_ms": bin_ms, "speedup_vs_round1_register_kernel": bin_ms["cuda_core_register_kernel"] / (s2 * 1e3),
                                      "gather_roofline_128B_GBs": 4200.0, "frac": nq5 * m5 * 128 / s2 / 1e9 / 4200.0,
                                      "traffic": traffic("rescore_binary_cfg5")[0], "traffic_source": traffic("rescore_binary_cfg5")[1],
                                      "algorithmic_bytes": nq5 * m5 * 128,
                                      "peak_source": "random 128-byte rows stream at 4.0-4.4 TB/s (profiles/microbench/gather_bench_r01.txt)"}
        del pos5, qf5, sc5, sc5b
    # adversarial steps.  (a) 64 of the 1024 queries have 32 exact duplicates each inside the first sampled tiles.  With round 1's
    # list-based sample pass their sampled threshold was 0 and the exact fallback pass ran; the list-free sample pass keeps at most
    # four distances per epilogue thread, so a cluster inside one strip cannot drag the threshold down any more.  (b) thresholds
    # forced far too tight (k' = 1 of a 100 k-row sample): the verification fails for about half of the queries and the gated
    # exact fallback - a second dense pass, thresholds of the short queries at infinity - runs.
    try:
        nadv = 64
        qf_a, qb_a = qf_d[1].clone(), qb_d[1].clone()
        save = index.read_rows(L.ROWS_CODES, 0, 2048)
        dup = qb_a[:nadv].cpu().numpy().repeat(32, axis=0)
        index.write_rows(L.ROWS_CODES, 0, dup)
        lab = torch.empty((NQ, K), dtype=torch.int64, device=dev)
        ham = torch.empty((NQ, K), dtype=torch.int32, device=dev)
        sb = torch.empty((NQ, K), dtype=torch.float64, device=dev)
        sc = torch.empty((NQ, K), dtype=torch.float64, device=dev)
        cnt = torch.empty((NQ,), dtype=torch.int32, device=dev)

        def step(qf_, qb_):
            return lambda: L.check(lib.vrq_index_search3(index._h, NQ, L.ptr(qf_), L.ptr(qb_), K, BO, IO, L.ptr(lab), L.ptr(ham),
                                                         L.ptr(sb), L.ptr(sc), L.ptr(cnt)))

        s_adv = _timed(env, step(qf_a, qb_a), 3)
        d_h, _ = index.search(qb_a.cpu().numpy(), K * BO)  # Phase I of the same batch: every duplicate must be found, at distance 0
        ok = bool((d_h[:nadv, :32] == 0).all() and (d_h[:nadv, 32] > 0).all() and (d_h[:, -1] < 1024).all()) and bool((cnt == K).all().item())
        index.write_rows(L.ROWS_CODES, 0, save)
        saved = {k_: os.environ.get(k_) for k_ in ("VRQ_MMA_SAMPLE_K", "VRQ_MMA_SAFETY")}
        os.environ["VRQ_MMA_SAMPLE_K"], os.environ["VRQ_MMA_SAFETY"] = "1", "1"
        s_fb = _timed(env, step(qf_d[2], qb_d[2]), 3)
        ok_fb = bool((cnt == K).all().item())
        for k_, v in saved.items():
            if v is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v
        out["adversarial"] = {"clustered_duplicates": {"what": f"{nadv} of the 1024 queries have 32 exact duplicates inside the first sampled tiles",
                                                       "ms_per_step": s_adv * 1e3, "ratio": s_adv * 1e3 / ms_per_step,
                                                       "results_complete_and_duplicates_found": ok},
                              "forced_fallback": {"what": "thresholds from k'=1 of a 100k-row sample (VRQ_MMA_SAMPLE_K=1, VRQ_MMA_SAFETY=1): about half of the "
                                                          "queries come up short in the dense pass, the gated exact fallback pass runs for them",
                                                  "ms_per_step": s_fb * 1e3, "ratio": s_fb * 1e3 / ms_per_step, "results_complete": ok_fb},
                              "normal_ms_per_step": ms_per_step}
    except Exception as e:
        out["adversarial"] = {"error": repr(e)}
    return out


def encode_legs(env, hbm_peak, peak_src):
    """BASELINE config 2: global-limit encode of 10 M x 1024 float32 rows resident in HBM (41 GB in)."""
    torch, L, lib, ctx, dev = env.torch, env.L, env.lib, env.ctx, env.dev
    out = {}
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
        s = _timed(env, fn, 5)
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
    tr, tr_src = traffic("encode_int8_global_10M")
    out["roofline_encode"] = {"kernel": "encode1024_ring_kernel<INT8_GLOBAL, ubinary fused> (cp.async ring of rows per warp, magic-number rounding)",
                              "bound": "hbm", "unit": "GB/s",
                              "achieved": enc["int8_global+ubinary"]["GB/s"], "peak": hbm_peak,
                              "frac": enc["int8_global+ubinary"]["frac"], "peak_source": peak_src, "rows": n_enc, "traffic": tr,
                              "traffic_source": tr_src, "algorithmic_bytes": n_enc * 5248,
                              "workload": "cfg2: 10M x 1024 float32 rows resident in HBM", "all_codecs": enc}
    del x, ub, q8, q16, lo, hi
    torch.cuda.empty_cache()
    return out


def cfg1_leg(env):
    """BASELINE.json configs[0]: VectorDBInt8 per-document int8 quantise + search over 10 k synthetic 1024-d embeddings,
    k=10, binary_oversample=10, 100 queries.  CPU side = the reference's flow restated literally (VectorDBInt8.py:148-242):
    per document quantise + _to_binary (NumPy, one vector at a time), per query the faiss-equivalent scan + the
    per-candidate dequantise / np.dot loop.  GPU side = the drop-in class: add_documents in batches of 64 and search() one
    query at a time (the reference's call pattern), and the bulk entry points add_embeddings / search_batch."""
    V = env.V
    n, nq, k, bo = 10_000, 100, 10, 10
    res = {"workload": "cfg1: VectorDBInt8 per-document int8 quantise + 2-phase search, 10k synthetic 1024-d rows, 100 queries, k=10"}
    try:
        from oracle import oracle_c as oc
        from oracle import vrq_oracle as o
        cores = oc.use_all_cores()
        x = oc.synth_f32(1, 0, n, D, True)
        qx = oc.synth_f32(2, 0, nq, D, True)
        # ---- CPU, literal per-document loop
        t0 = time.perf_counter()
        q8 = np.empty((n, D), np.int8)
        lo = np.empty(n, np.float32)
        hi = np.empty(n, np.float32)
        codes = np.empty((n, D // 8), np.uint8)
        for i in range(n):
            q8[i], lo[i], hi[i] = o.quantize_int8_perdoc_one(x[i])
            codes[i] = o.to_binary_one(x[i])
        t_add_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        cpu_top = []
        for qi in range(nq):
            qb = o.to_binary_one(qx[qi])[None]
            dist, pos = oc.hamming_topk(codes, qb, k * bo)
            hits = []
            for p in pos[0]:
                if p == -1:
                    continue
                emb = o.dequantize_int8_perdoc_one(q8[p], lo[p], hi[p])
                hits.append((int(p), float(np.dot(qx[qi], emb))))
            hits.sort(key=lambda h_: h_[1], reverse=True)
            cpu_top.append([h_[0] for h_ in hits[:k]])
        t_search_cpu = time.perf_counter() - t0
        # ---- GPU, the class
        table = {f"doc {i}": i for i in range(n)}
        table.update({f"query {i}": n + i for i in range(nq)})
        allx = np.concatenate([x, qx])
        docs = [f"doc {i}" for i in range(n)]
        with tempfile.TemporaryDirectory() as tmp:
            # a throw-away database first: the timed one does not pay the process's first encode launch / allocation
            V.VectorDBInt8(os.path.join(tmp, "w"), embedder=lambda texts: allx[[table[t] for t in texts]], ctx=env.ctx).add_documents(
                list(range(256)), docs[:256], batch_size=64, save=False)
            db = V.VectorDBInt8(os.path.join(tmp, "a"), embedder=lambda texts: allx[[table[t] for t in texts]], ctx=env.ctx)
            t0 = time.perf_counter()
            db.add_documents(list(range(n)), docs, batch_size=64, save=False)
            env.ctx.sync()
            t_add_gpu = time.perf_counter() - t0
            t0 = time.perf_counter()
            gpu_top = [[h_["doc_id"] for h_ in db.search(f"query {qi}", k=k, binary_oversample=bo)] for qi in range(nq)]
            t_search_gpu = time.perf_counter() - t0
            db2 = V.VectorDBInt8(os.path.join(tmp, "b"), embedder=lambda texts: allx[[table[t] for t in texts]], ctx=env.ctx)
            t0 = time.perf_counter()
            db2.add_embeddings(list(range(n)), x, docs, keep_float=False)
            env.ctx.sync()
            t_bulk_add = time.perf_counter() - t0
            t0 = time.perf_counter()
            labels, scores, cnt = db2.search_batch(qx, k, bo)
            t_bulk_search = time.perf_counter() - t0
            codes_equal = bool(np.array_equal(db2.index.read_rows(env.L.ROWS_CODES, 0, n), codes) and
                               np.array_equal(db2.index.read_rows(env.L.ROWS_PAYLOAD, 0, n), q8))
        same = sum(1 for a, b in zip(cpu_top, gpu_top) if a == b)
        same_bulk = sum(1 for a, b in zip(cpu_top, labels.tolist()) if a == b)
        res.update({"cpu": {"cores": cores, "add_docs_per_s": n / t_add_cpu, "search_queries_per_s": nq / t_search_cpu,
                            "what": "literal per-document NumPy quantise + _to_binary; per query C Hamming scan + per-candidate dequantise/np.dot loop"},
                    "gpu_class_api": {"add_docs_per_s": n / t_add_gpu, "search_queries_per_s": nq / t_search_gpu,
                                      "what": "VectorDBInt8.add_documents(batch_size=64) + search() one query at a time, host buffers through the C ABI "
                                              "(document store, duplicate check and index growth included; a 256-document warm-up database "
                                              "comes first)"},
                    "gpu_bulk_api": {"add_docs_per_s": n / t_bulk_add, "search_queries_per_s": nq / t_bulk_search,
                                     "what": "add_embeddings (one encode call) + search_batch (one call for the 100 queries)"},
                    "codes_and_int8_bit_exact_vs_cpu": codes_equal,
                    "queries_with_identical_top10": {"class_api": same, "bulk_api": same_bulk, "of": nq,
                                                     "note": "float32 dot on the CPU vs float64-accumulated dot rounded to float32 on the GPU: "
                                                             "ranks of near-tied scores may swap (tests compare with the 1e-5 tolerance)"}})
    except Exception as e:
        res["error"] = repr(e)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_PER_GPU, help="rows per GPU (default: the 100M of the headline config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the encode / stream-scan / cfg1 / cfg5 legs")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 1-billion-row leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity probe")
    args = ap.parse_args()
    args.warmup_actual = max(3, args.warmup) if args.impl == "b200" else args.warmup  # timing rule: >= 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
