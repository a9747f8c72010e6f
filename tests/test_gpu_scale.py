"""Scale checks (gpu): sizes where a full CPU oracle pass is too slow for every query are covered with
size-independent properties - sortedness, recomputed distances for the returned rows, shard-merge invariance, a full
oracle scan for two queries, and bit-exact sampled rows for the bulk encoders (device-resident path)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle_c as oc  # noqa: E402
from oracle import vrq_oracle as o  # noqa: E402

N = 8_000_000


@pytest.fixture(scope="module")
def big():
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.reserve(N)
    ix.add_synthetic(77, 0, N, 0)
    return V, L, ix


def test_topk_properties_at_scale(big):
    V, L, ix = big
    nq, k = 48, 1000
    qf = oc.synth_f32(78, 0, nq)
    qb = o.synth_ubinary_from_f32(qf)
    dist, lab = ix.search(qb, k)
    # sorted by (distance, position), no duplicates
    key = dist.astype(np.int64) << 40 | lab
    assert np.all(np.diff(key, axis=1) > 0)
    # recompute the distance of every returned row from the counter-based generator
    for qi in range(0, nq, 7):
        rows = np.stack([oc.synth_codes_int8(77, int(p), 1, want_int8=False)[0][0] for p in lab[qi][::25]])
        assert np.array_equal(o.hamming_distances(rows, qb[qi])[0], dist[qi][::25])
    # full oracle scan for two queries over the whole database
    codes, _ = oc.synth_codes_int8(77, 0, N, want_int8=False)
    rd, rp = oc.hamming_topk(codes, qb[:2], k)
    assert np.array_equal(rd, dist[:2]) and np.array_equal(rp, lab[:2])
    # batch invariance: the same queries inside a different batch size / regime give the same answer
    d1, l1 = ix.search(qb[:1], k)
    d9, l9 = ix.search(qb[:9], k)
    assert np.array_equal(d1, dist[:1]) and np.array_equal(l1, lab[:1])
    assert np.array_equal(d9, dist[:9]) and np.array_equal(l9, lab[:9])


def test_search3_shard_invariance_at_scale(big):
    """3-phase search over the whole index == merge of the two half shards (device-pointer path + vrq_merge3)."""
    import torch
    V, L, ix = big
    nq, k, bo, io = 32, 100, 10, 3
    qf = oc.synth_f32(79, 0, nq)
    qb = o.synth_ubinary_from_f32(qf)
    want = ix.search3(qf, qb, k, bo, io)
    halves = []
    for a, b in ((0, N // 2), (N // 2, N)):
        h = V.BinaryIndex(1024, ctx=ix.ctx, payload_kind=L.PAYLOAD_INT8_RAW)
        h.add_synthetic(77, a, b - a, a)
        halves.append((h, a))
    dev = torch.device("cuda", ix.ctx.device)
    bk = k * bo
    packed = torch.empty((2, 4, nq, bk), dtype=torch.int64, device=dev)
    qf_d, qb_d = torch.from_numpy(qf).to(dev), torch.from_numpy(qb).to(dev)
    torch.cuda.synchronize()
    for w, (h, a) in enumerate(halves):
        h.search3_local_into(qf_d, qb_d, nq, bk, a, packed[w, 0], packed[w, 1], packed[w, 2], packed[w, 3])
    out = [torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.int32, device=dev),
           torch.empty((nq, k), dtype=torch.float64, device=dev), torch.empty((nq, k), dtype=torch.float64, device=dev),
           torch.empty(nq, dtype=torch.int32, device=dev)]
    L.check(L.load().vrq_merge3(ix.ctx.handle, 2, nq, bk, 4 * nq * bk, L.ptr(packed[0, 0]), L.ptr(packed[0, 1]), L.ptr(packed[0, 2]),
                                L.ptr(packed[0, 3]), k, k * io, *[L.ptr(t) for t in out]))
    ix.ctx.sync()
    for a, b in zip(want, out):
        assert np.array_equal(a, b.cpu().numpy())
    # scores of the winners recomputed by the oracle from regenerated rows
    for qi in (0, 17):
        ids = want[0][qi]
        rows = [oc.synth_codes_int8(77, int(p), 1) for p in ids]
        sb = o.rescore_binary(qf[qi], np.stack([r[0][0] for r in rows]))
        sc = o.rescore_int8cos(qf[qi], np.stack([r[1][0] for r in rows]))
        fl = o.rescore_int8cos_absfloor(qf[qi], np.stack([r[1][0] for r in rows]))
        assert np.all(np.abs(want[2][qi] - sb) <= 1e-5 * np.abs(sb) + 1e-12)
        assert np.all(np.abs(want[3][qi] - sc) <= 1e-5 * np.abs(sc) + fl)
        assert np.all(np.diff(want[3][qi]) <= 0)  # final order: score_cosine descending


def test_bulk_encode_device_path_at_scale():
    """4 M rows (16 GB of float32) generated and encoded entirely on the device; sampled rows bit-exact vs the oracle,
    global invariants on the whole output."""
    import torch
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    n = 4_000_000
    ctx = V.default_context()
    lib = L.load()
    dev = torch.device("cuda", ctx.device)
    ctx.reset_stream()
    x = torch.empty((n, 1024), dtype=torch.float32, device=dev)
    L.check(lib.vrq_synth_f32(ctx.handle, 5, 0, n, 1024, 1, L.ptr(x)))
    q8 = torch.empty((n, 1024), dtype=torch.int8, device=dev)
    q4 = torch.empty((n, 512), dtype=torch.int8, device=dev)
    ub = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    ub2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    lo = torch.empty(n, dtype=torch.float64, device=dev)
    hi = torch.empty(n, dtype=torch.float64, device=dev)
    L.check(lib.vrq_quantize_int8_global(ctx.handle, L.ptr(x), n, 1024, 0.3, L.ptr(q8), L.ptr(ub)))
    L.check(lib.vrq_quantize_int4(ctx.handle, L.ptr(x), n, 1024, L.ptr(q4), L.ptr(lo), L.ptr(hi), L.ptr(ub2)))
    ctx.sync()
    assert torch.equal(ub, ub2)  # the fused code is the same whichever codec it rides with
    assert int(q8.abs().max()) <= 127
    rng = np.random.default_rng(0)
    rows = np.concatenate([np.arange(0, 4096), np.arange(n - 4096, n), rng.integers(0, n, 8192)])
    idx = torch.from_numpy(rows).to(dev)
    xs = x[idx].cpu().numpy()
    assert np.array_equal(xs[:4096], oc.synth_f32(5, 0, 4096, 1024, True))
    assert np.array_equal(q8[idx].cpu().numpy(), oc.quantize_int8_global(xs, 0.3))
    p4, l4, h4 = oc.quantize_int4(xs)
    assert np.array_equal(q4[idx].cpu().numpy(), p4)
    assert np.array_equal(lo[idx].cpu().numpy(), l4) and np.array_equal(hi[idx].cpu().numpy(), h4)
    assert np.array_equal(ub[idx].cpu().numpy(), oc.to_binary_f32(xs))
    # round trip: |dequant(quant(x)) - clip(x)| <= half a step (+ float32 rounding) everywhere
    deq = q8[idx].float() * np.float32(0.3 / 127.0)
    err = (deq - x[idx].clamp(-0.3, 0.3)).abs().max().item()
    assert err <= 0.5 * 0.3 / 127.0 * 1.0001


def test_phase1_at_full_baseline_size(monkeypatch):
    """BASELINE.json configs[2] at its real size: 100 M x 1024-bit codes, 1024 queries, Phase-I k = 1000.  The tensor-core
    scan (+-1 x {0,1} e2m1 contraction, CTA pairs, sampled thresholds) and the integer-pipe scan (XOR + carry-save POPC,
    exact prefix pass) share nothing but the merge tree; they must return identical (distance, position) lists, which must
    be sorted, duplicate-free, and carry distances that recompute from the counter-based generator."""
    import torch
    import vectorragquantization_b200 as V
    if torch.cuda.mem_get_info()[0] < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    n, nq, k = 100_000_000, 1024, 1000
    ix = V.BinaryIndex(1024)
    ix.reserve(n)
    for off in range(0, n, 10_000_000):
        ix.add_synthetic(1, off, 10_000_000, off)
    qf = oc.synth_f32(2, 0, nq)
    qb = o.synth_ubinary_from_f32(qf)
    for kk in ("VRQ_SCAN_MMA", "VRQ_MMA_KIND", "VRQ_MMA_PAIR", "VRQ_MMA_SAFETY", "VRQ_MMA_SAMPLE_K"):
        monkeypatch.delenv(kk, raising=False)
    dist, lab = ix.search(qb, k)  # tensor cores
    key = dist.astype(np.int64) << 40 | lab
    assert np.all(np.diff(key, axis=1) > 0) and lab.min() >= 0 and lab.max() < n
    monkeypatch.setenv("VRQ_SCAN_MMA", "0")
    d0, l0 = ix.search(qb, k)  # integer pipes
    assert np.array_equal(dist, d0) and np.array_equal(lab, l0)
    monkeypatch.setenv("VRQ_SCAN_MMA", "1")
    monkeypatch.setenv("VRQ_MMA_KIND", "8")
    d8, l8 = ix.search(qb[:256], k)  # int8 operand kind
    assert np.array_equal(dist[:256], d8) and np.array_equal(lab[:256], l8)
    for qi in (0, 511, 1023):
        rows = np.stack([oc.synth_codes_int8(1, int(p), 1, want_int8=False)[0][0] for p in lab[qi][::50]])
        assert np.array_equal(o.hamming_distances(rows, qb[qi])[0], dist[qi][::50])
    ix.close()
