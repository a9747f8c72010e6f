"""GPU parity tests for the device-resident binary index and the fused searches (through the C ABI)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle_c as oc  # noqa: E402
from oracle import vrq_oracle as o  # noqa: E402


@pytest.fixture(scope="module")
def V():
    import vectorragquantization_b200 as v
    return v


def check_topk(index, codes, q, k, ids=None):
    dist, labels = index.search(q, k)
    rd, rp = oc.hamming_topk(codes, q, k)
    assert np.array_equal(dist, rd)
    rl = rp.copy()
    if ids is not None:
        rl = np.where(rp >= 0, ids[np.clip(rp, 0, None)], -1)
    assert np.array_equal(labels, rl)


@pytest.mark.parametrize("n,nq,k", [(1, 1, 1), (1000, 1, 100), (1000, 3, 1000), (999, 2, 1500), (70000, 5, 10),
                                    (70000, 37, 100), (300000, 2, 1000), (300000, 300, 50), (1 << 20, 1, 1000)])
def test_hamming_topk_random(V, n, nq, k):
    rng = np.random.default_rng(n + nq + k)
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ids = rng.permutation(n).astype(np.int64) * 7 + 3
    ix = V.IndexBinaryIDMap2(V.IndexBinaryFlat(1024))
    ix.add_with_ids(codes, ids)
    assert ix.ntotal == n
    check_topk(ix, codes, q, k, ids)


@pytest.mark.parametrize("n,nq,k", [(3_000_000, 4, 1000), (1_500_000, 300, 100)])
def test_hamming_topk_prefix_pass(V, n, nq, k):
    """Large enough that the prefix pass seeds the thresholds of the main pass (scan.cu: topk_batch)."""
    codes, _ = oc.synth_codes_int8(61, 0, n, want_int8=False)
    qx = oc.synth_f32(62, 0, nq)
    q = o.synth_ubinary_from_f32(qx)
    q[0] = codes[n - 5]  # an exact hit near the end of the last strip
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    check_topk(ix, codes, q, k)


def test_hamming_topk_massive_ties(V):
    """Thousands of codes at the k-th distance: ties must resolve by ascending position, across strips."""
    rng = np.random.default_rng(42)
    n = 400000
    base = rng.integers(0, 256, (16, 128), dtype=np.uint8)
    codes = base[rng.integers(0, 16, n)]  # only 16 distinct codes -> huge tie groups
    q = np.concatenate([base[:3], rng.integers(0, 256, (2, 128), dtype=np.uint8)])
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (1, 100, 1000, 4096):
        check_topk(ix, codes, q, k)
    # all-equal database: every candidate ties
    codes[:] = base[0]
    ix2 = V.BinaryIndex(1024)
    ix2.add_with_ids(codes, np.arange(n))
    check_topk(ix2, codes, q[:2], 1000)


def test_hamming_topk_clustered_and_sorted(V):
    """Adversarial order for a streaming threshold: the database is sorted by DEcreasing distance to the query, so
    every later row beats the current threshold (worst case for list compaction)."""
    rng = np.random.default_rng(4)
    n = 200000
    q = rng.integers(0, 256, (1, 128), dtype=np.uint8)
    flips = np.sort(rng.integers(0, 1024, n))[::-1]
    bits = np.unpackbits(np.repeat(q, n, 0), axis=1)
    mask = np.arange(1024)[None, :] < flips[:, None]
    codes = np.packbits(bits ^ mask, axis=1)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    check_topk(ix, codes, q, 1000)
    check_topk(ix, codes, np.concatenate([q, ~q]), 77)


@pytest.mark.parametrize("d", [256, 384, 2048])
def test_hamming_topk_other_code_sizes(V, d):
    rng = np.random.default_rng(d)
    n = 50000
    codes = rng.integers(0, 256, (n, d // 8), dtype=np.uint8)
    q = rng.integers(0, 256, (9, d // 8), dtype=np.uint8)
    ix = V.BinaryIndex(d)
    ix.add_with_ids(codes, np.arange(n))
    check_topk(ix, codes, q, 200)


def test_index_faiss_surface(V, tmp_path, golden_dbs):
    codes = golden_dbs["db_cohere_enhanced.codes"]
    ids = np.arange(1000, dtype=np.int64)
    ix = V.IndexBinaryIDMap2(V.IndexBinaryFlat(1024))
    for s in range(0, 1000, 64):  # the reference adds in batches of 64 (:195)
        ix.add_with_ids(codes[s:s + 64], ids[s:s + 64])
    assert ix.ntotal == 1000
    assert np.array_equal(ix.reconstruct(851), codes[851])
    # byte-compatible index.bin (KAT-5)
    p = os.path.join(tmp_path, "index.bin")
    V.write_index_binary(ix, p)
    raw = open(p, "rb").read()
    assert raw == o.write_index_binary_bytes(1024, codes, ids)
    assert raw[:58].hex() == golden_dbs.headers["db_cohere_enhanced"]["index_header_hex"]
    ix2 = V.read_index_binary(p)
    assert ix2.ntotal == 1000 and ix2.d == 1024
    q = codes[[5, 77]]
    assert all(np.array_equal(a, b) for a, b in zip(ix.search(q, 50), ix2.search(q, 50)))
    # remove_ids keeps order (faiss compacts) and reconstruct follows
    assert ix2.remove_ids(np.array([5, 6, 999, 123456])) == 3
    keep = np.setdiff1d(ids, [5, 6, 999])
    assert ix2.ntotal == 997
    d1, l1 = ix2.search(q, 20)
    rd, rp = oc.hamming_topk(codes[keep], q, 20)
    assert np.array_equal(d1, rd) and np.array_equal(l1, keep[rp])
    with pytest.raises(V.VrqError):
        ix2.reconstruct(5)
    # duplicate ids: last added wins for reconstruct (IDMap2)
    ix2.add_with_ids(codes[:1] ^ 0xFF, np.array([7]))
    assert np.array_equal(ix2.reconstruct(7), codes[0] ^ 0xFF)
    # k > ntotal pads with (INT32_MAX, -1)
    small = V.BinaryIndex(1024)
    small.add_with_ids(codes[:3], ids[:3])
    d3, l3 = small.search(q, 5)
    assert np.all(l3[:, 3:] == -1) and np.all(d3[:, 3:] == 2147483647)


def test_search3_matches_reference_flow(V):
    """CohereEnhancedVectorDB.search phases I-III on synthetic Cohere-like data vs the oracle's literal restatement."""
    from vectorragquantization_b200 import _lib as L
    n, nq, k, bo, io = 30000, 12, 10, 10, 3
    x = oc.synth_f32(31, 0, n)
    codes, i8 = oc.synth_codes_int8(31, 0, n)
    ids = np.arange(n, dtype=np.int64) + 1000
    # queries: perturbed database rows, so there are true neighbours (and a few exact duplicates -> ties)
    qf = (x[np.arange(nq) * 997] + oc.synth_f32(32, 0, nq) * np.float32(0.5)).astype(np.float32)
    qf[3] = x[2991]
    qb = o.synth_ubinary_from_f32(qf)
    ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_with_ids(codes, ids, payload=i8)
    labels, ham, sb, sc, cnt = ix.search3(qf, qb, k, bo, io)
    for qi in range(nq):
        ref = o.search3(codes, ids, i8, qf[qi], qb[qi], k, bo, io)
        assert cnt[qi] == len(ref) == k
        assert [h["doc_id"] for h in ref] == labels[qi].tolist()
        assert [h["score_hamming"] for h in ref] == ham[qi].tolist()
        rb = np.array([h["score_binary"] for h in ref])
        rc = np.array([h["score_cosine"] for h in ref])
        assert np.all(np.abs(sb[qi] - rb) <= 1e-5 * np.abs(rb) + 1e-12)
        fl = o.rescore_int8cos_absfloor(qf[qi], i8[labels[qi] - 1000])
        assert np.all(np.abs(sc[qi] - rc) <= 1e-5 * np.abs(rc) + fl)
    # k * binary_oversample > ntotal: clamp (:267)
    small = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    small.add_with_ids(codes[:40], ids[:40], payload=i8[:40])
    labels, ham, sb, sc, cnt = small.search3(qf[:2], qb[:2], 10, 10, 3)
    for qi in range(2):
        ref = o.search3(codes[:40], ids[:40], i8[:40], qf[qi], qb[qi], 10, 10, 3)
        assert cnt[qi] == len(ref) == 10 and [h["doc_id"] for h in ref] == labels[qi].tolist()
    labels, ham, sb, sc, cnt = small.search3(qf[:2], qb[:2], 50, 10, 3)
    assert np.all(cnt == 40) and np.all(labels[:, 40:] == -1)
    # empty index -> no results (:247-249)
    empty = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    assert np.all(empty.search3(qf[:2], qb[:2], 10)[4] == 0)


def test_search2_all_payload_kinds(V):
    from vectorragquantization_b200 import _lib as L
    n, nq, k, bo = 8000, 6, 10, 10
    x = oc.synth_f32(41, 0, n, row_scale=True)
    ub = o.to_binary_f32(x)
    ids = np.arange(n, dtype=np.int64)
    qf = (x[np.arange(nq) * 501] * np.float32(0.9) + oc.synth_f32(42, 0, nq) * np.float32(0.3)).astype(np.float32)
    qb = o.to_binary_f32(qf)
    q8, lo, hi = o.quantize_int8_perdoc(x)
    p4, lo4, hi4 = o.quantize_int4(x)
    cases = [
        (L.PAYLOAD_INT8_PERDOC, 0.0, q8, np.stack([lo, hi], 1), lambda p: o.dequantize_int8_perdoc(q8[p], lo[p], hi[p])),
        (L.PAYLOAD_INT8_GLOBAL, 0.3, o.quantize_int8_global(x, 0.3), None,
         lambda p: o.dequantize_int8_global(o.quantize_int8_global(x[p], 0.3), 0.3)),
        (L.PAYLOAD_INT16_GLOBAL, 1.0, o.quantize_int16_global(x, 1.0), None,
         lambda p: o.dequantize_int16_global(o.quantize_int16_global(x[p], 1.0), 1.0)),
        (L.PAYLOAD_INT4_PERDOC, 0.0, p4, np.stack([lo4, hi4], 1), lambda p: o.dequantize_int4_perdoc(p4[p], 1024, lo4[p], hi4[p])),
        (L.PAYLOAD_INT4_GLOBAL, 0.18, p4, None, lambda p: o.dequantize_int4_global(p4[p], 1024, 0.18)),
        (L.PAYLOAD_F32, 0.0, x, None, lambda p: x[p]),
    ]
    for kind, lim, payload, aux, deq in cases:
        ix = V.BinaryIndex(1024, payload_kind=kind, global_limit=lim)
        ix.add_with_ids(ub, ids, payload=payload, aux=aux)
        labels, score, cnt = ix.search2(qf, qb, k, bo)
        # q_ubin = NULL: the library derives query_bin = _to_binary(query float) on the device (VectorDBInt8.py:213)
        l0, s0, c0 = ix.search2(qf, None, k, bo)
        assert np.array_equal(l0, labels) and np.array_equal(s0, score) and np.array_equal(c0, cnt), kind
        l1, s1, c1 = ix.search2(qf[0], None, k, bo)  # one query per call, as the classes search
        assert np.array_equal(l1[0], labels[0]) and np.array_equal(s1[0], score[0])
        for qi in range(nq):
            ref = o.search2(ub, ids, deq, qf[qi], qb[qi], k, bo)
            assert cnt[qi] == k
            rs = np.array([h["score"] for h in ref], np.float64)
            # float32 dot: tolerance 1e-5 relative + the float32 accumulation floor of the reference's sdot
            floor = 4 * 2.0 ** -24 * np.abs(deq(labels[qi]).astype(np.float64)) @ np.abs(qf[qi].astype(np.float64))
            assert np.all(np.abs(score[qi] - rs) <= 1e-5 * np.abs(rs) + floor + 1e-9), kind
            # ids equal wherever the reference's own scores are separated by more than the tolerance
            same = labels[qi] == np.array([h["doc_id"] for h in ref])
            gap_ok = np.ones(k, bool)
            gap_ok[:-1] &= np.abs(np.diff(rs)) > 2 * (1e-5 * np.abs(rs[:-1]) + floor[:-1])
            gap_ok[1:] &= np.abs(np.diff(rs)) > 2 * (1e-5 * np.abs(rs[1:]) + floor[1:])
            assert np.all(same | ~gap_ok), kind


def test_merge3_equals_single_index(V):
    """Shard the database by rows into 4 indexes on one GPU, run search3_local per shard into the layout an
    NCCL all-gather produces, merge3 -> must equal search3 on the unsharded index bit for bit."""
    import torch
    from vectorragquantization_b200 import _lib as L
    n, nq, k, bo, io, W = 40000, 9, 10, 10, 3, 4
    codes, i8 = oc.synth_codes_int8(51, 0, n)
    ids = np.arange(n, dtype=np.int64) * 2 + 1
    x = oc.synth_f32(51, 0, n)
    qf = (x[np.arange(nq) * 1234] + oc.synth_f32(52, 0, nq) * np.float32(0.7)).astype(np.float32)
    qb = o.synth_ubinary_from_f32(qf)
    full = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    full.add_with_ids(codes, ids, payload=i8)
    want = full.search3(qf, qb, k, bo, io)
    ctx = full.ctx
    dev = torch.device("cuda", ctx.device)
    bk = k * bo
    keys = torch.empty((W, nq, bk), dtype=torch.int64, device=dev)
    labels = torch.empty((W, nq, bk), dtype=torch.int64, device=dev)
    sbin = torch.empty((W, nq, bk), dtype=torch.float64, device=dev)
    scos = torch.empty((W, nq, bk), dtype=torch.float64, device=dev)
    qf_d = torch.from_numpy(qf).to(dev)
    qb_d = torch.from_numpy(qb).to(dev)
    torch.cuda.synchronize()
    shards = []
    for w in range(W):
        a, b = w * n // W, (w + 1) * n // W
        sh = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
        sh.add_with_ids(codes[a:b], ids[a:b], payload=i8[a:b])
        sh.search3_local_into(qf_d, qb_d, nq, bk, a, keys[w], labels[w], sbin[w], scos[w])
        shards.append(sh)
    ctx.sync()
    out = [np.empty((nq, k), np.int64), np.empty((nq, k), np.int32), np.empty((nq, k), np.float64),
           np.empty((nq, k), np.float64), np.empty(nq, np.int32)]
    lib = L.load()
    hk, hl, hb, hc = keys.cpu().numpy(), labels.cpu().numpy(), sbin.cpu().numpy(), scos.cpu().numpy()
    L.check(lib.vrq_merge3(ctx.handle, W, nq, bk, 0, L.ptr(hk), L.ptr(hl), L.ptr(hb), L.ptr(hc), k, k * io,
                           *[L.ptr(a) for a in out]))
    # and once more entirely on device pointers (the path sharded.py uses)
    dout = [torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.int32, device=dev),
            torch.empty((nq, k), dtype=torch.float64, device=dev), torch.empty((nq, k), dtype=torch.float64, device=dev),
            torch.empty(nq, dtype=torch.int32, device=dev)]
    torch.cuda.synchronize()
    L.check(lib.vrq_merge3(ctx.handle, W, nq, bk, 0, L.ptr(keys), L.ptr(labels), L.ptr(sbin), L.ptr(scos), k, k * io,
                           *[L.ptr(a) for a in dout]))
    ctx.sync()
    for a, b in zip(want, dout):
        assert np.array_equal(a, b.cpu().numpy())
    for a, b in zip(want, out):
        assert np.array_equal(a, b)


def test_limits_and_errors(V):
    """Argument checking at the C ABI: errors are loud VrqErrors with the documented codes, never silent fallbacks."""
    from vectorragquantization_b200 import _lib as L
    from vectorragquantization_b200 import kernels as K
    with pytest.raises(V.VrqError) as e:
        V.BinaryIndex(1027)  # faiss's own requirement: d % 8 == 0
    assert e.value.code == L.ERR_ARG
    ix = V.BinaryIndex(1024)
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 256, (5000, 128), dtype=np.uint8)
    ix.add_with_ids(codes, np.arange(5000))
    with pytest.raises(V.VrqError) as e:
        ix.search(codes[:1], 20000)  # k > 16384
    assert e.value.code == L.ERR_UNSUPPORTED
    with pytest.raises(V.VrqError):
        ix.search(codes[:1], 0)
    with pytest.raises(V.VrqError) as e:
        ix.search3(np.zeros((1, 1024), np.float32), codes[:1], 10)  # no int8 payload
    assert e.value.code == L.ERR_STATE
    with pytest.raises(V.VrqError):
        K.quantize_int8_global(np.zeros((2, 1020), np.float32), 0.3)  # d % 8 != 0
    with pytest.raises(V.VrqError):
        K.quantize_int8_global(np.zeros((2, 1024), np.float32), 0.0)  # limit must be > 0
    with pytest.raises(V.VrqError):
        ix.add_with_ids(codes[:2], np.arange(2), payload=np.zeros((2, 1024), np.int8))  # payload without a payload kind
    # maximum supported k, and a query count that spans several internal batches
    d, l = ix.search(codes[:3], 4096)
    rd, rp = oc.hamming_topk(codes, codes[:3], 4096)
    assert np.array_equal(d, rd) and np.array_equal(l, rp)
    q = rng.integers(0, 256, (1100, 128), dtype=np.uint8)
    d, l = ix.search(q, 7)
    rd, rp = oc.hamming_topk(codes, q, 7)
    assert np.array_equal(d, rd) and np.array_equal(l, rp)


@pytest.mark.parametrize("n,nq,k", [(30000, 3, 5000), (30000, 40, 10000), (9000, 2, 16384), (400000, 5, 10000)])
def test_hamming_topk_beyond_one_pass(V, n, nq, k):
    """k above what one scan pass keeps per list (4096): the ranking is produced in chunks, each an exact scan above the last
    key of the chunk before it (faiss has no limit on k; CohereEnhancedVectorDB.py:267 with k=1000 asks for 10000).  Includes
    k > ntotal (padding) and heavy ties across chunk boundaries."""
    rng = np.random.default_rng(n + k)
    if n == 9000:  # few distinct codes: tie groups far larger than a chunk
        base = rng.integers(0, 256, (4, 128), dtype=np.uint8)
        codes = base[rng.integers(0, 4, n)]
    else:
        codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    check_topk(ix, codes, q, k)


def test_search3_large_binary_k(V):
    """k * binary_oversample = 10000 (k=1000, oversample 10: a legal call of the reference) through all three phases."""
    from vectorragquantization_b200 import _lib as L
    n, nq, k, bo, io = 40000, 4, 1000, 10, 3
    x = oc.synth_f32(41, 0, n)
    codes, i8 = oc.synth_codes_int8(41, 0, n)
    ids = np.arange(n, dtype=np.int64)
    qf = (x[np.arange(nq) * 997] + oc.synth_f32(42, 0, nq) * np.float32(0.5)).astype(np.float32)
    qb = o.synth_ubinary_from_f32(qf)
    ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_with_ids(codes, ids, payload=i8)
    labels, ham, sb, sc, cnt = ix.search3(qf, qb, k, bo, io)
    for qi in range(nq):
        ref = o.search3(codes, ids, i8, qf[qi], qb[qi], k, bo, io, literal=False)
        assert cnt[qi] == len(ref) == k
        assert [h["doc_id"] for h in ref] == labels[qi].tolist()
        assert [h["score_hamming"] for h in ref] == ham[qi].tolist()
    # the 2-phase classes' search with 12000 phase-I hits
    ix2 = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_F32)
    ix2.add_with_ids(codes, ids, payload=x)
    l2, s2, c2 = ix2.search2(qf[:2], qb[:2], 1200, 10)
    for qi in range(2):
        ref = o.search2(codes, ids, lambda p: x[p], qf[qi], qb[qi], 1200, 10)
        assert c2[qi] == 1200 and len(set(l2[qi].tolist()) ^ {h["doc_id"] for h in ref}) <= 4  # float32 score ties may swap the last ranks


@pytest.mark.parametrize("d", [1000, 136, 8])
def test_code_widths_not_multiple_of_32_bits(V, d):
    """faiss binary indexes only require d % 8 == 0: index, Hamming top-k, both fused searches and the file round trip with
    code rows that are not word-aligned (d = 1000 -> 125-byte codes)."""
    from vectorragquantization_b200 import _lib as L
    rng = np.random.default_rng(d)
    n, nq = 5000, 6
    x = rng.normal(0, 0.05, (n, d)).astype(np.float32)
    codes = np.packbits(x > 0, axis=1)
    i8 = np.clip(np.rint(1259 * x - 0.69), -128, 127).astype(np.int8)
    ids = np.arange(n, dtype=np.int64) * 3
    qf = (x[:nq] + rng.normal(0, 0.02, (nq, d))).astype(np.float32)
    qb = np.packbits(qf > 0, axis=1)
    ix = V.BinaryIndex(d, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_with_ids(codes, ids, payload=i8)
    check_topk(ix, codes, qb, 100 if d > 8 else 20, ids)
    assert np.array_equal(ix.reconstruct(int(ids[77])), codes[77])
    if d >= 136:
        labels, ham, sb, sc, cnt = ix.search3(qf, qb, 10, 10, 3)
        for qi in range(nq):
            ref = o.search3(codes, ids, i8, qf[qi], qb[qi], 10, 10, 3)
            assert [h["doc_id"] for h in ref] == labels[qi].tolist() and [h["score_hamming"] for h in ref] == ham[qi].tolist()
            rb = np.array([h["score_binary"] for h in ref])
            assert np.all(np.abs(sb[qi] - rb) <= 1e-5 * np.abs(rb) + 1e-12)
    ix2 = V.BinaryIndex(d, payload_kind=L.PAYLOAD_F32)
    ix2.add_with_ids(codes, ids, payload=x)
    l2, s2, c2 = ix2.search2(qf, qb, 5, 10)
    for qi in range(nq):
        ref = o.search2(codes, ids, lambda p: x[p], qf[qi], qb[qi], 5, 10)
        assert [h["doc_id"] for h in ref] == l2[qi].tolist()


def test_search3_other_dim(V):
    """d = 2048: generic (non-TMA) scan and the generic rescoring paths."""
    from vectorragquantization_b200 import _lib as L
    n, nq, d = 6000, 5, 2048
    x = o.synth_f32(91, 0, n, d)
    codes, i8 = o.synth_ubinary_from_f32(x), o.synth_int8_from_f32(x)
    ids = np.arange(n, dtype=np.int64)
    qf = (x[:nq] * np.float32(0.8) + o.synth_f32(92, 0, nq, d) * np.float32(0.4)).astype(np.float32)
    qb = o.synth_ubinary_from_f32(qf)
    ix = V.BinaryIndex(d, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_with_ids(codes, ids, payload=i8)
    labels, ham, sb, sc, cnt = ix.search3(qf, qb, 10, 10, 3)
    for qi in range(nq):
        ref = o.search3(codes, ids, i8, qf[qi], qb[qi], 10, 10, 3)
        assert [h["doc_id"] for h in ref] == labels[qi].tolist()
        assert [h["score_hamming"] for h in ref] == ham[qi].tolist()
        rc = np.array([h["score_cosine"] for h in ref])
        fl = o.rescore_int8cos_absfloor(qf[qi], i8[labels[qi]])
        assert np.all(np.abs(sc[qi] - rc) <= 1e-5 * np.abs(rc) + fl)


@pytest.mark.parametrize("explicit_ids", [False, True])
def test_remove_ids_lazy_compaction(V, explicit_ids, tmp_path):
    """remove_ids (CohereEnhancedVectorDB.py:334) records the rows and compacts the device arrays once, in place, before
    the next read: many single-id removals (the reference removes one document per call) cost one pass.  Every
    observable - ntotal, search order incl. ties, reconstruct, payload, index.bin - is what faiss's immediate
    order-preserving compaction gives."""
    from vectorragquantization_b200 import _lib as L
    rng = np.random.default_rng(5)
    n = 300_000
    base = rng.integers(0, 256, (64, 128), dtype=np.uint8)
    codes = base[rng.integers(0, 64, n)]  # few distinct codes: tie order across removed rows matters
    codes[::7] = rng.integers(0, 256, (len(codes[::7]), 128), dtype=np.uint8)
    payload = rng.integers(-128, 128, (n, 1024), dtype=np.int8)
    ids = (np.arange(n, dtype=np.int64) if not explicit_ids else rng.permutation(n).astype(np.int64) * 5 + 1)
    ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_with_ids(codes, ids, payload=payload)
    kill_pos = np.unique(np.concatenate([rng.integers(0, n, 500), [0, 1, 2, n - 1, n - 2, 77777]]))
    removed = 0
    for chunk in np.array_split(ids[kill_pos], 40):  # 40 calls, nothing read in between
        removed += ix.remove_ids(chunk)
    assert removed == len(kill_pos)
    assert ix.remove_ids(ids[kill_pos[:10]]) == 0  # already gone
    assert ix.ntotal == n - len(kill_pos)
    keep = np.setdiff1d(np.arange(n), kill_pos)
    q = np.concatenate([base[:8], codes[kill_pos[:4]], rng.integers(0, 256, (28, 128), dtype=np.uint8)])
    for nq in (3, 40):  # integer-pipe and tensor-core scans
        d, lab = ix.search(q[:nq], 300)
        rd, rp = oc.hamming_topk(codes[keep], q[:nq], 300)
        assert np.array_equal(d, rd) and np.array_equal(lab, ids[keep][rp])
    with pytest.raises(V.VrqError):
        ix.reconstruct(int(ids[kill_pos[3]]))
    assert np.array_equal(ix.reconstruct(int(ids[keep[12345]])), codes[keep[12345]])
    got = ix.get_payload(np.array([0, 1, len(keep) - 1, 4242]), np.int8, 1024)
    got = got[0] if isinstance(got, tuple) else got
    assert np.array_equal(got, payload[keep][[0, 1, len(keep) - 1, 4242]])
    # remove + re-add (the dedupe path of add_documents), then the file is byte-identical to a fresh index of the same rows
    ix.remove_ids(ids[keep[:3]])
    ix.add_with_ids(codes[keep[:3]], ids[keep[:3]], payload=payload[keep[:3]])
    order = np.concatenate([keep[3:], keep[:3]])
    p = str(tmp_path / "index.bin")
    V.write_index_binary(ix, p)
    assert open(p, "rb").read() == o.write_index_binary_bytes(1024, codes[order], ids[order])
    # duplicate ids: IDMap2 removes every row carrying the id
    dup = V.BinaryIndex(1024)
    dup.add_with_ids(codes[:10], np.array([1, 2, 3, 1, 2, 3, 1, 9, 9, 4], dtype=np.int64))
    assert dup.remove_ids(np.array([1, 9])) == 5 and dup.ntotal == 5
    d, lab = dup.search(codes[:1], 5)
    assert sorted(lab[0].tolist()) == [2, 2, 3, 3, 4]
