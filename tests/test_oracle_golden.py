"""Pin the CPU oracle against the reference's own artefacts and against outputs of
the reference's own static methods (tests/golden/reference_static_methods.npz, made by make_golden.py).
CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import vrq_oracle as o

LIMITS = (0.18, 0.3, 1.0)


# ---------------------------------------------------------------- reference static methods ----


def test_int8_perdoc_matches_reference(golden_static):
    g = golden_static
    q, lo, hi = o.quantize_int8_perdoc(g["x"])
    assert np.array_equal(q, g["int8_perdoc.q"])
    assert np.array_equal(np.stack([lo, hi], 1), g["int8_perdoc.min_max"])
    deq = o.dequantize_int8_perdoc(q, lo, hi)
    assert np.array_equal(deq, g["int8_perdoc.deq"])
    q2, lo2, hi2 = oc.quantize_int8_perdoc(g["x"])
    assert np.array_equal(q2, q) and np.array_equal(lo2, lo) and np.array_equal(hi2, hi)


@pytest.mark.parametrize("lim", LIMITS)
def test_global_int8_int16_match_reference(golden_static, lim):
    g = golden_static
    assert np.array_equal(o.quantize_int8_global(g["x"], lim), g[f"int8_global.q.{lim}"])
    assert np.array_equal(o.quantize_int16_global(g["x"], lim), g[f"int16_global.q.{lim}"])
    assert np.array_equal(oc.quantize_int8_global(g["x"], lim), g[f"int8_global.q.{lim}"])
    assert np.array_equal(oc.quantize_int16_global(g["x"], lim), g[f"int16_global.q.{lim}"])
    if f"int8_global.deq.{lim}" in g.files:
        assert np.array_equal(o.dequantize_int8_global(g[f"int8_global.q.{lim}"], lim), g[f"int8_global.deq.{lim}"])
    if f"int16_global.deq.{lim}" in g.files:
        assert np.array_equal(o.dequantize_int16_global(g[f"int16_global.q.{lim}"], lim),
                              g[f"int16_global.deq.{lim}"])


def test_int4_matches_reference(golden_static):
    g = golden_static
    p, lo, hi = o.quantize_int4(g["x"])
    assert np.array_equal(p, g["int4.q"])
    assert np.array_equal(np.stack([lo, hi], 1), g["int4.min_max"])
    p2, lo2, hi2 = oc.quantize_int4(g["x"])
    assert np.array_equal(p2, p) and np.array_equal(lo2, lo) and np.array_equal(hi2, hi)


def test_int4_dequant_is_numpy1_intent(golden_static):
    """Trap T8: the reference's loop crashes on NumPy>=2; restate out[i] = f32((nib-8)*scale64) literally."""
    g = golden_static
    p, lo, hi = g["int4.q"][:8], g["int4.min_max"][:8, 0], g["int4.min_max"][:8, 1]
    got = o.dequantize_int4_perdoc(p, 1024, lo, hi)
    gotg = o.dequantize_int4_global(p, 1024, 0.18)
    for r in range(8):
        scale = max(abs(float(lo[r])), abs(float(hi[r]))) / 7.0
        exp = np.zeros(1024, np.float32)
        expg = np.zeros(1024, np.float32)
        for i, byte in enumerate(p[r]):
            b = int(byte) if byte >= 0 else int(byte) + 256
            for j, nib in enumerate(((b >> 4) & 15, b & 15)):
                if lo[r] != hi[r]:
                    exp[2 * i + j] = (nib - 8) * scale
                expg[2 * i + j] = (nib - 8) * (0.18 / 7.0)
        assert np.array_equal(got[r], exp)
        assert np.array_equal(gotg[r], expg)


def test_to_binary_matches_reference(golden_static):
    g = golden_static
    assert np.array_equal(o.to_binary_f32(g["x"]), g["ubinary_f32"])
    assert np.array_equal(o.to_binary_f32(g["x"], ge=True), g["ubinary_f32_ge"])
    assert np.array_equal(oc.to_binary_f32(g["x"]), g["ubinary_f32"])
    assert np.array_equal(oc.to_binary_f32(g["x"], ge=True), g["ubinary_f32_ge"])
    assert np.array_equal(o.to_binary_int(g["i8"]), g["ubinary_i8"])
    assert np.array_equal(o.to_binary_int(g["i16"]), g["ubinary_i16"])
    assert np.array_equal(oc.to_binary_int(g["i8"]), g["ubinary_i8"])
    assert np.array_equal(oc.to_binary_int(g["i16"]), g["ubinary_i16"])


def test_pairwise_tree_is_numpy_mean():
    """the explicit tree == np.add.reduce on float32, for D=1024 and for ragged lengths."""
    rng = np.random.default_rng(7)
    for n in (1, 7, 8, 9, 127, 128, 129, 136, 256, 384, 768, 1000, 1024, 1536, 4096):
        for _ in range(20):
            a = (rng.normal(0, 1, n) * 10.0 ** rng.integers(-3, 3)).astype(np.float32)
            ref = np.add.reduce(a)
            assert o.pairwise_sum_f32(a) == ref, n
            assert oc.pairwise_sum_f32(a) == ref, n


# ---------------------------------------------------------------- KATs on the committed DBs ----


def test_kat1_cohere_int8_codes(golden_dbs):
    """KAT-1: packbits(int8 > mean) of the stored Cohere int8 payload == index.bin codes, bit for bit."""
    pay = golden_dbs["db_cohere_int8.payload"]
    codes = golden_dbs["db_cohere_int8.codes"]
    assert np.array_equal(o.to_binary_int(pay), codes[: pay.shape[0]])
    assert np.array_equal(oc.to_binary_int(pay), codes[: pay.shape[0]])


def test_kat1b_enhanced_codes_are_sign_bits(golden_dbs):
    """Trap T4: Cohere's ubinary ~= packbits(int8 > ~0): the stored codes agree with the sign of the stored
    int8 on all but a handful of bits (statistical pin of the synthetic stand-in we use for Cohere)."""
    pay = golden_dbs["db_cohere_enhanced.payload"]
    codes = golden_dbs["db_cohere_enhanced.codes"][: pay.shape[0]]
    bits = np.unpackbits(codes, axis=1)
    agree = (bits == (pay >= 0)).mean()
    assert agree > 0.99


def test_kat2_int4_global_ignores_limit(golden_dbs):
    """KAT-2 / trap T2: db_int4 and db_int4_global hold byte-identical packed payloads."""
    assert np.array_equal(golden_dbs["db_int4.payload"], golden_dbs["db_int4_global.payload"])
    assert str(golden_dbs["db_int4.payload_sha256"]) == str(golden_dbs["db_int4_global.payload_sha256"])


def test_kat3_rounding_modes(golden_dbs):
    """KAT-3 (statistical: the float32 inputs are not stored).  Reconstruct x from the per-doc int16 payload of
    the older VectorDBInt16 and check which rounding each codec used."""
    i16 = golden_dbs["db_int16.payload"].astype(np.float64)
    mm = golden_dbs["db_int16.min_max"]
    m = np.maximum(np.abs(mm[:, 0]), np.abs(mm[:, 1]))
    x = (i16 * (m / 32767.0)[:, None]).astype(np.float32)
    n = x.shape[0]
    q8, _, _ = o.quantize_int8_perdoc(x)
    assert (q8 == golden_dbs["db_int8.payload"][:n]).mean() > 0.99  # truncation (trap T1)
    q8r = np.round(x * (np.float32(127) / m.astype(np.float32))[:, None]).astype(np.int8)
    assert (q8r == golden_dbs["db_int8.payload"][:n]).mean() < 0.6
    assert (o.quantize_int8_global(x, 0.3) == golden_dbs["db_int8_global.payload"][:n]).mean() > 0.995
    p4, _, _ = o.quantize_int4(x)
    assert (p4 == golden_dbs["db_int4.payload"][:n]).mean() > 0.999
    d16 = np.abs(o.quantize_int16_global(x, 1.0).astype(np.int32) - golden_dbs["db_int16_global.payload"][:n])
    assert d16.max() <= 1
    bits = np.unpackbits(o.to_binary_f32(x), axis=1) != np.unpackbits(golden_dbs["db_int8.codes"][:n], axis=1)
    assert bits.sum() < 64  # 35 / 1 024 000 over all 1000 rows in the survey


def test_kat4_tie_order(golden_dbs):
    """KAT-4: the reference's FAISS output (1.log:78-127) is sorted by (distance, id) - ties by ascending id."""
    kd = golden_dbs["kat4_id_dist"]
    key = kd[:, 1] * 100000 + kd[:, 0]
    assert np.all(np.diff(key) > 0)


def test_kat4_oracle_topk_order():
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 256, (5000, 128), dtype=np.uint8)
    codes[100:400] = codes[50]  # massive tie groups
    q = rng.integers(0, 256, (4, 128), dtype=np.uint8)
    q[1] = codes[50]
    for k in (1, 10, 333, 5000, 5100):
        d1, p1 = o.hamming_topk(codes, q, k, pos_base=7, chunk=777)
        d2, p2 = oc.hamming_topk(codes, q, k, pos_base=7)
        assert np.array_equal(d1, d2) and np.array_equal(p1, p2)
        full = o.hamming_distances(codes, q)
        for i in range(4):
            order = np.lexsort((np.arange(5000), full[i]))[:k]
            m = len(order)
            assert np.array_equal(p1[i, :m], order + 7)
            assert np.array_equal(d1[i, :m], full[i][order])
            assert np.all(p1[i, m:] == -1) and np.all(d1[i, m:] == 2147483647)


def test_kat5_index_bin_layout(golden_dbs):
    """KAT-5: write_index_binary restatement reproduces the reference's header bytes and file size."""
    for db, h in golden_dbs.headers.items():
        codes = golden_dbs[f"{db}.codes"]
        b = o.write_index_binary_bytes(1024, codes, np.arange(1000))
        assert len(b) == h["index_size"] == 66 + 136 * 1000
        assert b[:58].hex() == h["index_header_hex"]
        d, c2, ids = o.read_index_binary_bytes(b)
        assert d == 1024 and np.array_equal(c2, codes) and np.array_equal(ids, np.arange(1000))
    h = golden_dbs.headers
    assert h["db_int8"]["config_json"] == o.config_json("snowflake-arctic-embed2", 1024)
    assert h["db_int8_global"]["config_json"] == o.config_json("snowflake-arctic-embed2", 1024, 0.3)
    assert h["db_cohere_enhanced"]["config_json"] == o.config_json("embed-english-v3.0", 1024)


# ---------------------------------------------------------------- rescoring + pipeline ----


def test_rescore_literal_vs_batched():
    rng = np.random.default_rng(5)
    q = rng.normal(0, 0.03, 1024).astype(np.float32) + np.float32(0.01)
    codes = rng.integers(0, 256, (300, 128), dtype=np.uint8)
    i8 = rng.integers(-128, 128, (300, 1024)).astype(np.int8)
    i8[3] = 0
    a, b = o.rescore_binary(q, codes, True), o.rescore_binary(q, codes, False)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-13)
    assert np.allclose(oc.rescore_binary(q, codes), a, rtol=1e-12, atol=1e-13)
    a, b = o.rescore_int8cos(q, i8, True), o.rescore_int8cos(q, i8, False)
    assert a[3] == -np.inf and b[3] == -np.inf
    fl = o.rescore_int8cos_absfloor(q, i8)
    ok = np.isfinite(a)
    assert np.all(np.abs(a[ok] - b[ok]) <= 1e-5 * np.abs(b[ok]) + fl[ok])
    assert np.array_equal(oc.int8_sumsq(i8), (i8.astype(np.int64) ** 2).sum(1))


def test_search3_equals_sharded_merge():
    """The multi-GPU rule: per-shard candidates + global merge == search3 on the concatenated database."""
    n, k, bo, io = 6000, 7, 10, 3
    x = o.synth_f32(11, 0, n)
    codes, i8 = o.synth_ubinary_from_f32(x), o.synth_int8_from_f32(x)
    ids = np.arange(n, dtype=np.int64) * 3 + 5
    qf = o.synth_f32(12, 0, 3)
    qb = o.synth_ubinary_from_f32(qf)
    for qi in range(3):
        full = o.search3(codes, ids, i8, qf[qi], qb[qi], k, bo, io)
        shards = []
        for s in range(4):
            a, b = s * n // 4, (s + 1) * n // 4
            # each shard returns ALL its phase-I candidates with all three scores attached
            hits = o.search3(codes[a:b], ids[a:b], i8[a:b], qf[qi], qb[qi], k * bo, 1, 1, pos_base=a)
            shards.append(hits)
        merged = o.merge_shard_results(shards, k, bo, io)
        assert [h["doc_id"] for h in merged] == [h["doc_id"] for h in full]
        assert [h["score_cosine"] for h in merged] == [h["score_cosine"] for h in full]


def test_synth_generators_agree():
    a = o.synth_f32(5, 1000, 64, 1024, row_scale=True)
    b = oc.synth_f32(5, 1000, 64, 1024, row_scale=True)
    assert np.array_equal(a, b)
    x = o.synth_f32(9, 123456789012, 32)
    c, i8 = oc.synth_codes_int8(9, 123456789012, 32)
    assert np.array_equal(c, o.synth_ubinary_from_f32(x)) and np.array_equal(i8, o.synth_int8_from_f32(x))
    assert abs(float(x.std()) - 0.0375) < 0.003 and np.abs(x).max() < 0.15


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the authoring container")
def test_live_reference_static_methods():
    """When /root/reference is mounted, run its own static methods live on fresh random rows."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg

    rng = np.random.default_rng(99)
    x = rng.normal(0, 0.05, (64, 1024)).astype(np.float32)
    M8 = mg.load_reference_module("VectorDBInt8").VectorDBInt8
    M8G = mg.load_reference_module("VectorDBInt8Global").VectorDBInt8Global
    M16G = mg.load_reference_module("VectorDBInt16Global").VectorDBInt16Global
    M4 = mg.load_reference_module("VectorDBInt4").VectorDBInt4
    q, lo, hi = o.quantize_int8_perdoc(x)
    p4, _, _ = o.quantize_int4(x)
    for r in range(64):
        qr, a, b = M8._quantize_to_int8(x[r])
        assert np.array_equal(qr, q[r]) and a == lo[r] and b == hi[r]
        assert np.array_equal(M8._to_binary(x[r]), o.to_binary_f32(x[r]))
        assert np.array_equal(M8G._quantize_to_int8(x[r], 0.07), o.quantize_int8_global(x[r], 0.07))
        assert np.array_equal(M16G._quantize_to_int16(x[r], 0.07), o.quantize_int16_global(x[r], 0.07))
        assert np.array_equal(M4._quantize_to_int4(x[r])[0], p4[r])


def test_float_index_file_layout_matches_reference_artefact():
    """KAT: the oracle's IndexIDMap(IndexFlatIP) writer reproduces the header and size of the reference's committed
    db_cohere_float/index.faiss (1000 x 1024 float32 rows, ids 0..999)."""
    import json
    h = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "float_index_header.json")))
    x = np.zeros((h["ntotal"], h["d"]), np.float32)
    b = o.write_index_float_bytes(h["d"], x, np.arange(h["ntotal"]))
    assert b[:82].hex() == h["header_hex"] and len(b) == h["size"] and h["ids_are_arange"]
    assert h["config_json"] == json.dumps({"model": "embed-english-v3.0", "embedding_dim": 1024})
