"""GPU tests of the reference-shaped classes: same calls the reference's drivers make (main.py:401-428,
maisnowflake.py:288-381), results checked against the oracle's literal restatement of each class's search."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle_c as oc  # noqa: E402
from oracle import vrq_oracle as o  # noqa: E402

DOCS = [f"AI topic number {i} is transforming field {i % 37}." for i in range(600)]
IDS = list(range(len(DOCS)))
QUERY = "Artificial intelligence is transforming industries."


def synth_rows(texts, seed=1, row_scale=False):
    from vectorragquantization_b200.embedder import text_row
    return np.stack([oc.synth_f32(seed, text_row(t), 1, 1024, row_scale)[0] for t in texts])


def close(a, b, floor):
    return abs(a - b) <= 1e-5 * abs(b) + floor


def check_search2(res, ref, floor=1e-7):
    assert len(res) == len(ref)
    rs = np.array([h["score"] for h in ref])
    for i, (r, h) in enumerate(zip(res, ref)):
        assert close(r["score"], h["score"], floor)
        sep = (i == 0 or rs[i - 1] - rs[i] > 4e-5 * abs(rs[i]) + 4 * floor) and (
            i == len(ref) - 1 or rs[i] - rs[i + 1] > 4e-5 * abs(rs[i]) + 4 * floor)
        if sep:
            assert r["doc_id"] == h["doc_id"]


@pytest.mark.parametrize("cls,kw,deq", [
    ("VectorDBInt8", {}, lambda x: o.dequantize_int8_perdoc(*o.quantize_int8_perdoc(x))),
    ("VectorDBInt8Global", {"global_limit": 0.3}, lambda x: o.dequantize_int8_global(o.quantize_int8_global(x, 0.3), 0.3)),
    ("VectorDBInt16Global", {"global_limit": 1.0}, lambda x: o.dequantize_int16_global(o.quantize_int16_global(x, 1.0), 1.0)),
    ("VectorDBInt4", {}, lambda x: o.dequantize_int4_perdoc(*((lambda p, lo, hi: (p, 1024, lo, hi))(*o.quantize_int4(x))))),
    ("VectorDBInt4Global", {"global_limit": 0.18}, lambda x: o.dequantize_int4_global(o.quantize_int4(x)[0], 1024, 0.18)),
])
def test_vectordb_classes(tmp_path, cls, kw, deq):
    import vectorragquantization_b200 as V
    C = getattr(V, cls)
    folder = os.path.join(tmp_path, "db")
    db = C(folder, **kw)
    db.add_documents(IDS, DOCS, batch_size=64)
    assert len(db) == len(DOCS) and db.index.ntotal == len(DOCS)
    cfg = json.load(open(os.path.join(folder, "config.json")))
    assert cfg["model"] == "snowflake-arctic-embed2" and cfg["embedding_dim"] == 1024
    if kw:
        assert cfg["global_limit"] == kw["global_limit"] and db.global_limit == kw["global_limit"]
    x = synth_rows(DOCS)
    ub = o.to_binary_f32(x)
    assert open(os.path.join(folder, "index.bin"), "rb").read() == o.write_index_binary_bytes(1024, ub, np.array(IDS))
    qf = synth_rows([QUERY])[0]
    qb = o.to_binary_f32(qf)
    emb = deq(x)
    for k, bo in ((10, 10), (5, 3), (100, 10)):
        ref = o.search2(ub, np.array(IDS), lambda p: emb[p], qf, qb, k, bo)
        res = db.search(QUERY, k=k, binary_oversample=bo)
        check_search2(res, ref)
        assert all(r["doc"] == DOCS[r["doc_id"]] for r in res)
        ref32 = o.search2(ub, np.array(IDS), lambda p: x[p], qf, qb, k, bo)
        check_search2(db.search(QUERY, k=k, binary_oversample=bo, compare_float32=True), ref32)
    with pytest.raises(ValueError):
        db.add_documents([1, 2], ["only one"])
    # re-adding an id replaces it (remove + add), order-preserving like faiss
    db.add_documents([3], ["a replaced document"], save=False)
    assert len(db) == len(DOCS)
    assert db.index.position_of(3) == len(DOCS) - 1
    db.remove_document(7)
    assert len(db) == len(DOCS) - 1 and "7" not in db.doc_db
    db.save()
    # reopen: config + index + quantised vectors come back; float_embeddings do not (reference: RAM only)
    db2 = C(folder, **({"global_limit": 9.9} if kw else {}))
    if kw:
        assert db2.global_limit == kw["global_limit"]  # stored limit wins (VectorDBInt8Global.py:73)
    assert len(db2) == len(DOCS) - 1
    r1, r2 = db.search(QUERY, k=10), db2.search(QUERY, k=10)
    assert [h["doc_id"] for h in r1] == [h["doc_id"] for h in r2] and [h["score"] for h in r1] == [h["score"] for h in r2]
    with pytest.raises(KeyError):
        db2.search(QUERY, k=10, compare_float32=True)


def test_static_methods_like_reference():
    import vectorragquantization_b200 as V
    x = oc.synth_f32(3, 0, 4, row_scale=True)
    q, lo, hi = V.VectorDBInt8._quantize_to_int8(x[0])
    rq, rlo, rhi = o.quantize_int8_perdoc(x[0])
    assert np.array_equal(q, rq) and lo == rlo and hi == rhi
    assert np.array_equal(V.VectorDBInt8._dequantize_int8(q, (lo, hi)), o.dequantize_int8_perdoc(rq, rlo, rhi))
    assert np.array_equal(V.VectorDBInt8._to_binary(x[1]), o.to_binary_f32(x[1]))
    assert np.array_equal(V.VectorDBInt8Global._quantize_to_int8(x[2], 0.3), o.quantize_int8_global(x[2], 0.3))
    assert np.array_equal(V.VectorDBInt16Global._quantize_to_int16(x[2], 1.0), o.quantize_int16_global(x[2], 1.0))
    p, a, b = V.VectorDBInt4._quantize_to_int4(x[3])
    rp, ra, rb = o.quantize_int4(x[3])
    assert np.array_equal(p, rp) and a == ra and b == rb and isinstance(a, float)
    assert np.array_equal(V.VectorDBInt4Global._quantize_to_int4(x[3], 0.18), rp)  # limit ignored (trap T2)
    assert np.array_equal(V.VectorDBInt4._dequantize_int4(p, 1024, (a, b)), o.dequantize_int4_perdoc(rp, 1024, ra, rb))
    assert np.array_equal(V.VectorDBInt4Global._dequantize_int4(p, 1024, 0.18), o.dequantize_int4_global(rp, 1024, 0.18))


def test_vectordb_int16_hamming_only(tmp_path):
    import vectorragquantization_b200 as V
    db = V.VectorDBInt16(os.path.join(tmp_path, "db16"))
    db.add_documents(IDS, DOCS)
    x = synth_rows(DOCS)
    i16 = o.quantize_int16_global(x, 1.0)
    codes = o.to_binary_int(i16)
    q16 = o.quantize_int16_global(synth_rows([QUERY]), 1.0)
    d, p = oc.hamming_topk(codes, o.to_binary_int(q16), 100)
    res = db.search(QUERY, k=10)
    assert [r["doc_id"] for r in res] == p[0][:10].tolist() and [r["score"] for r in res] == d[0][:10].tolist()
    assert db.search(QUERY, k=1000, binary_oversample=10)[-1]["doc_id"] == oc.hamming_topk(codes, o.to_binary_int(q16), 600)[1][0][-1]


def test_cohere_enhanced_class(tmp_path):
    import vectorragquantization_b200 as V
    from vectorragquantization_b200.embedder import text_row
    folder = os.path.join(tmp_path, "enh")
    db = V.CohereEnhancedVectorDB(folder)
    db.add_documents(IDS, DOCS, batch_size=64)
    assert len(db) == len(DOCS)
    assert open(os.path.join(folder, "config.json")).read() == o.config_json("embed-english-v3.0", 1024)
    rows = [text_row(t) for t in DOCS]
    pairs = [oc.synth_codes_int8(1, r, 1) for r in rows]
    codes = np.stack([p[0][0] for p in pairs])
    i8 = np.stack([p[1][0] for p in pairs])
    assert open(os.path.join(folder, "index.bin"), "rb").read() == o.write_index_binary_bytes(1024, codes, np.array(IDS))
    qr = text_row(QUERY)
    qf = oc.synth_f32(1, qr, 1)[0]
    qb = oc.synth_codes_int8(1, qr, 1)[0][0]
    for k, bo, io in ((10, 10, 3), (50, 10, 3), (3, 2, 2)):
        ref = o.search3(codes, np.array(IDS), i8, qf, qb, k, bo, io)
        res = db.search(QUERY, k=k, binary_oversample=bo, int8_oversample=io)
        assert [r["doc_id"] for r in res] == [h["doc_id"] for h in ref]
        assert [r["score_hamming"] for r in res] == [h["score_hamming"] for h in ref]
        for r, h in zip(res, ref):
            assert isinstance(r["doc_id"], int) and r["doc"] == DOCS[r["doc_id"]]
            assert close(r["score_binary"], h["score_binary"], 1e-12)
            assert close(r["score_cosine"], h["score_cosine"], float(o.rescore_int8cos_absfloor(qf, i8[[h["doc_id"]]])[0]))
    db.remove_document(ref[0]["doc_id"])
    assert db.search(QUERY, k=3, binary_oversample=2, int8_oversample=2)[0]["doc_id"] != ref[0]["doc_id"]
    db2 = V.CohereEnhancedVectorDB(folder)
    assert len(db2) == len(DOCS) - 1
    a, b = db.search(QUERY, k=10), db2.search(QUERY, k=10)
    assert a == b
    assert V.CohereEnhancedVectorDB(os.path.join(tmp_path, "empty")).search(QUERY) == []
    with pytest.raises(Exception):
        os.makedirs(os.path.join(tmp_path, "junk"))
        open(os.path.join(tmp_path, "junk", "x"), "w").write("x")
        V.CohereEnhancedVectorDB(os.path.join(tmp_path, "junk"))


def test_sharded_world1_cuda_engine():
    """ShardedSearch3 with the CUDA engine and no process group == BinaryIndex.search3."""
    import torch
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import _lib as L
    from vectorragquantization_b200.sharded import CudaEngine, ShardedSearch3
    n, nq = 50000, 16
    ix = V.BinaryIndex(1024, payload_kind=L.PAYLOAD_INT8_RAW)
    ix.add_synthetic(5, 0, n, 100)
    x = oc.synth_f32(5, 0, nq) + oc.synth_f32(6, 0, nq) * np.float32(0.5)
    qf, qb = x.astype(np.float32), o.synth_ubinary_from_f32(x.astype(np.float32))
    want = ix.search3(qf, qb, 10, 10, 3)
    s = ShardedSearch3(CudaEngine(ix), pos_base=0)
    out = s.search(qf, qb, 10, 10, 3)
    torch.cuda.synchronize()
    got = [out[k].cpu().numpy() for k in ("labels", "hamming", "score_binary", "score_cosine", "count")]
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    codes, i8 = oc.synth_codes_int8(5, 0, n)
    ref = o.search3(codes, np.arange(n) + 100, i8, qf[0], qb[0], 10, 10, 3)
    assert [h["doc_id"] for h in ref] == got[0][0].tolist()


def test_cohere_binary_class(tmp_path):
    """CohereVectorDBBinary: '>=' mean threshold, rescoring = float32 dot(query, +-1 unpacked code)."""
    import vectorragquantization_b200 as V
    db = V.CohereVectorDBBinary(os.path.join(tmp_path, "bin"))
    db.add_documents(IDS, DOCS)
    x = synth_rows(DOCS)
    codes = o.to_binary_f32(x, ge=True)
    assert open(os.path.join(tmp_path, "bin", "index.bin"), "rb").read() == o.write_index_binary_bytes(1024, codes, np.array(IDS))
    qf = synth_rows([QUERY])[0]
    qb = o.to_binary_f32(qf, ge=True)
    pm1 = np.where(np.unpackbits(codes, axis=1) == 0, -1, 1).astype(np.float32)
    for k, bo in ((10, 10), (7, 3)):
        check_search2(db.search(QUERY, k=k, binary_oversample=bo), o.search2(codes, np.array(IDS), lambda p: pm1[p], qf, qb, k, bo), 1e-5)
        check_search2(db.search(QUERY, k=k, binary_oversample=bo, compare_float32=True),
                      o.search2(codes, np.array(IDS), lambda p: x[p], qf, qb, k, bo))
    sb = V.CohereVectorDBBinary._to_signed_binary(x[5])
    assert np.array_equal(sb, np.where(x[5] >= x[5].mean(), 1, -1)) and sb.dtype == np.int8
    assert np.array_equal(V.CohereVectorDBBinary._pack_signed_binary(sb), codes[5])
    assert np.array_equal(V.CohereVectorDBBinary._unpack_signed_binary(codes[5], 1024), pm1[5])
    db2 = V.CohereVectorDBBinary(os.path.join(tmp_path, "bin"))
    assert [h["doc_id"] for h in db2.search(QUERY, k=10)] == [h["doc_id"] for h in db.search(QUERY, k=10)]


def test_cohere_int8_class(tmp_path):
    """CohereVectorDBInt8: packbits(int8 > mean(int8)) codes, Hamming-only search."""
    import vectorragquantization_b200 as V
    from vectorragquantization_b200.embedder import text_row
    db = V.CohereVectorDBInt8(os.path.join(tmp_path, "ci8"))
    db.add_documents(IDS, DOCS)
    i8 = np.stack([oc.synth_codes_int8(1, text_row(t), 1, want_codes=False)[1][0] for t in DOCS])
    codes = o.to_binary_int(i8)
    assert open(os.path.join(tmp_path, "ci8", "index.bin"), "rb").read() == o.write_index_binary_bytes(1024, codes, np.array(IDS))
    q8 = oc.synth_codes_int8(1, text_row(QUERY), 1, want_codes=False)[1]
    d, p = oc.hamming_topk(codes, o.to_binary_int(q8), 100)
    res = db.search(QUERY, k=10)
    assert [r["doc_id"] for r in res] == p[0][:10].tolist() and [r["score"] for r in res] == d[0][:10].tolist()
    with pytest.raises(NotImplementedError):
        db.search_rerank_cohere(QUERY)


def _golden_walker():
    """tests/golden/make_golden.py's own decoder of the reference's database files (test infrastructure, independent of the
    product's importer)."""
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m, GOLDEN


def test_open_reference_written_cohere_enhanced(tmp_path):
    """A database folder WRITTEN BY THE REFERENCE (config.json + faiss index.bin + rocksdict docs/, committed in the
    reference repo as db_cohere_enhanced) opens as-is (CohereEnhancedVectorDB.py:116-128); the 3-phase search over it equals
    the oracle run on the arrays decoded independently from the same files.  Queries: Cohere-like (float, ubinary) pairs
    derived from stored documents, so there are true neighbours."""
    import shutil
    import vectorragquantization_b200 as V
    g, GOLDEN = _golden_walker()
    folder = os.path.join(tmp_path, "db_cohere_enhanced")
    shutil.copytree(os.path.join(GOLDEN, "db_cohere_enhanced"), folder)
    d, codes, ids, _ = g.read_index_bin(os.path.join(folder, "index.bin"))
    docs = g.read_docs(os.path.join(folder, "docs", "000009.sst"))
    i8 = np.stack([np.asarray(docs[int(i)]["int8"]) for i in ids])
    rng = np.random.default_rng(3)
    qsrc = [5, 123, 777, 999]
    qf = np.stack([(i8[j].astype(np.float32) + rng.normal(0, 12, 1024).astype(np.float32)) / np.float32(1259.0) for j in qsrc])
    qb = np.packbits(qf > 0, axis=1)
    table = {f"q{j}": i for i, j in enumerate(qsrc)}

    def embedder(texts, input_type, embedding_types):
        sel = [table[t] for t in texts]
        return {"float": qf[sel], "ubinary": qb[sel]}

    db = V.CohereEnhancedVectorDB(folder, embedder=embedder)
    assert len(db) == 1000 and db.index.payload_kind == V._lib.PAYLOAD_INT8_RAW
    assert np.array_equal(db.index.read_rows(V._lib.ROWS_PAYLOAD, 0, 1000), i8)
    for k, bo, io in ((10, 10, 3), (50, 10, 3)):
        for qi, j in enumerate(qsrc):
            ref = o.search3(codes, ids, i8, qf[qi], qb[qi], k, bo, io)
            res = db.search(f"q{j}", k=k, binary_oversample=bo, int8_oversample=io)
            assert [r["doc_id"] for r in res] == [h["doc_id"] for h in ref]
            assert [r["score_hamming"] for r in res] == [h["score_hamming"] for h in ref]
            for r, h in zip(res, ref):
                assert r["doc"] == docs[r["doc_id"]]["doc"]
                assert close(r["score_binary"], h["score_binary"], 1e-12)
                assert close(r["score_cosine"], h["score_cosine"], float(o.rescore_int8cos_absfloor(qf[qi], i8[[h["doc_id"]]])[0]))
    # save() next to the reference's files (index.bin byte-identical, streamed sidecar, docs.log overlay), reopen, same answers
    before = open(os.path.join(folder, "index.bin"), "rb").read()
    db.remove_document(int(ids[17]), save=False)
    db.add_embeddings([5000], i8[17:18], codes[17:18], docs=["re-added"], save=True)
    db2 = V.CohereEnhancedVectorDB(folder, embedder=embedder)
    assert len(db2) == 1000 and os.path.exists(os.path.join(folder, "payload.vrqp"))
    assert db2.search("q5", k=10) == db.search("q5", k=10)
    assert db2.doc_db.get("5000")["doc"] == "re-added" and "17" not in db2.doc_db
    assert len(before) == 136066 and g.read_index_bin(os.path.join(folder, "index.bin"))[2][-1] == 5000


def test_open_reference_written_int8(tmp_path):
    """The reference's committed db_int8 folder (VectorDBInt8: per-document int8 + min_max in the docs store) opens and its
    2-phase search equals the oracle on the independently decoded arrays."""
    import shutil
    import vectorragquantization_b200 as V
    g, GOLDEN = _golden_walker()
    folder = os.path.join(tmp_path, "db_int8")
    shutil.copytree(os.path.join(GOLDEN, "db_int8"), folder)
    d, codes, ids, _ = g.read_index_bin(os.path.join(folder, "index.bin"))
    docs = g.read_docs(os.path.join(folder, "docs", "000009.sst"))
    q8 = np.stack([np.asarray(docs[int(i)]["emb_int8"]) for i in ids])
    lo = np.array([docs[int(i)]["min_max"][0] for i in ids], np.float32)
    hi = np.array([docs[int(i)]["min_max"][1] for i in ids], np.float32)
    emb = o.dequantize_int8_perdoc(q8, lo, hi)
    db = V.VectorDBInt8(folder, embedder=lambda texts: np.stack([emb[int(t)] for t in texts]))
    assert len(db) == 1000
    for j in (3, 500, 998):
        qf = emb[j]
        qb = o.to_binary_f32(qf)
        for k, bo in ((10, 10), (100, 10)):
            check_search2(db.search(str(j), k=k, binary_oversample=bo), o.search2(codes, ids, lambda p: emb[p], qf, qb, k, bo))
    assert db.search("3", k=1)[0]["doc"] == docs[db.search("3", k=1)[0]["doc_id"]]["doc"]
    db.save()
    db2 = V.VectorDBInt8(folder, embedder=lambda texts: np.stack([emb[int(t)] for t in texts]))
    assert db2.search("500", k=10) == db.search("500", k=10)


def test_remove_and_readd_without_float_rows(tmp_path):
    """add_embeddings(keep_float=False) with a duplicate id, and a re-add after a reopen, must neither raise nor
    leave the document half removed (the reference raises KeyError from `del float_embeddings[id]` after a reopen)."""
    import vectorragquantization_b200 as V
    folder = os.path.join(tmp_path, "db")
    db = V.VectorDBInt8Global(folder, global_limit=0.3)
    x = synth_rows(DOCS[:50])
    db.add_embeddings(IDS[:50], x, DOCS[:50], keep_float=False)
    db.add_embeddings([7], x[8:9], ["seven again"], keep_float=False)  # duplicate id: replaced, not an exception
    assert len(db) == 50 and db.doc_db.get("7")["doc"] == "seven again" and db.index.position_of(7) == 49
    db.save()
    db2 = V.VectorDBInt8Global(folder)
    db2.add_documents([3], ["a replaced document"])  # re-add after a reopen (float_embeddings is empty)
    assert len(db2) == 50 and db2.index.position_of(3) == 49
    db2.remove_document(4)
    assert len(db2) == 49 and "4" not in db2.doc_db


def test_cohere_float_class(tmp_path):
    """CohereVectorDBFloat: float32 rows, brute-force inner-product top-k (faiss IndexIDMap(IndexFlatIP)), index.faiss bytes."""
    import vectorragquantization_b200 as V
    from vectorragquantization_b200.embedder import text_row
    folder = os.path.join(tmp_path, "flt")
    db = V.CohereVectorDBFloat(folder)
    db.add_documents(IDS, DOCS, batch_size=64)
    assert len(db) == len(DOCS)
    assert open(os.path.join(folder, "config.json")).read() == json.dumps({"model": "embed-english-v3.0", "embedding_dim": 1024})
    x = np.stack([oc.synth_f32(1, text_row(t), 1)[0] for t in DOCS])
    assert open(os.path.join(folder, "index.faiss"), "rb").read() == o.write_index_float_bytes(1024, x, np.array(IDS))
    qf = oc.synth_f32(1, text_row(QUERY), 1)[0]
    for k in (10, 100, 700):
        rs, rl = o.search_ip(x, np.array(IDS), qf, k)
        res = db.search(QUERY, k=k)
        got_s = np.array([r["score"] for r in res])
        assert len(res) == min(k, len(DOCS))
        assert np.all(np.abs(got_s - rs[0][:len(res)]) <= 1e-5 * np.abs(rs[0][:len(res)]) + 1e-7)
        ref_s = rs[0][:len(res)]
        gap = np.abs(np.diff(ref_s))
        apart = np.r_[True, gap > 1e-6] & np.r_[gap > 1e-6, True]  # a rank is only defined where both neighbouring scores differ
        same = np.array([r["doc_id"] for r in res]) == rl[0][:len(res)]
        assert np.all(same[apart]) and np.mean(same) > 0.9
        assert all(r["doc"] == DOCS[r["doc_id"]] for r in res)
    # batch of queries incl. more than one pass of 8, and k > ntotal padding
    Q = np.stack([oc.synth_f32(3, i, 1)[0] for i in range(19)])
    s, l = db.search_batch(Q, 5)
    rs, rl = o.search_ip(x, np.array(IDS), Q, 5)
    assert np.allclose(s, rs, rtol=1e-5, atol=1e-7) and np.mean(l == rl) > 0.98
    db.remove_document(rl[0][0])
    assert db.search_batch(Q[:1], 1)[1][0, 0] != rl[0][0]
    db2 = V.CohereVectorDBFloat(folder)
    assert len(db2) == len(DOCS) - 1
    assert [h["doc_id"] for h in db2.search(QUERY, k=10)] == [h["doc_id"] for h in db.search(QUERY, k=10)]
    small = V.CohereVectorDBFloat(os.path.join(tmp_path, "small"))
    small.add_documents(IDS[:3], DOCS[:3])
    assert len(small.search(QUERY, k=10)) == 3
    assert V.CohereVectorDBFloat(os.path.join(tmp_path, "empty")).search(QUERY) == []
