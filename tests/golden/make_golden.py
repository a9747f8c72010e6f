#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the reference's committed 1000-document databases.

Run HERE (authoring container, /root/reference mounted read-only); the GPU box never sees
/root/reference, it only sees the small .npz files this script writes.  Nothing in the
product imports this file.

What is decoded:
  * db_*/index.bin          faiss IndexBinaryIDMap2(IndexBinaryFlat) -> codes u8[1000,128], ids i64[1000]
  * db_*/docs/000009.sst    RocksDB block-based table written by rocksdict -> per-document pickles
The SST walker is a from-scratch reader of the public RocksDB table format (footer -> metaindex ->
index block -> data blocks, raw or Snappy).  Pickles are loaded through a whitelist unpickler: the
reference is untrusted content.

Also runs the reference's own static NumPy methods (faiss / rocksdict stubbed in sys.modules) on
seeded inputs and stores input/output pairs, so the oracle can be pinned against the reference's
actual code without /root/reference being present later.
"""
import hashlib
import io
import json
import os
import pickle
import re
import struct
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
NROWS = 256  # payload rows kept per database (codes keep all 1000 rows)

_OK = {("numpy.core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
       ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "_reconstruct"),
       ("numpy._core.multiarray", "scalar")}


class _RU(pickle.Unpickler):
    def find_class(self, m, n):
        if (m, n) not in _OK:
            raise pickle.UnpicklingError(f"blocked {m}.{n}")
        return super().find_class(m, n)


def _varint(b, o):
    r = s = 0
    while True:
        c = b[o]
        o += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, o


def _snappy(b):
    n, o = _varint(b, 0)
    out = bytearray()
    while o < len(b):
        t = b[o]
        o += 1
        k = t & 3
        if k == 0:
            ln = t >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(b[o:o + nb], "little")
                o += nb
            ln += 1
            out += b[o:o + ln]
            o += ln
        else:
            if k == 1:
                ln = ((t >> 2) & 7) + 4
                off = ((t >> 5) << 8) | b[o]
                o += 1
            elif k == 2:
                ln = (t >> 2) + 1
                off = int.from_bytes(b[o:o + 2], "little")
                o += 2
            else:
                ln = (t >> 2) + 1
                off = int.from_bytes(b[o:o + 4], "little")
                o += 4
            for _ in range(ln):
                out.append(out[-off])
    assert len(out) == n
    return bytes(out)


def _block(b, off, size):
    raw = b[off:off + size]
    return _snappy(raw) if b[off + size] == 1 else raw


def _entries(raw, index=False):
    nr = struct.unpack_from("<I", raw, len(raw) - 4)[0] & 0x7FFFFFFF
    end, o, key = len(raw) - 4 - 4 * nr, 0, b""
    while o < end:
        sh, o = _varint(raw, o)
        ns, o = _varint(raw, o)
        if index:
            key = key[:sh] + raw[o:o + ns]
            o += ns
            ho, o = _varint(raw, o)
            hs, o = _varint(raw, o)
            yield key, (ho, hs)
        else:
            vl, o = _varint(raw, o)
            key = key[:sh] + raw[o:o + ns]
            o += ns
            yield key, raw[o:o + vl]
            o += vl


def read_docs(sst):
    b = open(sst, "rb").read()
    L = len(b)
    assert b[-8:] == bytes.fromhex("f7cff485b741e288")
    msz = struct.unpack_from("<I", b, L - 53 + 13)[0]
    meta = dict(_entries(_block(b, L - 53 - 5 - msz, msz)))
    v = meta[b"rocksdb.index"]
    ho, o = _varint(v, 0)
    hs, _ = _varint(v, o)
    out = {}
    for _, (bo, bs) in _entries(_block(b, ho, hs), index=True):
        for k, val in _entries(_block(b, bo, bs)):
            out[int(k[1:-8])] = _RU(io.BytesIO(val[1:])).load()
    return out


def read_index_bin(path):
    b = open(path, "rb").read()
    assert b[0:4] == b"IBM2" and b[25:29] == b"IBxF"
    d, cs, nt = struct.unpack_from("<iiq", b, 4)
    nbytes = struct.unpack_from("<Q", b, 50)[0]
    assert nbytes == nt * cs
    codes = np.frombuffer(b, np.uint8, nbytes, 58).reshape(nt, cs).copy()
    nid = struct.unpack_from("<Q", b, 58 + nbytes)[0]
    ids = np.frombuffer(b, np.int64, nid, 66 + nbytes).copy()
    return d, codes, ids, b


def load_reference_module(name):
    """Import /root/reference/<name>.py with faiss / rocksdict stubbed (not installed here)."""
    for stub in ("faiss", "rocksdict"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.IndexBinaryFlat = object
            m.Rdict = object
            sys.modules[stub] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return __import__(name)


def copy_fixture_dbs():
    """Two of the reference's committed 1000-document database folders, byte for byte (DATA, not source: config.json,
    index.bin and the RocksDB files of docs/), so that the GPU box - which never sees /root/reference - can check that
    a folder WRITTEN BY THE REFERENCE opens in this package (tests/test_gpu_classes.py::test_open_reference_written_*)."""
    import shutil
    for db in ("db_cohere_enhanced", "db_int8"):
        dst = os.path.join(OUT, db)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        os.makedirs(os.path.join(dst, "docs"))
        for f in ("config.json", "index.bin"):
            shutil.copyfile(f"{REF}/{db}/{f}", os.path.join(dst, f))
        for f in os.listdir(f"{REF}/{db}/docs"):
            if f.endswith(".sst") or f in ("CURRENT", "rocksdict-config.json") or f.startswith("MANIFEST"):
                shutil.copyfile(f"{REF}/{db}/docs/{f}", os.path.join(dst, "docs", f))
        os.chmod(dst, 0o755)
        print("copied", db, sum(os.path.getsize(os.path.join(r, x)) for r, _, fs in os.walk(dst) for x in fs), "bytes")


def float_index_header():
    """Header bytes, size and the first / last stored values of the reference's committed float index (db_cohere_float):
    pins the oracle's restatement of faiss's IndexIDMap(IndexFlatIP) file layout."""
    b = open(f"{REF}/db_cohere_float/index.faiss", "rb").read()
    d, nt = struct.unpack_from("<iq", b, 4)
    rows = np.frombuffer(b, np.float32, nt * d, 82).reshape(nt, d)
    ids = np.frombuffer(b, np.int64, nt, 82 + 4 * nt * d + 8)
    out = {"header_hex": b[:82].hex(), "size": len(b), "d": d, "ntotal": nt, "config_json": open(f"{REF}/db_cohere_float/config.json").read(),
           "row0_first8": rows[0, :8].tolist(), "rowlast_last8": rows[-1, -8:].tolist(), "ids_first": int(ids[0]), "ids_last": int(ids[-1]),
           "ids_are_arange": bool(np.array_equal(ids, np.arange(nt))), "sha256": hashlib.sha256(b).hexdigest()}
    json.dump(out, open(os.path.join(OUT, "float_index_header.json"), "w"), indent=1)
    print("float index header", out["size"], out["d"], out["ntotal"])


def main():
    copy_fixture_dbs()
    float_index_header()
    rows = {}
    dbs = ["db_int8", "db_int8_global", "db_int4", "db_int4_global", "db_int16", "db_int16_global",
           "db_cohere_int8", "db_cohere_enhanced"]
    payload_key = {"db_int8": "emb_int8", "db_int8_global": "emb_int8", "db_int4": "emb_int4",
                   "db_int4_global": "emb_int4", "db_int16": "emb_int16", "db_int16_global": "emb_int16",
                   "db_cohere_int8": "int8", "db_cohere_enhanced": "int8"}
    out = {}
    headers = {}
    for db in dbs:
        d, codes, ids, raw = read_index_bin(f"{REF}/{db}/index.bin")
        docs = read_docs(f"{REF}/{db}/docs/000009.sst")
        assert d == 1024 and codes.shape == (1000, 128) and np.array_equal(ids, np.arange(1000))
        key = payload_key[db]
        pay = np.stack([np.asarray(docs[i][key]) for i in range(1000)])
        out[f"{db}.codes"] = codes
        out[f"{db}.payload"] = pay[:NROWS]
        out[f"{db}.payload_sha256"] = np.array(hashlib.sha256(pay.tobytes()).hexdigest())  # all 1000 rows
        if "min_max" in docs[0]:
            out[f"{db}.min_max"] = np.array([[float(docs[i]["min_max"][0]), float(docs[i]["min_max"][1])]
                                             for i in range(NROWS)], dtype=np.float64)
            out[f"{db}.min_max_dtype"] = np.array(type(docs[0]["min_max"][0]).__name__)
        headers[db] = {"config_json": open(f"{REF}/{db}/config.json").read(),
                       "index_header_hex": raw[:58].hex(),
                       "index_size": len(raw)}
        print(db, key, pay.dtype, pay.shape, "min_max" in docs[0])
    out["headers_json"] = np.array(json.dumps(headers))
    # de-duplicate byte-identical arrays (five Snowflake DBs share codes; int4 == int4_global, trap T2)
    seen, alias = {}, {}
    for k in sorted(out):
        if isinstance(out[k], np.ndarray) and out[k].ndim == 2:
            h = hashlib.sha256(out[k].tobytes()).hexdigest()
            if h in seen:
                alias[k] = seen[h]
            else:
                seen[h] = k
    for k in alias:
        del out[k]
    out["alias_json"] = np.array(json.dumps(alias))

    # KAT-4: top-50 (doc_id, hamming) printed by the reference's FAISS run (1.log:78-127)
    kat4 = []
    log = open(f"{REF}/1.log", errors="replace").read().split("\n")
    start = next(i for i, l in enumerate(log) if "Raw Search Results (Cohere Int8)" in l)
    for line in log[start + 1:]:
        m = re.search(r"DocID=(\d+), Raw Score=(\d+),", line)
        if m:
            kat4.append((int(m.group(1)), int(m.group(2))))
        elif kat4:
            break
    assert len(kat4) == 50
    out["kat4_id_dist"] = np.array(kat4, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "reference_dbs.npz"), **out)

    # Reference static methods run on seeded inputs -> input/output pairs.
    rng = np.random.default_rng(20261018)
    n = 96
    x = np.empty((n, 1024), np.float32)
    sig = [1e-4, 0.036, 0.2, 1.0, 5.0, 0.036]
    for i in range(n):
        x[i] = rng.normal(0, sig[i % len(sig)], 1024).astype(np.float32)
    # adversarial rows: constants, exact half steps, +-limit, elements equal to the mean
    x[0] = 0.0
    x[1] = 0.25
    x[2] = np.where(np.arange(1024) % 2 == 0, 0.3, -0.3)
    x[3] = (np.arange(1024) - 512 + 0.5) * np.float32(0.3 / 127.0)
    x[4] = (np.arange(1024) - 512 + 0.5) * np.float32(1.0 / 32767.0)
    x[5] = np.nextafter(np.float32(0.3), np.float32(1.0))
    x[5, ::3] = -x[5, ::3]
    x[6] = np.where(np.arange(1024) < 512, 1.0, -1.0)  # mean exactly 0, no element equals it
    x[7] = np.where(np.arange(1024) % 4 == 0, 0.0, np.where(np.arange(1024) % 4 == 1, 0.5, -0.25))
    x[8, :] = 0.0
    x[8, 17] = 1e-30
    ref = {"x": x}
    M8 = load_reference_module("VectorDBInt8").VectorDBInt8
    M8G = load_reference_module("VectorDBInt8Global").VectorDBInt8Global
    M16G = load_reference_module("VectorDBInt16Global").VectorDBInt16Global
    M4 = load_reference_module("VectorDBInt4").VectorDBInt4
    M4G = load_reference_module("VectorDBInt4Global").VectorDBInt4Global
    M16 = load_reference_module("VectorDBInt16").VectorDBInt16
    q8, mm8 = [], []
    for r in x:
        q, lo, hi = M8._quantize_to_int8(r)
        q8.append(q)
        mm8.append((lo, hi))
    ref["int8_perdoc.q"] = np.stack(q8)
    ref["int8_perdoc.min_max"] = np.array(mm8, np.float32)
    ref["int8_perdoc.deq"] = np.stack([M8._dequantize_int8(q, (np.float32(a), np.float32(b)))
                                        for q, (a, b) in zip(q8, mm8)])
    ref["ubinary_f32"] = np.stack([M8._to_binary(r) for r in x])
    for lim in (0.18, 0.3, 1.0):
        q = np.stack([M8G._quantize_to_int8(r, lim) for r in x])
        ref[f"int8_global.q.{lim}"] = q
        if lim == 0.3:
            ref[f"int8_global.deq.{lim}"] = np.stack([M8G._dequantize_int8(r, lim) for r in q])
        q = np.stack([M16G._quantize_to_int16(r, lim) for r in x])
        ref[f"int16_global.q.{lim}"] = q
        if lim == 1.0:
            ref[f"int16_global.deq.{lim}"] = np.stack([M16G._dequantize_int16(r, lim) for r in q])
    q4, mm4 = [], []
    for r in x:
        q, lo, hi = M4._quantize_to_int4(r)
        q4.append(q)
        mm4.append((lo, hi))
        assert np.array_equal(q, M4G._quantize_to_int4(r, 0.18))  # trap T2: limit is ignored
    ref["int4.q"] = np.stack(q4)
    ref["int4.min_max"] = np.array(mm4, np.float64)
    i16 = rng.integers(-32768, 32768, size=(64, 1024)).astype(np.int16)
    i16[0] = 7
    i16[1, :] = np.where(np.arange(1024) % 2 == 0, 3, 4)
    ref["i16"] = i16
    ref["ubinary_i16"] = np.stack([M16._to_binary(r) for r in i16])
    i8 = rng.integers(-128, 128, size=(64, 1024)).astype(np.int8)
    i8[0] = -5
    ref["i8"] = i8
    MC = load_reference_module("CohereVectorDBInt8").CohereVectorDBInt8
    ref["ubinary_i8"] = np.stack([MC._to_binary(r) for r in i8])
    # CohereVectorDBBinary (needs the absent azure SDK at import time -> stub it too): '>=' threshold (trap T10)
    for stub in ("azure", "azure.ai", "azure.ai.inference", "azure.core", "azure.core.credentials"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.EmbeddingsClient = object
            m.AzureKeyCredential = object
            sys.modules[stub] = m
    MB = load_reference_module("CohereVectorDBBinary").CohereVectorDBBinary
    ref["ubinary_f32_ge"] = np.stack([MB._pack_signed_binary(MB._to_signed_binary(r)) for r in x])
    np.savez_compressed(os.path.join(OUT, "reference_static_methods.npz"), **ref)
    for f in ("reference_dbs.npz", "reference_static_methods.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
