"""CPU tests: the read-only importer of the reference's docs/ store (RocksDB table files written by rocksdict) and the
DocStore log, against the payloads tests/golden/make_golden.py decoded from the same committed databases with its own
walker, and against hashes of all 1000 rows."""
import hashlib
import os
import shutil

import numpy as np

from conftest import GOLDEN
from vectorragquantization_b200 import rocks_import as R
from vectorragquantization_b200.docstore import DocStore


def test_import_cohere_enhanced_docs(golden_dbs):
    d = R.read_rocksdict_folder(os.path.join(GOLDEN, "db_cohere_enhanced", "docs"))
    assert len(d) == 1000 and set(d) == {str(i) for i in range(1000)}
    rows = np.stack([np.asarray(d[str(i)]["int8"]) for i in range(1000)])
    assert rows.dtype == np.int8 and rows.shape == (1000, 1024)
    assert np.array_equal(rows[:256], golden_dbs["db_cohere_enhanced.payload"])
    assert hashlib.sha256(rows.tobytes()).hexdigest() == str(golden_dbs["db_cohere_enhanced.payload_sha256"])
    assert all(isinstance(d[str(i)]["doc"], str) and d[str(i)]["doc"] for i in (0, 499, 999))


def test_import_int8_perdoc_docs(golden_dbs):
    d = R.read_rocksdict_folder(os.path.join(GOLDEN, "db_int8", "docs"))
    assert len(d) == 1000
    rows = np.stack([np.asarray(d[str(i)]["emb_int8"]) for i in range(1000)])
    assert hashlib.sha256(rows.tobytes()).hexdigest() == str(golden_dbs["db_int8.payload_sha256"])
    mm = np.array([[float(d[str(i)]["min_max"][0]), float(d[str(i)]["min_max"][1])] for i in range(256)])
    assert np.array_equal(mm, golden_dbs["db_int8.min_max"])


def test_unflushed_wal_is_refused(tmp_path):
    dst = tmp_path / "docs"
    shutil.copytree(os.path.join(GOLDEN, "db_int8", "docs"), dst)
    (dst / "000012.log").write_bytes(b"x" * 10)
    try:
        R.read_rocksdict_folder(str(dst))
        assert False, "expected RocksImportError"
    except R.RocksImportError as e:
        assert "write-ahead log" in str(e)


def test_docstore_overlay_on_imported_store_and_torn_tail(tmp_path):
    dst = tmp_path / "docs"
    shutil.copytree(os.path.join(GOLDEN, "db_cohere_enhanced", "docs"), dst)
    s = DocStore(str(dst))
    assert len(s) == 1000 and "7" in s and set(s.get("7")) == {"doc"}  # the vectors are not kept in the doc store
    assert "int8" in s.imported_raw["7"]
    del s["7"]
    s["1000"] = {"doc": "new"}
    s.set_many([("1001", {"doc": "a"}), ("1002", {"doc": "b"})])
    s.close()
    with open(dst / "docs.log", "ab") as f:
        f.write(b"\x80\x04\x95\xff\xff")  # a torn record
    s2 = DocStore(str(dst))
    assert len(s2) == 1002 and "7" not in s2 and s2["1000"]["doc"] == "new" and s2.get("1002")["doc"] == "b"
    s2["1003"] = {"doc": "c"}  # appending after the cut works
    s2.close()
    assert DocStore(str(dst)).get("1003")["doc"] == "c"


def test_docstore_compacts_a_mostly_dead_log(tmp_path):
    s = DocStore(str(tmp_path / "docs"))
    for r in range(5):
        for i in range(600):
            s[str(i)] = {"doc": f"{r}-{i}"}
    s.close()
    size = os.path.getsize(tmp_path / "docs" / "docs.log")
    s2 = DocStore(str(tmp_path / "docs"))
    assert len(s2) == 600 and s2["599"]["doc"] == "4-599"
    assert os.path.getsize(tmp_path / "docs" / "docs.log") < size / 3
