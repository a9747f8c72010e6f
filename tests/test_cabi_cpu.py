"""CPU-only checks of the drop-in boundary: libvrq.so loads, exports every symbol include/vrq.h declares, the ctypes
table matches the header, and the product fails loudly (never falls back) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vrq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(vrq_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from vectorragquantization_b200 import _lib as L
    lib = ctypes.CDLL(L.LIB_PATH)
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vrq.h but not exported by libvrq.so"


def test_ctypes_table_matches_header():
    from vectorragquantization_b200 import _lib as L
    assert set(L.SIGNATURES) == header_functions()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vrq.h")).read(), flags=re.S)
    for name, (_, args) in L.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less box")
    import vectorragquantization_b200 as V
    from vectorragquantization_b200 import kernels as K
    assert V._lib.load().vrq_version() == 100
    with pytest.raises(V.VrqError):
        K.quantize_int8_global(np.zeros((4, 1024), np.float32), 0.3)
    with pytest.raises(V.VrqError):
        V.BinaryIndex(1024)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vectorragquantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "vrqo_" not in txt, f


def test_docstore_roundtrip(tmp_path):
    from vectorragquantization_b200.docstore import DocStore
    p = os.path.join(tmp_path, "docs")
    d = DocStore(p)
    d["1"] = {"doc": "a"}
    d["2"] = {"doc": "b"}
    del d["1"]
    assert "1" not in d and d.get("2")["doc"] == "b" and d.get("9") is None and len(d) == 1
    d.close()
    d2 = DocStore(p)
    assert "1" not in d2 and d2["2"]["doc"] == "b"


def test_shard_ranges_cover_exactly():
    from vectorragquantization_b200.sharded import shard_range
    for n in (0, 1, 7, 1000, 10 ** 9 + 7):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_text_row_is_stable():
    from vectorragquantization_b200.embedder import text_row
    assert text_row("Artificial intelligence is transforming industries.") == text_row(
        "Artificial intelligence is transforming industries.")
    assert text_row("a") != text_row("b") and 0 <= text_row("a") < (1 << 40)
