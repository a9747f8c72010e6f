"""The reference's ``VectorDBInt4/Int8/Int16`` and ``*Global`` classes on the B200 path.

Same constructor arguments, methods, return shapes, config.json / index.bin layout and error behaviour as the
reference (VectorDBInt8.py, VectorDBInt8Global.py, VectorDBInt16Global.py, VectorDBInt4.py, VectorDBInt4Global.py,
VectorDBInt16.py).  What changed underneath:

  * ``faiss.IndexBinaryIDMap2``  -> ``BinaryIndex`` (device-resident codes, TMA Hamming top-k kernel)
  * the NumPy static methods      -> CUDA encoders / decoders (kernels.py); still callable as Class._fn(np.ndarray)
  * the per-candidate Python loop in ``search`` (RocksDB get + dequantise + np.dot, VectorDBInt8.py:226-240)
                                  -> one fused gather + dequantise + dot kernel and a per-query sort kernel
                                     (``BinaryIndex.search2``); the quantised vectors live next to the codes in HBM
  * ``rocksdict.Rdict``           -> ``DocStore`` (document text only)
  * ``requests.post`` to Ollama   -> an injectable ``embedder`` (default: synthetic, there is no network)

Additions the reference API cannot express (bulk entry points for the benchmark configurations):
``add_embeddings(doc_ids, x, docs=None)`` and ``search_batch(q_float, k, binary_oversample, compare_float32)``.
"""
from __future__ import annotations

import json
import logging
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from . import kernels as K
from .binary_index import BinaryIndex, read_index_binary, write_index_binary
from .docstore import DocStore
from .embedder import OllamaHttpEmbedder, SyntheticEmbedder, warn_synthetic

logger = logging.getLogger(__name__)

try:  # tqdm is what the reference shows progress with; optional here
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None


class _Progress:
    def __init__(self, total, desc):
        self.bar = _tqdm(total=total, desc=desc, disable=total < 2048) if _tqdm is not None else None

    def __enter__(self):
        return self

    def update(self, n):
        if self.bar is not None:
            self.bar.update(n)

    def __exit__(self, *a):
        if self.bar is not None:
            self.bar.close()


class _VectorDBBase:
    """Shared flow of the six classes (they are copies of one another in the reference, modulo the codec)."""

    _payload_kind = L.PAYLOAD_NONE
    _desc = "Indexing docs"
    _has_global_limit = False
    _ge = False  # '>=' mean threshold (CohereVectorDBBinary) instead of '>'
    _GPU_CHUNK = 4096  # embedded rows per encode + index-append call inside add_documents

    def __init__(self, folder: str, model: str = "snowflake-arctic-embed2", embedding_dim: int = 1024, rdict_options=None,
                 embed_url: str = "http://localhost:11434/api/embed", embedder: Optional[Callable] = None, ctx=None):
        self.embedding_dim = embedding_dim
        self.embed_url = embed_url
        self._ctx = ctx if ctx is not None else L.default_context()
        if embedder == "http":
            embedder = OllamaHttpEmbedder(embed_url, model)  # the reference's service call (VectorDBInt8.py:76-100)
        elif embedder is None:
            warn_synthetic(type(self).__name__, "embed_url")
            embedder = SyntheticEmbedder(embedding_dim, ctx=self._ctx)
        self._embedder = embedder
        self.folder = folder
        self._setup_config(folder, model, embedding_dim)
        self.doc_db = DocStore(os.path.join(folder, "docs"), rdict_options)
        self.index = self._initialize_faiss_index(folder, embedding_dim)
        self.doc_db.imported_raw = None  # the vectors of an imported reference store are in the index now
        self.float_embeddings: Dict[str, np.ndarray] = {}
        self._findex: Optional[BinaryIndex] = None  # codes + float32 rows, backs compare_float32=True

    # ---- config.json (VectorDBInt8.py:41-58, VectorDBInt8Global.py:50-73) ------------------------------------
    def _config_dict(self, model, embedding_dim):
        cfg = {"version": "1.0", "model": model, "embedding_dim": embedding_dim}
        if self._has_global_limit:
            cfg["global_limit"] = self.global_limit
        return cfg

    def _setup_config(self, folder: str, model: str, embedding_dim: int):
        config_path = os.path.join(folder, "config.json")
        if not os.path.exists(config_path):
            if os.path.exists(folder) and len(os.listdir(folder)) > 0:
                raise Exception(f"Folder {folder} contains files, but no config.json. "
                                "If you want to create a new database, the folder must be empty.")
            os.makedirs(folder, exist_ok=True)
            with open(config_path, "w") as f:
                json.dump(self._config_dict(model, embedding_dim), f)
        with open(config_path, "r") as f:
            self.config = json.load(f)
        if self._has_global_limit:
            # the stored limit wins over the constructor argument (VectorDBInt8Global.py:73)
            self.global_limit = float(self.config.get("global_limit", self.global_limit))

    # ---- index.bin + the quantised vectors -------------------------------------------------------------------
    def _limit(self) -> float:
        return float(getattr(self, "global_limit", 0.0))

    # key of the quantised vector inside the reference's per-document pickles (VectorDBInt8.py:179-183 and siblings)
    _ref_payload_key: Optional[str] = None

    def _initialize_faiss_index(self, folder: str, embedding_dim: int) -> BinaryIndex:
        """index.bin holds codes + ids only (faiss layout) and is streamed into device memory.  The quantised rows the
        reference keeps in RocksDB pickles come from the streamed sidecar ``payload.vrqp`` written by ``save()``, from
        round 1's ``payload.npz``, or - for a folder written by the REFERENCE - from its ``docs/`` store (rocks_import.py)."""
        path = os.path.join(folder, "index.bin")
        kind = self._payload_kind
        if not os.path.exists(path):
            logger.info(f"New FAISS index created with embedding dimension {embedding_dim}.")
            return BinaryIndex(embedding_dim, ctx=self._ctx, payload_kind=kind, global_limit=self._limit())
        index = read_index_binary(path, ctx=self._ctx)
        logger.info("Existing FAISS index loaded.")
        n = index.ntotal
        if kind == L.PAYLOAD_NONE:
            return index
        if kind == L.PAYLOAD_CODES_PM1:  # the rescoring payload is the code itself
            index.attach_payload(kind)
            return index
        legacy = os.path.join(folder, "payload.npz")
        if os.path.exists(self._payload_path()):
            index.read_payload(self._payload_path())
            if index.payload_kind != kind:
                raise Exception(f"{self._payload_path()} holds payload kind {index.payload_kind}, this class needs {kind}")
            return index
        index.attach_payload(kind, self._limit())
        dt, ln, adt = index.payload_layout()
        if n > 0 and os.path.exists(legacy):
            z = np.load(legacy)
            if z["payload"].shape[0] != n:
                raise Exception(f"{legacy} has {z['payload'].shape[0]} rows, index.bin has {n} codes")
            index.write_rows(L.ROWS_PAYLOAD, 0, np.ascontiguousarray(z["payload"], dt))
            if adt is not None:
                index.write_rows(L.ROWS_AUX, 0, np.ascontiguousarray(z["aux"], adt))
        elif n > 0 and self.doc_db.imported_raw is not None and self._ref_payload_key:
            ids = index.read_rows(L.ROWS_IDS, 0, n)
            for a in range(0, n, 65536):
                rows, aux = [], []
                for i in ids[a:a + 65536]:
                    entry = self.doc_db.imported_raw.get(str(int(i)))
                    if entry is None or self._ref_payload_key not in entry:
                        raise Exception(f"document {int(i)} of index.bin has no '{self._ref_payload_key}' vector in {folder}/docs")
                    rows.append(np.asarray(entry[self._ref_payload_key]).astype(dt, copy=False).reshape(ln))
                    if adt is not None:
                        aux.append([entry["min_max"][0], entry["min_max"][1]])
                index.write_rows(L.ROWS_PAYLOAD, a, np.stack(rows))
                if adt is not None:
                    index.write_rows(L.ROWS_AUX, a, np.asarray(aux, adt))
        elif n > 0:
            raise Exception(f"{folder}: index.bin has {n} codes but there are no quantised vectors (payload.vrqp / docs store)")
        return index

    def _payload_path(self):
        return os.path.join(self.folder, "payload.vrqp")

    # ---- embedding -----------------------------------------------------------------------------------------
    def _embed_float(self, texts: Sequence[str]) -> Optional[np.ndarray]:
        try:
            x = np.asarray(self._embedder(list(texts)), dtype=np.float32)
        except Exception as e:  # the reference logs and skips (VectorDBInt8.py:110-111)
            logger.error(f"Failed to generate embeddings. Error: {e}")
            return None
        if x.ndim == 1:
            x = x[None]
        if x.shape != (len(texts), self.embedding_dim):
            logger.error(f"Unexpected embedding shape: {x.shape}. Expected: ({len(texts)}, {self.embedding_dim}).")
            return None
        return x

    # codec hooks: encode a batch -> (payload rows, aux rows or None, ubinary), and the per-text result dict
    def _encode(self, x: np.ndarray) -> Tuple[np.ndarray, Optional[np.ndarray], np.ndarray]:
        raise NotImplementedError

    def _result_entry(self, x, payload, aux, ub) -> dict:
        raise NotImplementedError

    def _generate_embeddings(self, texts: List[str]) -> Dict[str, Dict[str, np.ndarray]]:
        x = self._embed_float(texts)
        if x is None:
            return {}
        payload, aux, ub = self._encode(x)
        return {t: self._result_entry(x[i], payload[i], None if aux is None else aux[i], ub[i]) for i, t in enumerate(texts)}

    # ---- add -------------------------------------------------------------------------------------------------
    def add_documents(self, doc_ids: List[int], docs: List[str], batch_size: int = 64, save: bool = True):
        if len(doc_ids) != len(docs):
            raise ValueError("doc_ids and docs must have the same length.")
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        # The embedder is called once per batch of `batch_size` texts, like the reference (:163-165); the GPU work (encode +
        # append to the device-resident index) is issued for up to _GPU_CHUNK embedded rows at a time: a 64-row call is pure
        # launch / copy latency.  Order, ids and results are those of per-batch insertion.
        pend_ids, pend_x, pend_docs = [], [], []

        def flush():
            if pend_ids:
                self._add_batch(pend_ids, np.concatenate(pend_x), pend_docs)
                pend_ids.clear(), pend_x.clear(), pend_docs.clear()

        with _Progress(len(docs), self._desc) as pbar:
            for start in range(0, len(docs), batch_size):
                batch_ids = doc_ids[start:start + batch_size]
                batch_docs = docs[start:start + batch_size]
                x = self._embed_float(batch_docs)
                if x is None:
                    logger.error(f"Embedding generation failed for batch: {batch_docs}")
                    continue
                pend_ids.extend(batch_ids), pend_x.append(x), pend_docs.extend(batch_docs)
                if len(pend_ids) >= self._GPU_CHUNK:
                    flush()
                pbar.update(len(batch_docs))
        flush()
        if save:
            self.save()

    def add_embeddings(self, doc_ids: Sequence[int], x: np.ndarray, docs: Optional[Sequence[str]] = None, save: bool = False,
                       keep_float: bool = True):
        """Bulk path: quantise + index precomputed float32 embeddings [n, D] (no embedder call)."""
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim != 2 or x.shape[1] != self.embedding_dim or x.shape[0] != len(doc_ids):
            raise ValueError("x must be float32[len(doc_ids), embedding_dim]")
        if docs is not None and len(docs) != len(doc_ids):
            raise ValueError("doc_ids and docs must have the same length.")
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        self._add_batch(list(doc_ids), x, docs, keep_float=keep_float)
        if save:
            self.save()

    def _add_batch(self, batch_ids, x, batch_docs, keep_float: bool = True):
        payload, aux, ub = self._encode(x)
        ids = np.array(batch_ids, dtype=np.int64)
        self.index.add_with_ids(ub, ids, payload=payload, aux=aux)
        if keep_float:
            if self._findex is None:
                self._findex = BinaryIndex(self.embedding_dim, ctx=self._ctx, payload_kind=L.PAYLOAD_F32)
            self._findex.add_with_ids(ub, ids, payload=x)
        self.doc_db.set_many((str(doc_id), {"doc": batch_docs[i] if batch_docs is not None else ""}) for i, doc_id in enumerate(batch_ids))
        if keep_float:
            for i, doc_id in enumerate(batch_ids):
                self.float_embeddings[str(doc_id)] = x[i]

    # ---- search ------------------------------------------------------------------------------------------------
    def _embed_query(self, query: str):
        x = self._embed_float([query])
        if x is None:
            return None
        return x[0]

    def search(self, query: str, k: int = 10, binary_oversample: int = 10, compare_float32: bool = False) -> List[Dict]:
        if self.index.ntotal == 0:
            logger.error("No documents indexed. Please add documents before searching.")
            return []
        qf = self._embed_query(query)
        if qf is None:
            logger.error("Query embedding generation failed. Returning empty results.")
            return []
        labels, scores, cnt = self.search_batch(qf[None], k, binary_oversample, compare_float32)
        out = []
        for doc_id, score in zip(labels[0][:cnt[0]], scores[0][:cnt[0]]):
            doc_data = self.doc_db.get(str(doc_id))
            if not doc_data:
                continue
            out.append({"doc_id": doc_id, "score": float(score), "doc": doc_data["doc"]})
        return out

    def search_batch(self, q_float: np.ndarray, k: int = 10, binary_oversample: int = 10, compare_float32: bool = False):
        """Batched ``search`` on precomputed query embeddings: (doc_ids i64[nq,k], scores f32[nq,k], count i32[nq])."""
        qf = np.ascontiguousarray(q_float, np.float32)
        # query_bin = self._to_binary(query float) (VectorDBInt8.py:213): with the `>` threshold the library derives it on the device
        # inside the search call; the `>=` variant (CohereVectorDBBinary.py:141) goes through its own call
        qb = K.to_binary(qf, ge=self._ge, ctx=self._ctx) if self._ge else None
        if compare_float32:
            if self._findex is None or self._findex.ntotal != self.index.ntotal:
                # the reference's float_embeddings dict is RAM-only and gone after a reopen: KeyError (VectorDBInt8.py:232)
                missing = next((k_ for k_ in self.doc_db.keys() if k_ not in self.float_embeddings), "?")
                raise KeyError(missing)
            return self._findex.search2(qf, qb, k, binary_oversample)
        return self.index.search2(qf, qb, k, binary_oversample)

    # ---- remove / save ---------------------------------------------------------------------------------------
    def remove_document(self, doc_id: int, save: bool = True):
        doc_id_str = str(doc_id)
        if doc_id_str in self.doc_db:
            self.index.remove_ids(np.array([doc_id], dtype=np.int64))
            if self._findex is not None:
                self._findex.remove_ids(np.array([doc_id], dtype=np.int64))
            del self.doc_db[doc_id_str]
            # (the reference does `del self.float_embeddings[id]` here, VectorDBInt8.py:252, and raises KeyError after a reopen
            # with the document already half removed; popping keeps remove / re-add usable after a reopen and with
            # add_embeddings(keep_float=False))
            self.float_embeddings.pop(doc_id_str, None)
            logger.info(f"Document {doc_id} removed from the database.")
        else:
            logger.warning(f"Document {doc_id} not found in the database.")
        if save:
            self.save()

    def save(self):
        """index.bin (faiss layout) + the payload sidecar, both streamed from device memory and renamed into place."""
        write_index_binary(self.index, os.path.join(self.folder, "index.bin"))
        if self._payload_kind not in (L.PAYLOAD_NONE, L.PAYLOAD_CODES_PM1):
            self.index.write_payload(self._payload_path())
            legacy = os.path.join(self.folder, "payload.npz")
            if os.path.exists(legacy):
                os.remove(legacy)
        logger.info("FAISS index saved to disk.")

    def __len__(self):
        return self.index.ntotal


# ------------------------------------------------------------------------------------------------------------------
class VectorDBInt8(_VectorDBBase):
    """Per-document symmetric int8 (VectorDBInt8.py): scale = 127/max|x| per vector, truncating cast."""

    _payload_kind = L.PAYLOAD_INT8_PERDOC
    _ref_payload_key = "emb_int8"
    _desc = "Indexing docs (Int8)"

    @staticmethod
    def _quantize_to_int8(embedding: np.ndarray):
        """VectorDBInt8.py:114-126 -> (int8[D], np.float32 min, np.float32 max)."""
        q, lo, hi = K.quantize_int8_perdoc(embedding)
        return q, lo, hi

    @staticmethod
    def _dequantize_int8(emb_int8: np.ndarray, min_max) -> np.ndarray:
        """VectorDBInt8.py:128-138."""
        return K.dequantize_int8_perdoc(emb_int8, np.float32(min_max[0]), np.float32(min_max[1]))

    @staticmethod
    def _to_binary(embedding: np.ndarray) -> np.ndarray:
        """VectorDBInt8.py:140-146."""
        return K.to_binary(embedding)

    def _encode(self, x):
        q, lo, hi, ub = K.quantize_int8_perdoc(x, want_binary=True, ctx=self._ctx)
        return q, np.stack([lo, hi], 1), ub

    def _result_entry(self, x, payload, aux, ub):
        return {"float": x, "ubinary": ub, "int8": payload, "min_max": (aux[0], aux[1])}


class VectorDBInt8Global(_VectorDBBase):
    """Global-limit int8 (VectorDBInt8Global.py): clip to +-limit, scale 127/limit, round half to even."""

    _payload_kind = L.PAYLOAD_INT8_GLOBAL
    _ref_payload_key = "emb_int8"
    _desc = "Indexing docs (Global Int8)"
    _has_global_limit = True

    def __init__(self, folder: str, model: str = "snowflake-arctic-embed2", embedding_dim: int = 1024,
                 global_limit: float = 0.3, rdict_options=None, embed_url: str = "http://localhost:11434/api/embed", **kw):
        self.global_limit = float(global_limit)
        super().__init__(folder, model, embedding_dim, rdict_options, embed_url, **kw)

    @staticmethod
    def _quantize_to_int8(embedding: np.ndarray, limit: float) -> np.ndarray:
        """VectorDBInt8Global.py:130-142."""
        return K.quantize_int8_global(embedding, limit)

    @staticmethod
    def _dequantize_int8(emb_int8: np.ndarray, limit: float) -> np.ndarray:
        """VectorDBInt8Global.py:144-152."""
        return K.dequantize_int8_global(emb_int8, limit)

    _to_binary = staticmethod(K.to_binary)

    def _encode(self, x):
        q, ub = K.quantize_int8_global(x, self.global_limit, want_binary=True, ctx=self._ctx)
        return q, None, ub

    def _result_entry(self, x, payload, aux, ub):
        return {"float": x, "ubinary": ub, "int8": payload}


class VectorDBInt16Global(_VectorDBBase):
    """Global-limit int16 (VectorDBInt16Global.py)."""

    _payload_kind = L.PAYLOAD_INT16_GLOBAL
    _ref_payload_key = "emb_int16"
    _desc = "Indexing docs (Global Int16)"
    _has_global_limit = True

    def __init__(self, folder: str, model: str = "snowflake-arctic-embed2", embedding_dim: int = 1024,
                 global_limit: float = 1.0, rdict_options=None, embed_url: str = "http://localhost:11434/api/embed", **kw):
        self.global_limit = float(global_limit)
        super().__init__(folder, model, embedding_dim, rdict_options, embed_url, **kw)

    @staticmethod
    def _quantize_to_int16(embedding: np.ndarray, limit: float) -> np.ndarray:
        """VectorDBInt16Global.py:130-142."""
        return K.quantize_int16_global(embedding, limit)

    @staticmethod
    def _dequantize_int16(emb_int16: np.ndarray, limit: float) -> np.ndarray:
        """VectorDBInt16Global.py:144-152."""
        return K.dequantize_int16_global(emb_int16, limit)

    _to_binary = staticmethod(K.to_binary)

    def _encode(self, x):
        q, ub = K.quantize_int16_global(x, self.global_limit, want_binary=True, ctx=self._ctx)
        return q, None, ub

    def _result_entry(self, x, payload, aux, ub):
        return {"float": x, "ubinary": ub, "int16": payload}


class VectorDBInt4(_VectorDBBase):
    """Per-document int4, two nibbles per byte (VectorDBInt4.py)."""

    _payload_kind = L.PAYLOAD_INT4_PERDOC
    _ref_payload_key = "emb_int4"
    _desc = "Indexing docs (Int4)"

    @staticmethod
    def _quantize_to_int4(embedding: np.ndarray):
        """VectorDBInt4.py:116-154 -> (packed int8[D/2], float min, float max)."""
        return K.quantize_int4(embedding)

    @staticmethod
    def _dequantize_int4(q_packed: np.ndarray, length: int, min_max) -> np.ndarray:
        """VectorDBInt4.py:156-184 (its NumPy-1.x result; the reference loop raises OverflowError on NumPy >= 2)."""
        return K.dequantize_int4_perdoc(q_packed, length, float(min_max[0]), float(min_max[1]))

    _to_binary = staticmethod(K.to_binary)

    def _encode(self, x):
        p, lo, hi, ub = K.quantize_int4(x, want_binary=True, ctx=self._ctx)
        return p, np.stack([lo, hi], 1), ub

    def _result_entry(self, x, payload, aux, ub):
        return {"float": x, "ubinary": ub, "int4": payload, "min_max": (float(aux[0]), float(aux[1]))}


class VectorDBInt4Global(_VectorDBBase):
    """"Global-limit" int4 (VectorDBInt4Global.py).  Faithful to the reference's code, not its docstring: the encoder
    IGNORES the limit and scales per document (VectorDBInt4Global.py:142-149), the decoder uses limit/7 (:177)."""

    _payload_kind = L.PAYLOAD_INT4_GLOBAL
    _ref_payload_key = "emb_int4"
    _desc = "Indexing docs (Global Int4)"
    _has_global_limit = True

    def __init__(self, folder: str, model: str = "snowflake-arctic-embed2", embedding_dim: int = 1024,
                 global_limit: float = 0.18, rdict_options=None, embed_url: str = "http://localhost:11434/api/embed", **kw):
        self.global_limit = float(global_limit)
        super().__init__(folder, model, embedding_dim, rdict_options, embed_url, **kw)

    @staticmethod
    def _quantize_to_int4(embedding: np.ndarray, limit: float) -> np.ndarray:
        """VectorDBInt4Global.py:129-164 (``limit`` unused, as in the reference)."""
        return K.quantize_int4(embedding)[0]

    @staticmethod
    def _dequantize_int4(q_packed: np.ndarray, length: int, limit: float) -> np.ndarray:
        """VectorDBInt4Global.py:166-188."""
        return K.dequantize_int4_global(q_packed, length, limit)

    _to_binary = staticmethod(K.to_binary)

    def _encode(self, x):
        p, _, _, ub = K.quantize_int4(x, want_binary=True, ctx=self._ctx)
        return p, None, ub

    def _result_entry(self, x, payload, aux, ub):
        return {"float": x, "ubinary": ub, "int4": payload}


class VectorDBInt16(_VectorDBBase):
    """VectorDBInt16.py: int16 embeddings come FROM THE SERVICE (there is no encoder in this class, SURVEY trap T3),
    1 bit/dim index, Hamming-only search (no rescoring, no compare_float32)."""

    _payload_kind = L.PAYLOAD_NONE
    _desc = "Indexing docs (Int16->1bit)"

    def __init__(self, folder: str, model: str = "snowflake-arctic-embed2", embedding_dim: int = 1024, rdict_options=None,
                 embed_url: str = "http://localhost:11434/api/embed", embedder: Optional[Callable] = None, ctx=None):
        super().__init__(folder, model, embedding_dim, rdict_options, embed_url, embedder, ctx)
        self.model = model
        self._int16_store: Dict[str, np.ndarray] = {}

    @staticmethod
    def _to_binary(embedding: np.ndarray) -> np.ndarray:
        """VectorDBInt16.py:148-157 (exact-mean threshold on int16)."""
        return K.to_binary(np.asarray(embedding, np.int16))

    def _generate_int16_embeddings(self, texts: List[str]) -> Dict[str, np.ndarray]:
        """VectorDBInt16.py:92-146.  The synthetic service returns round(clip(x, +-1) * 32767) of the synthetic float row."""
        if not texts:
            return {}
        try:
            e = np.asarray(self._embedder(list(texts)))
        except Exception as ex:
            logger.error(f"Int16 embedding generation failed: {ex}")
            return {}
        if e.dtype != np.int16:
            e = K.quantize_int16_global(np.asarray(e, np.float32), 1.0, ctx=self._ctx)
        if e.shape != (len(texts), self.embedding_dim):
            logger.error(f"Mismatch: got {e.shape[0]} embeddings for {len(texts)} texts.")
            return {}
        return {t: e[i] for i, t in enumerate(texts)}

    def add_documents(self, doc_ids: List[int], docs: List[str], batch_size: int = 64, save: bool = True):
        if len(doc_ids) != len(docs):
            raise ValueError("doc_ids and docs must have the same length.")
        for d_id in doc_ids:
            if str(d_id) in self.doc_db:
                self.remove_document(d_id, save=False)
        with _Progress(len(docs), self._desc) as pbar:
            for start in range(0, len(docs), batch_size):
                batch_ids = doc_ids[start:start + batch_size]
                batch_texts = docs[start:start + batch_size]
                emb_map = self._generate_int16_embeddings(batch_texts)
                if not emb_map:
                    logger.error("No embeddings returned for this batch.")
                    continue
                rows = np.stack([emb_map[t] for t in batch_texts])
                self.index.add_with_ids(K.to_binary(rows, ctx=self._ctx), np.array(batch_ids, dtype=np.int64))
                for d_id, text in zip(batch_ids, batch_texts):
                    self.doc_db[str(d_id)] = {"doc": text}
                    self._int16_store[str(d_id)] = emb_map[text]
                pbar.update(len(batch_texts))
        if save:
            self.save()

    def search(self, query: str, k: int = 10, binary_oversample: int = 10) -> List[Dict]:
        """VectorDBInt16.py:221-263: Hamming top min(k*oversample, ntotal), stable sort by distance, first k."""
        if self.index.ntotal == 0:
            logger.error("No documents indexed. Please add documents before searching.")
            return []
        emb_map = self._generate_int16_embeddings([query])
        if not emb_map or query not in emb_map:
            logger.error("Query embedding generation failed; returning empty.")
            return []
        query_bin = self._to_binary(emb_map[query])
        binary_k = min(k * binary_oversample, self.index.ntotal)
        distances, ids = self.index.search(query_bin.reshape(1, -1), binary_k)
        initial_hits = [(doc_id, dist) for doc_id, dist in zip(ids[0], distances[0]) if doc_id != -1]
        initial_hits.sort(key=lambda x: x[1])
        results = []
        for doc_id, dist in initial_hits[:k]:
            doc_data = self.doc_db.get(str(doc_id), {})
            results.append({"doc_id": doc_id, "score": dist, "doc": doc_data.get("doc", "N/A")})
        return results

    def remove_document(self, doc_id: int, save: bool = True):
        doc_id_str = str(doc_id)
        if doc_id_str in self.doc_db:
            self.index.remove_ids(np.array([doc_id], dtype=np.int64))
            del self.doc_db[doc_id_str]
            self._int16_store.pop(doc_id_str, None)
            logger.info(f"Document {doc_id} removed.")
        else:
            logger.warning(f"Document {doc_id} not found in the database.")
        if save:
            self.save()
