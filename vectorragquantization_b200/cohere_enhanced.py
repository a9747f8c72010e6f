"""``CohereEnhancedVectorDB`` (CohereEnhancedVectorDB.py) on the B200 path.

Same constructor, ``add_documents`` / ``search`` / ``remove_document`` / ``save`` / ``len`` and result dicts as the
reference.  The three phases of ``search`` (:267-322) run as CUDA kernels over data that stays in HBM:

  Phase I   faiss IndexBinaryFlat.search          -> TMA Hamming top-(k*binary_oversample) scan      (scan.cu)
  Phase II  Python loop: reconstruct + unpackbits + float.dot   -> in-register bit-select dot, f64  (rescore.cu)
  Phase III Python loop: RocksDB get + int8 "cosine"            -> gathered int8 dot / norm, f64    (rescore.cu)
  the three list.sort calls                                     -> per-query sort kernel             (rescore.cu)

The Cohere HTTP endpoint is out of scope (no network): embeddings come from an injectable ``embedder`` with the
call shape of ``_get_embeddings`` (default: the synthetic Cohere-like generator).  The reference insists on the
COHERE_EMBED_ENDPOINT / COHERE_EMBED_KEY environment variables (:67-75); that requirement is kept only when no
embedder is injected AND ``require_env=True`` is passed, so the class stays constructible offline.

Extensions: ``add_embeddings(doc_ids, int8, ubinary, docs=None)`` and ``search_batch(q_float, q_ubinary, ...)``.
"""
from __future__ import annotations

import json
import logging
import os
import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .binary_index import BinaryIndex, IndexBinaryFlat, read_index_binary, write_index_binary
from .docstore import DocStore
from .embedder import CohereHttpEmbedder, SyntheticCohereEmbedder, warn_synthetic

logger = logging.getLogger(__name__)


class CohereEnhancedVectorDB:
    def __init__(self, folder: str, model: str = "embed-english-v3.0", embedding_dim: int = 1024, index_type=IndexBinaryFlat,
                 index_args: List = None, rdict_options=None, embedder: Optional[Callable] = None, require_env: bool = False,
                 ctx=None):
        if index_args is None:
            index_args = [embedding_dim]
        if index_type is not IndexBinaryFlat:
            raise NotImplementedError("only IndexBinaryFlat (the reference's default and only used type) is supported")
        self.endpoint = os.environ.get("COHERE_EMBED_ENDPOINT")
        self.api_key = os.environ.get("COHERE_EMBED_KEY")
        if embedder is None and require_env:
            if not self.endpoint:
                raise Exception("COHERE_EMBED_ENDPOINT is not set in the environment.")
            if not self.api_key:
                raise Exception("COHERE_EMBED_KEY is not set in the environment.")
        if self.endpoint and "/v2/embed" not in self.endpoint:
            self.endpoint = self.endpoint.rstrip("/") + "/v2/embed"
        self.embedding_dim = embedding_dim
        self.model = model
        self.folder = folder
        self._ctx = ctx if ctx is not None else L.default_context()
        if embedder == "http":
            embedder = CohereHttpEmbedder(self.endpoint, self.api_key, model)
        elif embedder is None:
            warn_synthetic(type(self).__name__, "COHERE_EMBED_ENDPOINT / COHERE_EMBED_KEY")
            embedder = SyntheticCohereEmbedder(embedding_dim, ctx=self._ctx)
        self._embedder = embedder
        self._setup_config(folder, model, embedding_dim)
        self.doc_db = DocStore(os.path.join(folder, "docs"), rdict_options)
        self.index = self._initialize_faiss_index(folder, embedding_dim, index_type, index_args)
        self.doc_db.imported_raw = None  # the vectors of an imported reference store are in the index now

    # ---- config.json (:90-115): a mismatch is overwritten with a warning -------------------------------------
    def _setup_config(self, folder: str, model: str, embedding_dim: int):
        config_path = os.path.join(folder, "config.json")
        if not os.path.exists(config_path):
            if os.path.exists(folder) and os.listdir(folder):
                raise Exception(f"Folder {folder} contains files, but no config.json. "
                                "To create a new database, the folder must be empty.")
            os.makedirs(folder, exist_ok=True)
            with open(config_path, "w") as f:
                config = {"version": "1.0", "model": model, "embedding_dim": embedding_dim}
                json.dump(config, f)
        else:
            with open(config_path, "r") as f:
                config = json.load(f)
            if config.get("model") != model or config.get("embedding_dim") != embedding_dim:
                logger.warning("Config model or embedding_dim mismatch. Overwriting config.")
                config = {"version": "1.0", "model": model, "embedding_dim": embedding_dim}
                with open(config_path, "w") as fOut:
                    json.dump(config, fOut)
        self.config = config

    def _payload_path(self):
        return os.path.join(self.folder, "payload.vrqp")

    def _initialize_faiss_index(self, folder, embedding_dim, index_type, index_args) -> BinaryIndex:
        """:116-128.  index.bin is faiss's layout (codes + ids) and is streamed straight into device memory.  The int8 rows
        the reference keeps in RocksDB pickles come from, in this order: the streamed sidecar ``payload.vrqp`` written by
        ``save()``; round 1's ``int8.npy``; or - a folder written by the REFERENCE - its ``docs/`` store, imported
        read-only by rocks_import.py."""
        path = os.path.join(folder, "index.bin")
        if not os.path.exists(path):
            logger.info(f"New FAISS binary index created with embedding dimension {embedding_dim}.")
            return BinaryIndex(index_type(*index_args), ctx=self._ctx, payload_kind=L.PAYLOAD_INT8_RAW)
        index = read_index_binary(path, ctx=self._ctx)
        if index.d != embedding_dim:
            raise Exception(f"{path} holds {index.d}-bit codes, expected {embedding_dim}")
        n = index.ntotal
        legacy = os.path.join(folder, "int8.npy")
        if os.path.exists(self._payload_path()):
            index.read_payload(self._payload_path())
            if index.payload_kind != L.PAYLOAD_INT8_RAW:
                raise Exception(f"{self._payload_path()} does not hold int8 vectors")
        else:
            index.attach_payload(L.PAYLOAD_INT8_RAW)
            if n > 0 and os.path.exists(legacy):
                rows = np.load(legacy, mmap_mode="r")
                if rows.shape != (n, embedding_dim):
                    raise Exception(f"{legacy} has shape {rows.shape}, index.bin has {n} codes")
                for a in range(0, n, 65536):
                    index.write_rows(L.ROWS_PAYLOAD, a, np.ascontiguousarray(rows[a:a + 65536], np.int8))
            elif n > 0 and self.doc_db.imported_raw is not None:
                ids = index.read_rows(L.ROWS_IDS, 0, n)
                for a in range(0, n, 65536):
                    chunk = []
                    for i in ids[a:a + 65536]:
                        entry = self.doc_db.imported_raw.get(str(int(i)))
                        if entry is None or "int8" not in entry:
                            raise Exception(f"document {int(i)} of index.bin has no int8 vector in {folder}/docs")
                        chunk.append(np.asarray(entry["int8"], np.int8))
                    index.write_rows(L.ROWS_PAYLOAD, a, np.stack(chunk))
            elif n > 0:
                raise Exception(f"{folder}: index.bin has {n} codes but there are no int8 vectors (payload.vrqp / docs store)")
        logger.info("Existing FAISS binary index loaded.")
        return index

    def _to_binary(self, emb_int8: np.ndarray) -> np.ndarray:
        """:130-134 (dead code in the reference; kept for API parity)."""
        from . import kernels as K
        return K.to_binary(np.asarray(emb_int8, np.int8), ctx=self._ctx)

    def _get_embeddings(self, texts: List[str], input_type: str, embedding_types: List[str]) -> Dict:
        """:136-169 - returns {} on failure, like the reference."""
        try:
            return self._embedder(texts, input_type, embedding_types) or {}
        except Exception as e:
            logger.error("Embedding generation failed: %s", str(e))
            return {}

    # ---- add (:171-225) ----------------------------------------------------------------------------------------
    def add_documents(self, doc_ids: List[int], docs: List[str], batch_size: int = 64, save: bool = True):
        if len(doc_ids) != len(docs):
            raise ValueError("doc_ids and docs must have the same length.")
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        # one embedder call per batch like the reference (:195-199); the device-resident index is appended to for up to 4096
        # embedded documents at a time (a 64-row append is pure copy latency) - same order, ids and results
        pend_ids, pend_i8, pend_ub, pend_docs = [], [], [], []

        def flush():
            if pend_ids:
                self.index.add_with_ids(np.concatenate(pend_ub), np.array(pend_ids, dtype=np.int64), payload=np.concatenate(pend_i8))
                self.doc_db.set_many((str(doc_id), {"doc": doc}) for doc_id, doc in zip(pend_ids, pend_docs))
                pend_ids.clear(), pend_i8.clear(), pend_ub.clear(), pend_docs.clear()

        for start in range(0, len(docs), batch_size):
            batch_ids = doc_ids[start:start + batch_size]
            batch_docs = docs[start:start + batch_size]
            emb = self._get_embeddings(batch_docs, input_type="search_document", embedding_types=["int8", "ubinary"])
            if not emb:
                logger.error("Failed to retrieve embeddings for a batch.")
                continue
            try:
                int8_embs = np.array(emb["int8"], dtype=np.int8)
                ubinary_embs = np.array(emb["ubinary"], dtype=np.uint8)
                if int8_embs.shape != (len(batch_ids), self.embedding_dim) or ubinary_embs.shape != (len(batch_ids), self.embedding_dim // 8):
                    raise ValueError(f"unexpected embedding shapes {int8_embs.shape} / {ubinary_embs.shape}")
            except Exception as e:
                logger.error("Error processing embeddings: %s", str(e))
                continue
            pend_ids.extend(batch_ids), pend_i8.append(int8_embs), pend_ub.append(ubinary_embs), pend_docs.extend(batch_docs)
            if len(pend_ids) >= 4096:
                flush()
        flush()
        if save:
            self.save()

    def add_embeddings(self, doc_ids: Sequence[int], int8_embs: np.ndarray, ubinary_embs: np.ndarray,
                       docs: Optional[Sequence[str]] = None, save: bool = False):
        """Bulk path for precomputed Cohere-style embeddings (int8[n, D], ubinary[n, D/8])."""
        n = len(doc_ids)
        if int8_embs.shape != (n, self.embedding_dim) or ubinary_embs.shape != (n, self.embedding_dim // 8):
            raise ValueError("int8_embs must be int8[n, D] and ubinary_embs uint8[n, D/8]")
        if docs is not None and len(docs) != n:
            raise ValueError("doc_ids and docs must have the same length.")
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        self.index.add_with_ids(ubinary_embs, np.array(doc_ids, dtype=np.int64), payload=np.ascontiguousarray(int8_embs, np.int8))
        self.doc_db.set_many((str(doc_id), {"doc": docs[i] if docs is not None else ""}) for i, doc_id in enumerate(doc_ids))
        if save:
            self.save()

    # ---- search (:227-322) ---------------------------------------------------------------------------------------
    def search(self, query: str, k: int = 10, binary_oversample: int = 10, int8_oversample: int = 3) -> List[Dict]:
        if self.index.ntotal == 0:
            logger.error("No documents indexed. Please add documents before searching.")
            return []
        emb = self._get_embeddings([query], input_type="search_query", embedding_types=["float", "ubinary"])
        if not emb:
            logger.error("Query embedding generation failed.")
            return []
        try:
            query_float = np.array(emb["float"], dtype=np.float32)
            query_ubinary = np.array(emb["ubinary"], dtype=np.uint8)
        except Exception as e:
            logger.error("Error processing query embeddings: %s", str(e))
            return []
        t0 = time.time()
        labels, ham, sbin, scos, cnt = self.index.search3(query_float.reshape(1, -1), query_ubinary.reshape(1, -1), k,
                                                          binary_oversample, int8_oversample)
        logger.info("Phase I-III (Hamming scan, binary dot-product, int8 cosine) took %.2f ms", (time.time() - t0) * 1000)
        if cnt[0] == 0:
            logger.error("No candidates found.")
            return []
        results = []
        for i in range(int(cnt[0])):
            doc_entry = self.doc_db.get(str(int(labels[0, i])))
            if not doc_entry:
                continue
            results.append({"doc_id": int(labels[0, i]), "score_hamming": int(ham[0, i]), "score_binary": float(sbin[0, i]),
                            "score_cosine": float(scos[0, i]), "doc": doc_entry.get("doc", "N/A")})
        return results

    def search_batch(self, q_float: np.ndarray, q_ubinary: np.ndarray, k: int = 10, binary_oversample: int = 10,
                     int8_oversample: int = 3):
        """Batched phases I-III: (doc_ids i64[nq,k], hamming i32, score_binary f64, score_cosine f64, count i32[nq])."""
        return self.index.search3(q_float, q_ubinary, k, binary_oversample, int8_oversample)

    # ---- remove / save (:324-350) ----------------------------------------------------------------------------------
    def remove_document(self, doc_id: int, save: bool = True):
        doc_id_str = str(doc_id)
        if doc_id_str in self.doc_db:
            self.index.remove_ids(np.array([doc_id], dtype=np.int64))
            del self.doc_db[doc_id_str]
            logger.info(f"Document {doc_id} removed.")
        else:
            logger.warning(f"Document {doc_id} not found in the database.")
        if save:
            self.save()

    def save(self):
        """:342-347.  index.bin (faiss layout) + the int8 sidecar, both streamed from device memory in 64 MB chunks and
        renamed into place (no whole-matrix host copy, no second device copy)."""
        write_index_binary(self.index, os.path.join(self.folder, "index.bin"))
        self.index.write_payload(self._payload_path())
        legacy = os.path.join(self.folder, "int8.npy")
        if os.path.exists(legacy):
            os.remove(legacy)
        logger.info("FAISS binary index saved.")

    def __len__(self):
        return self.index.ntotal


def find_closest_document(db: CohereEnhancedVectorDB, query: str) -> Dict:
    """:355-360."""
    results = db.search(query, k=1)
    return results[0] if results else {}
