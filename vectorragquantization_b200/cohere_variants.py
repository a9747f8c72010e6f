"""``CohereVectorDBInt8`` (CohereVectorDBInt8.py) and ``CohereVectorDBBinary`` (CohereVectorDBBinary.py) on the B200
path - the two remaining binary-index classes of the reference.  Both reuse the kernels of the
main path: the integer ``_to_binary`` threshold, the Hamming top-k scan, and the fused gather + dot rescoring with the
1-bit code itself as the payload (unpacked to +-1.0 in registers, never in memory).

The Cohere / Azure embedding services are out of scope (no network): embeddings come from an injectable ``embedder``.
``search_rerank_cohere`` (CohereVectorDBInt8.py:237-339) calls Cohere's rerank web service and is not provided.
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, List, Optional

import numpy as np

from . import _lib as L
from . import kernels as K
from .vectordb import _Progress, _VectorDBBase

logger = logging.getLogger(__name__)


class CohereVectorDBInt8(_VectorDBBase):
    """Cohere int8 embeddings -> packbits(int8 > mean) -> Hamming-only search (CohereVectorDBInt8.py:130-235)."""

    _payload_kind = L.PAYLOAD_NONE
    _desc = "Indexing docs (Int8)"

    def __init__(self, folder: str, model: str = "embed-english-v3.0", embedding_dim: int = 1024, rdict_options=None,
                 embedder: Optional[Callable] = None, ctx=None):
        if embedder is None:
            from .embedder import SyntheticCohereEmbedder, warn_synthetic
            warn_synthetic(type(self).__name__, "the Cohere endpoint")
            coh = SyntheticCohereEmbedder(embedding_dim, ctx=ctx)
            embedder = lambda texts, input_type="search_document": coh(texts, input_type, ["int8"])["int8"]  # noqa: E731
        super().__init__(folder, model, embedding_dim, rdict_options, "", embedder, ctx)
        self._int8_store: Dict[str, np.ndarray] = {}

    @staticmethod
    def _to_binary(embedding: np.ndarray) -> np.ndarray:
        """CohereVectorDBInt8.py:130-135."""
        return K.to_binary(np.asarray(embedding, np.int8))

    def _generate_int8_embeddings(self, texts: List[str], input_type: str = "search_document") -> Dict[str, np.ndarray]:
        """CohereVectorDBInt8.py:82-128: {text: int8[D]}; {} on failure."""
        try:
            try:
                e = np.asarray(self._embedder(list(texts), input_type))
            except TypeError:
                e = np.asarray(self._embedder(list(texts)))
        except Exception as ex:
            logger.error(f"Int8 embedding generation failed: {ex}")
            return {}
        if e.dtype != np.int8 or e.shape != (len(texts), self.embedding_dim):
            logger.error(f"Unexpected int8 embeddings: dtype {e.dtype}, shape {e.shape}")
            return {}
        return {t: e[i] for i, t in enumerate(texts)}

    def add_documents(self, doc_ids: List[int], docs: List[str], batch_size: int = 64, save: bool = True):
        if len(doc_ids) != len(docs):
            raise ValueError("doc_ids and docs must have the same length.")
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        with _Progress(len(docs), self._desc) as pbar:
            for start in range(0, len(docs), batch_size):
                batch_ids = doc_ids[start:start + batch_size]
                batch_texts = docs[start:start + batch_size]
                emb_map = self._generate_int8_embeddings(batch_texts, input_type="search_document")
                if not emb_map:
                    logger.error(f"Int8 embedding generation failed for batch: {batch_texts}")
                    pbar.update(len(batch_texts))
                    continue
                rows = np.stack([emb_map[t] for t in batch_texts])
                self.index.add_with_ids(K.to_binary(rows, ctx=self._ctx), np.array(batch_ids, dtype=np.int64))
                for d_id, text in zip(batch_ids, batch_texts):
                    self.doc_db[str(d_id)] = {"doc": text}
                    self._int8_store[str(d_id)] = emb_map[text]
                pbar.update(len(batch_texts))
        if save:
            self.save()

    def search(self, query: str, k: int = 10, binary_oversample: int = 10) -> List[Dict]:
        """CohereVectorDBInt8.py:192-235: Hamming top min(k*oversample, ntotal), stable sort by distance, first k."""
        if self.index.ntotal == 0:
            logger.error("No documents indexed. Please add documents before searching.")
            return []
        emb_map = self._generate_int8_embeddings([query], input_type="search_query")
        if not emb_map or query not in emb_map:
            logger.error("Query embedding generation failed. Returning empty results.")
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

    def search_rerank_cohere(self, *a, **kw):
        raise NotImplementedError("Cohere's rerank web service is out of scope (no network); see DESIGN.md section 0")

    def remove_document(self, doc_id: int, save: bool = True):
        doc_id_str = str(doc_id)
        if doc_id_str in self.doc_db:
            self.index.remove_ids(np.array([doc_id], dtype=np.int64))
            del self.doc_db[doc_id_str]
            self._int8_store.pop(doc_id_str, None)
            logger.info(f"Document {doc_id} removed.")
        else:
            logger.warning(f"Document {doc_id} not found in the database.")
        if save:
            self.save()


class CohereVectorDBBinary(_VectorDBBase):
    """float32 embeddings -> signed binary (x >= mean -> +1) packed into the index; 2-phase search whose rescoring is
    float32 dot(query, +-1 vector) or, with compare_float32=True, dot(query, float32 row)  (CohereVectorDBBinary.py)."""

    _payload_kind = L.PAYLOAD_CODES_PM1
    _desc = "Indexing docs (signed binary)"
    _ge = True

    def __init__(self, folder: str, model: str = "embed-english-v3.0", embedding_dim: int = 1024, rdict_options=None,
                 embedder: Optional[Callable] = None, ctx=None):
        super().__init__(folder, model, embedding_dim, rdict_options, "", embedder, ctx)

    @staticmethod
    def _to_signed_binary(embedding: np.ndarray) -> np.ndarray:
        """CohereVectorDBBinary.py:133-141: +1 where x >= mean(x), else -1 (int8)."""
        bits = np.unpackbits(K.to_binary(np.asarray(embedding, np.float32), ge=True))[: np.asarray(embedding).shape[-1]]
        return np.where(bits == 1, 1, -1).astype(np.int8)

    @staticmethod
    def _pack_signed_binary(signed_binary: np.ndarray) -> np.ndarray:
        """CohereVectorDBBinary.py:143-150."""
        return np.packbits(((np.asarray(signed_binary) + 1) // 2).astype(np.uint8))

    @staticmethod
    def _unpack_signed_binary(packed: np.ndarray, length: int) -> np.ndarray:
        """CohereVectorDBBinary.py:152-159."""
        bits = np.unpackbits(np.asarray(packed, np.uint8))[:length]
        return np.where(bits == 0, -1, 1).astype(np.float32)

    def _encode(self, x):
        return None, None, K.to_binary(x, ge=True, ctx=self._ctx)

    def _result_entry(self, x, payload, aux, ub):
        return {"float": x, "packed_binary": ub}


class CohereVectorDBFloat:
    """``CohereVectorDBFloat`` (CohereVectorDBFloat.py): float32 Cohere embeddings in ``faiss.IndexIDMap(faiss.IndexFlatIP)``,
    brute-force inner-product search - the recall yardstick the quantised classes are compared with (main.py:289-331).
    Same constructor, ``add_documents`` / ``search`` / ``remove_document`` / ``save`` / ``len``, ``config.json`` (no
    "version" key, :48) and ``index.faiss`` bytes as the reference; the rows live in HBM and ``search`` is two kernels
    (float_ip.cu).  Embeddings come from an injectable ``embedder(texts, input_type, ["float"]) -> {"float": rows}``
    (default: the synthetic Cohere-like generator; ``embedder="http"`` is the reference's Cohere call)."""

    def __init__(self, folder: str, model: str = "embed-english-v3.0", embedding_dim: int = 1024, rdict_options=None,
                 embedder: Optional[Callable] = None, ctx=None):
        import json
        import os

        from .binary_index import BinaryIndex, read_index_float
        from .docstore import DocStore
        self.embedding_dim, self.folder, self.model = embedding_dim, folder, model
        self.endpoint = os.environ.get("COHERE_EMBED_ENDPOINT")
        self.api_key = os.environ.get("COHERE_EMBED_KEY")
        self._ctx = ctx if ctx is not None else L.default_context()
        if embedder == "http":
            from .embedder import CohereHttpEmbedder
            ep = self.endpoint if not self.endpoint or "/v2/embed" in self.endpoint else self.endpoint.rstrip("/") + "/v2/embed"
            embedder = CohereHttpEmbedder(ep, self.api_key, model)
        elif embedder is None:
            from .embedder import SyntheticCohereEmbedder, warn_synthetic
            warn_synthetic(type(self).__name__, "COHERE_EMBED_ENDPOINT / COHERE_EMBED_KEY")
            embedder = SyntheticCohereEmbedder(embedding_dim, ctx=self._ctx)
        self._embedder = embedder
        config_path = os.path.join(folder, "config.json")
        if not os.path.exists(config_path):
            if os.path.exists(folder) and os.listdir(folder):
                raise Exception(f"Folder {folder} not empty but no config.json found. "
                                "To create new DB, folder must be empty or have config.json.")
            os.makedirs(folder, exist_ok=True)
            self.config = {"model": model, "embedding_dim": embedding_dim}
            with open(config_path, "w") as f:
                json.dump(self.config, f)
        else:
            with open(config_path, "r") as f:
                self.config = json.load(f)
        path = os.path.join(folder, "index.faiss")
        if os.path.exists(path):
            self.index = read_index_float(path, ctx=self._ctx)
            logger.info("Existing float FAISS index loaded.")
        else:
            self.index = BinaryIndex(embedding_dim, ctx=self._ctx, payload_kind=L.PAYLOAD_F32)
            logger.info(f"New float FAISS index created (dim={embedding_dim}).")
        self.doc_db = DocStore(os.path.join(folder, "docs"), rdict_options)
        self.doc_db.imported_raw = None

    def _generate_float_embeddings(self, texts: List[str], input_type: str) -> Dict[str, np.ndarray]:
        """:66-104 - {} on failure; rows of the wrong width are skipped with a warning."""
        try:
            rows = self._embedder(list(texts), input_type, ["float"])["float"]
        except Exception as e:
            logger.error(f"Float embedding request failed: {e}")
            return {}
        out = {}
        for i, txt in enumerate(texts):
            emb = np.asarray(rows[i], dtype=np.float32)
            if emb.ndim > 1:
                emb = emb[0]
            if emb.shape[0] != self.embedding_dim:
                logger.warning(f"Dimension mismatch for text '{txt}': got {emb.shape[0]}, want {self.embedding_dim}")
                continue
            out[txt] = emb
        return out

    def add_documents(self, doc_ids: List[int], docs: List[str], batch_size: int = 64, save: bool = True):
        if len(doc_ids) != len(docs):
            raise ValueError("doc_ids and docs must match length.")
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        with _Progress(len(docs), "Indexing docs (Float)") as pbar:
            for start in range(0, len(docs), batch_size):
                batch_ids, batch_txts = doc_ids[start:start + batch_size], docs[start:start + batch_size]
                emb_map = self._generate_float_embeddings(batch_txts, input_type="search_document")
                pairs = [(i, t) for i, t in zip(batch_ids, batch_txts) if t in emb_map]
                if pairs:
                    self.index.add_float_rows(np.vstack([emb_map[t] for _, t in pairs]), np.array([i for i, _ in pairs], dtype=np.int64))
                    self.doc_db.set_many((str(i), {"doc": t}) for i, t in pairs)
                pbar.update(len(batch_txts))
        if save:
            self.save()

    def add_embeddings(self, doc_ids, x: np.ndarray, docs=None, save: bool = False):
        """Bulk path: precomputed float32 rows [n, D]."""
        for doc_id in doc_ids:
            if str(doc_id) in self.doc_db:
                self.remove_document(doc_id, save=False)
        self.index.add_float_rows(x, np.asarray(doc_ids, dtype=np.int64))
        self.doc_db.set_many((str(i), {"doc": docs[j] if docs is not None else ""}) for j, i in enumerate(doc_ids))
        if save:
            self.save()

    def search(self, query: str, k: int = 10) -> List[Dict]:
        """:142-172 - dot-product search; results sorted by score descending."""
        if self.index.ntotal == 0:
            logger.warning("No docs in index, add documents first.")
            return []
        emb_map = self._generate_float_embeddings([query], input_type="search_query")
        if not emb_map or query not in emb_map:
            logger.error("Query embedding generation failed.")
            return []
        scores, ids = self.index.search_ip(emb_map[query].reshape(1, -1), k)
        results = []
        for dist, did in zip(scores[0], ids[0]):
            if did == -1:
                continue
            results.append({"doc_id": did, "score": float(dist), "doc": self.doc_db.get(str(did), {}).get("doc", "N/A")})
        results.sort(key=lambda x: x["score"], reverse=True)
        return results

    def search_batch(self, q_float: np.ndarray, k: int = 10):
        """(scores f32[nq,k] descending, doc ids i64[nq,k])."""
        return self.index.search_ip(q_float, k)

    def remove_document(self, doc_id: int, save: bool = True):
        doc_id_str = str(doc_id)
        if doc_id_str in self.doc_db:
            self.index.remove_ids(np.array([doc_id], dtype=np.int64))
            del self.doc_db[doc_id_str]
        if save:
            self.save()

    def save(self):
        import os

        from .binary_index import write_index_float
        write_index_float(self.index, os.path.join(self.folder, "index.faiss"))
        logger.info("Float FAISS index saved to disk.")

    def __len__(self):
        return self.index.ntotal
