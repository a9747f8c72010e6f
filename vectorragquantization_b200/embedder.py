"""Synthetic embedding provider: replaces the HTTP embedding services (Ollama ``/api/embed`` at VectorDBInt8.py:82,
Cohere ``/v2/embed`` at CohereEnhancedVectorDB.py:163), which are out of scope (no network).

A text is mapped to a row number with a stable 64-bit hash; the row is produced by the counter-based generator
(``vrq_synth_f32`` / ``vrq_synth_codes_int8`` - CUDA, bit-identical to the oracle's generator).  The Cohere-like
variant returns the three types the Cohere API returns: float, int8 = clip(rint(1259 x - 0.69)), ubinary = x > 0
(SURVEY.md trap T4 - the semantics were probed from the reference's committed fixtures).
"""
from __future__ import annotations

import hashlib
import logging
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import kernels as K

logger = logging.getLogger(__name__)
_warned = set()


def warn_synthetic(cls_name: str, what: str) -> None:
    """Logged once per class: the default embedder produces NON-SEMANTIC vectors (a hash of the text picks a row of the
    counter-based generator).  Pass ``embedder=`` (any callable) or ``embedder="http"`` to reach a real service."""
    if cls_name in _warned:
        return
    _warned.add(cls_name)
    logger.warning("%s: no embedder was injected - using the SYNTHETIC embedder (%s is ignored); search results are not "
                   "semantically meaningful.  Pass embedder=<callable> or embedder='http'.", cls_name, what)


class OllamaHttpEmbedder:
    """The reference's embedding call (VectorDBInt8.py:76-100): POST {"model", "input"} to ``embed_url``, read
    ``["embeddings"]``.  Needs a reachable service; selected with ``embedder="http"``."""

    def __init__(self, embed_url: str, model: str, timeout: float = 60.0):
        self.embed_url, self.model, self.timeout = embed_url, model, timeout

    def __call__(self, texts: Sequence[str]) -> np.ndarray:
        import requests
        rows = []
        for text in texts:  # one request per text, like the reference; both response shapes it accepts
            r = requests.post(self.embed_url, json={"model": self.model, "input": text}, timeout=self.timeout)
            r.raise_for_status()
            data = r.json()
            if data.get("data"):
                e = np.asarray(data["data"][0]["embedding"], dtype=np.float32)
            else:
                e = np.asarray(data["embeddings"], dtype=np.float32)
            rows.append(e[0] if e.ndim > 1 else e)
        return np.stack(rows)


class CohereHttpEmbedder:
    """The reference's Cohere call (CohereEnhancedVectorDB.py:136-169): POST to COHERE_EMBED_ENDPOINT /v2/embed with a
    bearer key; returns the ``embeddings`` object ({"float": ..., "int8": ..., "ubinary": ...})."""

    def __init__(self, endpoint: Optional[str], api_key: Optional[str], model: str, timeout: float = 60.0):
        if not endpoint or not api_key:
            raise Exception("COHERE_EMBED_ENDPOINT / COHERE_EMBED_KEY are not set in the environment.")
        self.endpoint, self.api_key, self.model, self.timeout = endpoint, api_key, model, timeout

    def __call__(self, texts: Sequence[str], input_type: str, embedding_types: List[str]) -> Dict:
        import requests
        r = requests.post(self.endpoint, headers={"Content-Type": "application/json", "Authorization": f"Bearer {self.api_key}"},
                          json={"model": self.model, "texts": list(texts), "input_type": input_type, "truncate": "NONE",
                                "embedding_types": embedding_types},
                          timeout=self.timeout)
        r.raise_for_status()
        return r.json().get("embeddings", {})


def text_row(text: str) -> int:
    """Stable (process-independent) 40-bit row number of a text."""
    return int.from_bytes(hashlib.blake2b(text.encode("utf-8"), digest_size=8).digest(), "little") & ((1 << 40) - 1)


class SyntheticEmbedder:
    """``embedder(texts) -> float32[len(texts), dim]`` for the VectorDBInt* classes."""

    def __init__(self, dim: int = 1024, seed: int = 1, row_scale: bool = False, ctx=None):
        self.dim, self.seed, self.row_scale, self.ctx = dim, seed, row_scale, ctx

    def __call__(self, texts: Sequence[str]) -> np.ndarray:
        out = np.empty((len(texts), self.dim), np.float32)
        for i, t in enumerate(texts):
            out[i] = K.synth_f32(self.seed, text_row(t), 1, self.dim, self.row_scale, ctx=self.ctx)[0]
        return out


class SyntheticCohereEmbedder:
    """``embedder(texts, input_type, embedding_types) -> {"float": ..., "int8": ..., "ubinary": ...}`` with the
    shapes of Cohere's ``embeddings`` object (lists are replaced by arrays)."""

    def __init__(self, dim: int = 1024, seed: int = 1, ctx=None):
        self.dim, self.seed, self.ctx = dim, seed, ctx

    def __call__(self, texts: Sequence[str], input_type: str, embedding_types: List[str]) -> Dict[str, np.ndarray]:
        n = len(texts)
        res: Dict[str, np.ndarray] = {}
        f = np.empty((n, self.dim), np.float32)
        i8 = np.empty((n, self.dim), np.int8)
        ub = np.empty((n, self.dim // 8), np.uint8)
        for i, t in enumerate(texts):
            r = text_row(t)
            if "float" in embedding_types:
                f[i] = K.synth_f32(self.seed, r, 1, self.dim, ctx=self.ctx)[0]
            if "int8" in embedding_types or "ubinary" in embedding_types:
                c, q = K.synth_codes_int8(self.seed, r, 1, self.dim, ctx=self.ctx)
                ub[i], i8[i] = c[0], q[0]
        if "float" in embedding_types:
            res["float"] = f
        if "int8" in embedding_types:
            res["int8"] = i8
        if "ubinary" in embedding_types:
            res["ubinary"] = ub
        return res
