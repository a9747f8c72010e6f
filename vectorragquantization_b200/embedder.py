"""Synthetic embedding provider: replaces the HTTP embedding services (Ollama ``/api/embed`` at VectorDBInt8.py:82,
Cohere ``/v2/embed`` at CohereEnhancedVectorDB.py:163), which are out of scope (no network).

A text is mapped to a row number with a stable 64-bit hash; the row is produced by the counter-based generator
(``vrq_synth_f32`` / ``vrq_synth_codes_int8`` - CUDA, bit-identical to the oracle's generator).  The Cohere-like
variant returns the three types the Cohere API returns: float, int8 = clip(rint(1259 x - 0.69)), ubinary = x > 0
(SURVEY.md trap T4 - the semantics were probed from the reference's committed fixtures).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence

import numpy as np

from . import kernels as K


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
