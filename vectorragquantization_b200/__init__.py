"""B200-native quantise + multi-phase search path of aitrailblazer/VectorRAGQuantization.

Host side (this package) mirrors the reference's Python class API; every array operation on the hot path runs in
hand-written sm_100a CUDA kernels behind the C ABI in ``include/vrq.h`` (``libvrq.so``, bound with ctypes).
There is no CPU fallback: without the built library or without a GPU the first kernel call raises ``VrqError``.
"""
from ._lib import Context, VrqError, default_context  # noqa: F401
from .binary_index import (BinaryIndex, IndexBinaryFlat, IndexBinaryIDMap2, read_index_binary,  # noqa: F401
                           read_index_float, write_index_binary, write_index_float)

__version__ = "0.1.0"
from .cohere_enhanced import CohereEnhancedVectorDB  # noqa: F401,E402
from .docstore import DocStore, Rdict  # noqa: F401,E402
from .embedder import SyntheticCohereEmbedder, SyntheticEmbedder  # noqa: F401,E402
from .vectordb import (VectorDBInt4, VectorDBInt4Global, VectorDBInt8, VectorDBInt8Global, VectorDBInt16,  # noqa: F401,E402
                       VectorDBInt16Global)
from .cohere_variants import CohereVectorDBBinary, CohereVectorDBFloat, CohereVectorDBInt8  # noqa: F401,E402
