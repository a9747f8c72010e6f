"""Module name of the reference (``from CohereVectorDBFloat import CohereVectorDBFloat``)."""
from .cohere_variants import CohereVectorDBFloat  # noqa: F401
