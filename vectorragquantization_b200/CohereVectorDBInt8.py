"""Module alias for the reference module name."""
from .cohere_variants import CohereVectorDBInt8  # noqa: F401
