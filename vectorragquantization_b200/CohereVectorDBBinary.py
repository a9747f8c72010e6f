"""Module alias for the reference module name."""
from .cohere_variants import CohereVectorDBBinary  # noqa: F401
