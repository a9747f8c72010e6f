"""Module alias so that ``from CohereEnhancedVectorDB import CohereEnhancedVectorDB`` ports by changing only the package prefix."""
from .cohere_enhanced import CohereEnhancedVectorDB, find_closest_document  # noqa: F401
