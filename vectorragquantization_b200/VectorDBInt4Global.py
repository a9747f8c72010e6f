"""Module alias so that ``from VectorDBInt4Global import VectorDBInt4Global`` ports by changing only the package prefix."""
from .vectordb import VectorDBInt4Global  # noqa: F401
