"""Module alias so that ``from VectorDBInt4 import VectorDBInt4`` ports by changing only the package prefix."""
from .vectordb import VectorDBInt4  # noqa: F401
