"""Module alias so that ``from VectorDBInt16 import VectorDBInt16`` ports by changing only the package prefix."""
from .vectordb import VectorDBInt16  # noqa: F401
