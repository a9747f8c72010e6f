"""Module alias so that ``from VectorDBInt16Global import VectorDBInt16Global`` ports by changing only the package prefix."""
from .vectordb import VectorDBInt16Global  # noqa: F401
