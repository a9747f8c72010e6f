"""Module alias so that ``from VectorDBInt8 import VectorDBInt8`` ports by changing only the package prefix."""
from .vectordb import VectorDBInt8  # noqa: F401
