"""Module alias so that ``from VectorDBInt8Global import VectorDBInt8Global`` ports by changing only the package prefix."""
from .vectordb import VectorDBInt8Global  # noqa: F401
