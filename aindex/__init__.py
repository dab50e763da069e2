"""Drop-in alias: `from aindex import AIndex` / `from aindex.core.aindex import AIndex` resolve to the
B200 implementation (aindex_b200), so code written against ad3002/aindex runs unchanged."""
from aindex_b200.core.aindex import AIndex  # noqa: F401

__version__ = "1.4.4+b200"
__all__ = ["AIndex"]
