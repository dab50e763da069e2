"""aindex_b200.core: the reference's Python surface (aindex/core) on the CUDA backend."""
from .aindex import AIndex, Strand, get_revcomp, hamming_distance  # noqa: F401

__all__ = ["AIndex", "get_revcomp", "hamming_distance", "Strand"]
