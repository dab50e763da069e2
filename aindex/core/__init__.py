from aindex_b200.core.aindex import AIndex, Strand, get_revcomp, hamming_distance  # noqa: F401
from aindex_b200.core import aindex_cpp  # noqa: F401

__all__ = ["AIndex", "get_revcomp", "hamming_distance", "Strand", "aindex_cpp"]
