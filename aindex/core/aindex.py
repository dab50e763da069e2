from aindex_b200.core.aindex import *  # noqa: F401,F403
from aindex_b200.core.aindex import AIndex, Strand, get_revcomp, hamming_distance  # noqa: F401
