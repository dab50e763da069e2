"""aindex_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of ad3002/aindex.

Layout:
  csrc/        hand-written CUDA kernels + the C-ABI (include/aindex_cuda.h) + the pybind11
               module source (python_wrapper.cpp) + GPU command-line tools
  capi.py      ctypes binding of the C-ABI (numpy in / numpy out)
  core/        aindex_cpp (pybind11, class AindexWrapper) and aindex.py (class AIndex):
               the reference's Python surface on top of the CUDA library
  dist.py      sharding helpers for one-process-per-GPU runs (torch.distributed)
  build.py     in-tree build of the native parts

There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
