"""Builds the native parts of aindex_b200 in-tree.

  aindex_b200/libaindex_cuda.so            CUDA kernels + C-ABI (include/aindex_cuda.h), sm_100a
  aindex_b200/core/aindex_cpp<EXT_SUFFIX>  pybind11 module mirroring the reference's aindex_cpp
  aindex_b200/bin/{count_kmers13,compute_aindex,compute_aindex13,compute_reads}  drop-in executables

nvcc cross-compiles without a GPU; the .so files are git-ignored and travel to the GPU box
with the source snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libaindex_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("AIX_CXX", "/usr/bin/g++")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-ccbin", CXX, "--expt-relaxed-constexpr",
              "-diag-suppress", "177,550"] + os.environ.get("AIX_EXTRA_NVCC_FLAGS", "").split()
CU_SOURCES = ["ctx.cu", "mphf.cu", "tf_query.cu", "count13.cu", "codec.cu", "coverage.cu",
              "mphf_build.cu", "positions.cu", "radix_sort.cu", "multi.cu"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hs.append(os.path.join(ROOT, "include", "aindex_cuda.h"))
    return hs


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1] if cmd else ""))
    return r.stdout


def build_cuda(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in CU_SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(op)
        if force or _newer(op, [sp] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out:
                    print(out)
    if force or jobs or _newer(LIB, objs):
        _run([NVCC, "-shared", "-cudart", "static", "-ccbin", CXX, "-gencode",
              "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB


def build_pybind(force: bool = False) -> str | None:
    src = os.path.join(CSRC, "python_wrapper.cpp")
    if not os.path.exists(src):
        return None
    import pybind11
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(HERE, "core", "aindex_cpp" + ext)
    deps = [src, LIB] + _headers()
    if force or _newer(out, deps):
        _run([CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
              "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(),
              "-I" + os.path.join(ROOT, "include"), src, "-o", out,
              "-L" + HERE, "-laindex_cuda", "-Wl,-rpath,$ORIGIN/.."])
    return out


def build_bins(force: bool = False):
    outs = []
    bindir = os.path.join(HERE, "bin")
    for name in ("count_kmers13", "compute_aindex", "compute_aindex13", "compute_reads"):
        src = os.path.join(CSRC, "tools", name + ".cpp")
        if not os.path.exists(src):
            continue
        os.makedirs(bindir, exist_ok=True)
        out = os.path.join(bindir, name)
        if force or _newer(out, [src, LIB] + _headers()):
            _run([CXX, "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, src,
                  "-o", out, "-L" + HERE, "-laindex_cuda", "-Wl,-rpath,$ORIGIN/.."])
        outs.append(out)
    return outs


def build_all(verbose: bool = False, force: bool = False):
    lib = build_cuda(verbose=verbose, force=force)
    ext = build_pybind(force=force)
    bins = build_bins(force=force)
    return lib, ext, bins


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
