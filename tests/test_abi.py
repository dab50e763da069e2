"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/aindex_cuda.h declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from aindex_b200 import build
    build.build_cuda()
    return build.LIB


def _declared():
    src = open(os.path.join(ROOT, "include", "aindex_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aix_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in _declared() if s not in exported]
    assert not missing, f"declared in aindex_cuda.h but not exported: {missing}"


def test_ctypes_signatures_cover_header(built):
    from aindex_b200 import capi
    assert sorted(capi.SIGNATURES) == _declared()
    L = capi.lib()
    assert L.aix_version().startswith(b"aindex_b200")


def test_sm100a_code_present(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out


def test_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aindex_b200 import capi
    with pytest.raises(capi.AixError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (test infrastructure)."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "aindex_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle\b|oracle/|liboracle|aindex_oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_bench_contract_without_gpu():
    """bench.py on a box without a GPU: our arm refuses loudly (no CPU fallback), the reference arm (CPU only, no
    product code) prints the one-line JSON the contract asks for, and nothing else reaches stdout."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible: this is the no-GPU contract")
    except ImportError:
        pytest.skip("torch missing")
    # the reference arm needs no GPU and none of the product: at a small size it runs here end to end
    # (oracle canonical counter -> compute_mphf_seq -> compute_index -> ref_harness), one JSON line on stdout
    small = ["--reads", "20011", "--genome", "100003", "--queries", "200000", "--steps", "2", "--warmup", "1"]
    import bench
    import shutil
    import types
    cache = bench.ref_index_cache_dir(types.SimpleNamespace(reads=20011, genome=100003))
    try:
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference"] + small, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference"
        if "unavailable" not in d:  # oracle/_ref present (build container and GPU box)
            assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["config"]["queries_per_gpu"] == 200000
            assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["index_keys"] > 90000
            assert "aindex_b200" not in r.stderr
    finally:
        shutil.rmtree(cache, ignore_errors=True)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CUDA device" in r.stderr
