"""CPU tests of the host-only drop-in tools: compute_reads against the compiled reference binary (oracle/_ref/bin),
byte for byte, for the four input kinds (src/compute_reads.cpp:76-217)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "aindex_b200", "bin", "compute_reads")
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "compute_reads")


def _seq(rng, n):
    return bytes(rng.choice(list(b"ACGTNacgt"), size=n, p=[.23, .23, .23, .23, .04, .01, .01, .01, .01]).tolist())


@pytest.mark.parametrize("trailing_newline", [True, False])
def test_compute_reads_equals_reference_binary(tmp_path, trailing_newline):
    if not os.path.exists(OURS):
        from aindex_b200 import build
        build.build_all()
    if not os.path.exists(REF):
        pytest.skip("reference compute_reads not compiled under oracle/_ref")
    rng = np.random.default_rng(3 + trailing_newline)
    r1 = b"".join(b"@r%d/1\n%s\n+\n%s\n" % (i, _seq(rng, int(rng.integers(1, 80))), b"I" * 5) for i in range(400))
    r2 = b"".join(b"@r%d/2\n%s\n+\n%s\n" % (i, _seq(rng, int(rng.integers(1, 80))), b"I" * 5) for i in range(400))
    fa = b"".join(b">s%d some desc\n%s\n" % (i, b"\n".join(_seq(rng, 60) for _ in range(int(rng.integers(0, 4))))) for i in range(150))
    plain = b"".join(_seq(rng, int(rng.integers(0, 90))) + b"\n" for _ in range(300))
    if not trailing_newline:
        r1, r2, fa, plain = r1[:-1], r2[:-1], fa[:-1], plain[:-1]
    for name, data in (("r1.fq", r1), ("r2.fq", r2), ("x.fa", fa), ("p.txt", plain)):
        (tmp_path / name).write_bytes(data)
    cases = ((["r1.fq", "r2.fq", "fastq"], [".reads", ".ridx"]), (["r1.fq", "-", "se"], [".reads", ".ridx"]),
             (["x.fa", "-", "fasta"], [".reads", ".ridx", ".header"]), (["p.txt", "-", "reads"], [".ridx"]))
    for args, exts in cases:
        for tool, tag in ((REF, "ref"), (OURS, "ours")):
            cmd = [tool, str(tmp_path / args[0]), str(tmp_path / args[1]) if args[1] != "-" else "-", args[2],
                   str(tmp_path / "sub" / "deeper" / f"{tag}_{args[2]}")]
            r = subprocess.run(cmd, capture_output=True)
            assert r.returncode == 0, (cmd, r.stderr)
        for e in exts:
            a, b = (str(tmp_path / "sub" / "deeper" / f"{t}_{args[2]}{e}") for t in ("ref", "ours"))
            assert filecmp.cmp(a, b, shallow=False), (args, e)
    r = subprocess.run([OURS, str(tmp_path / "p.txt"), "-", "nonsense", str(tmp_path / "z")], capture_output=True)
    assert r.returncode == 2
