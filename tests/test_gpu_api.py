"""GPU tests of the drop-in surface: the pybind11 module aindex_cpp.AindexWrapper, the AIndex
Python class and the command-line tools, against answers recorded from the unmodified reference
(tests/golden).  These read like the reference's own test_aindex_functionality.py checks."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpp():
    from aindex_b200.core import aindex_cpp
    assert aindex_cpp.backend == "cuda-sm_100a"
    return aindex_cpp


@pytest.fixture(scope="module")
def g23(golden_dir):
    return np.load(os.path.join(golden_dir, "golden23.npz"))


@pytest.fixture(scope="module")
def g13(golden_dir):
    return np.load(os.path.join(golden_dir, "golden13.npz"))


@pytest.fixture(scope="module")
def wrapper(cpp, golden_dir):
    w = cpp.AindexWrapper()
    prefix = os.path.join(golden_dir, "idx23")
    w.load_from_prefix_23mer(prefix, prefix + ".reads")
    w.load_aindex_from_prefix_23mer(prefix, 100000)
    return w


def _q(g):
    return [g["recs"][i, :g["lens"][i]].tobytes().decode("latin-1") for i in range(g["recs"].shape[0])]


def test_wrapper_queries_match_reference(wrapper, g23):
    q = _q(g23)
    assert wrapper.n_kmers == int(g23["n_kmers"][0]) == wrapper.get_hash_size()
    assert wrapper.get_tf_values(q) == g23["tf"].tolist()
    assert wrapper.get_tf_values_23mer(q) == g23["tf"].tolist()
    assert wrapper.get_total_tf_values_23mer(q) == g23["total"].tolist()
    assert [list(p) for p in wrapper.get_tf_both_directions_23mer_batch(q)] == g23["both"].tolist()
    assert wrapper.get_hash_values(q) == g23["hash"].tolist()
    for i in range(0, len(q), 41):
        assert wrapper.get_tf_value(q[i]) == int(g23["tf"][i])
        assert wrapper.get_kid_by_kmer(q[i]) == int(g23["kid"][i])
        assert wrapper.get_strand(q[i]) == int(g23["strand"][i])
        assert wrapper.get_total_tf_value_23mer(q[i]) == int(g23["total"][i])
        assert list(wrapper.get_tf_both_directions_23mer(q[i])) == g23["both"][i].tolist()
        assert wrapper.get_hash_value(q[i]) == int(g23["hash"][i])
    # buffer overload: uint8[q, 23] in, uint32 ndarray out
    sel = [i for i in range(len(q)) if g23["lens"][i] == 23]
    arr = np.ascontiguousarray(g23["recs"][sel, :23])
    out = wrapper.get_tf_values(arr)
    assert isinstance(out, np.ndarray) and out.dtype == np.uint32
    assert np.array_equal(out, g23["tf"][sel])
    assert wrapper.get_tf_values([]) == []


def test_wrapper_kid_info_and_revcomp(wrapper, g23, golden_dir):
    for kid, tf, km, rk in zip(g23["info_kid"], g23["info_tf"], g23["info_kmer"], g23["info_rkmer"]):
        assert wrapper.get_kmer_info(int(kid)) == (int(tf), km.decode(), rk.decode())
        assert wrapper.get_kmer_by_kid(int(kid)) == km.decode()
        assert wrapper.get_kid_by_kmer(km.decode()) == int(kid)  # round trip, test_aindex_functionality.py:770-775
    assert wrapper.get_kmer_by_kid(10 ** 9) == "" and wrapper.get_kmer_info(10 ** 9) == (0, "", "")
    kat = np.load(os.path.join(golden_dir, "golden_kat.npz"))
    for s, r in zip(kat["rc23_in"], kat["rc23_out"]):
        assert wrapper.get_reverse_complement_23mer(s.decode()) == r.decode()
    for s, r in zip(kat["rc13_in"], kat["rc13_out"]):
        assert wrapper.get_reverse_complement_13mer(s.decode()) == r.decode()
    assert wrapper.get_reverse_complement_23mer("ACGT") == ""
    assert "23-mer Index Statistics" in wrapper.get_23mer_statistics()
    assert "Mode: 23-mer" in wrapper.get_index_info()


def test_wrapper_positions_and_reads(wrapper, g23, golden_dir):
    q = _q(g23)
    for j, qi in enumerate(g23["pos_qidx"]):
        want = g23["pos_val"][int(g23["pos_off"][j]):int(g23["pos_off"][j + 1])].tolist()
        assert wrapper.get_positions(q[qi]) == want
        assert len(want) == int(g23["tf"][qi])  # test_aindex_functionality.py:376-380
    offs, vals = wrapper.get_positions_batch([q[i] for i in g23["pos_qidx"]], 23)
    assert np.array_equal(offs, g23["pos_off"]) and np.array_equal(vals, g23["pos_val"])
    # absent k-mer / unsupported length: [] (the reference aborts on the first)
    assert wrapper.get_positions("A" * 23) == [] or wrapper.get_tf_value("A" * 23) > 0
    assert wrapper.get_positions("ACGT") == []
    # positions point at the k-mer or its reverse complement (test_aindex_functionality.py:541-558)
    reads = open(os.path.join(golden_dir, "idx23.reads"), "rb").read().decode()
    comp = str.maketrans("ACGT", "TGCA")
    for qi in g23["pos_qidx"][:50]:
        for p in wrapper.get_positions(q[qi]):
            sub = reads[p:p + 23]
            assert sub == q[qi] or sub.translate(comp)[::-1] == q[qi]
    gr = np.load(os.path.join(golden_dir, "golden_reads.npz"))
    assert wrapper.n_reads == int(gr["n_reads"][0]) and wrapper.get_reads_size() == int(gr["reads_size"][0])
    assert [wrapper.get_rid(int(p)) for p in gr["pos"]] == gr["rid"].tolist()
    assert [wrapper.get_start(int(p)) for p in gr["pos"]] == gr["start"].tolist()
    assert [wrapper.get_read_by_rid(int(r)).encode() for r in gr["read_rid"]] == gr["read_str"].tolist()
    assert [wrapper.get_read(int(a), int(b), bool(c)).encode() for a, b, c in gr["span"]] == gr["span_str"].tolist()
    # batched forms agree with the single calls (which agree with the reference module's recorded answers)
    rid_b, start_b = wrapper.get_rids_and_starts(gr["pos"].astype(np.uint64))
    assert rid_b.tolist() == gr["rid"].tolist() and start_b.tolist() == gr["start"].tolist()
    rr = np.concatenate([gr["read_rid"].astype(np.uint64), np.array([10**9], dtype=np.uint64)])  # + an unknown rid -> ""
    blob, offs = wrapper.get_reads_by_rids(rr)
    assert [blob[int(offs[i]):int(offs[i + 1])] for i in range(rr.size)] == gr["read_str"].tolist() + [b""]
    some = q[int(g23["pos_qidx"][0])]
    got = wrapper.get_reads_se_by_kmer(some, 5)
    assert 1 <= len(got) <= 5 and all(some in r or some.translate(comp)[::-1] in r for r in got)


def test_wrapper_errors(cpp, tmp_path):
    w = cpp.AindexWrapper()
    with pytest.raises(FileNotFoundError):
        w.load_from_prefix_23mer(str(tmp_path / "missing"))
    with pytest.raises(FileNotFoundError):
        w.load_13mer_index(str(tmp_path / "a.pf"), str(tmp_path / "b.tf.bin"))
    with pytest.raises(RuntimeError):
        w.get_tf_values(["A" * 23])  # nothing loaded: loud failure, not a silent zero from a CPU path
    assert w.get_total_tf_values_13mer(["A" * 13]) == [0]  # 13-mer API outside 13-mer mode: zeros (:523-526)


def test_aindex_class(golden_dir, g23):
    from aindex import AIndex  # the drop-in import path
    from aindex_b200.core.aindex import Strand, get_revcomp, hamming_distance
    prefix = os.path.join(golden_dir, "idx23")
    ix = AIndex.load_from_prefix(prefix, max_tf=100000, load_reads=True)  # auto-detects 23-mers
    q = _q(g23)
    assert ix.k == 23 and len(ix) == int(g23["n_kmers"][0]) and ix.aindex_loaded
    assert ix.get_tf_values(q) == g23["tf"].tolist()
    hit = int(np.nonzero(g23["tf"])[0][0])
    assert ix[q[hit]] == int(g23["tf"][hit]) and q[hit] in ix and ix.get("A" * 23, 7) in (7, ix["A" * 23])
    assert ix.get_strand(q[hit]) in (Strand.FORWARD, Strand.REVERSE)
    so, co = g23["cov_seq_off"], g23["cov_off"]
    seqs = [g23["cov_seq"][so[j]:so[j + 1]].tobytes().decode("latin-1") for j in range(len(so) - 1)]
    for j, s in enumerate(seqs):
        want = g23["cov_val"][co[j]:co[j + 1]]
        assert ix.get_sequence_coverage(s) == want.tolist()
        assert ix.get_sequence_coverage(s, cutoff=4) == np.where(want >= 4, want, 0).tolist()
    offs, cov = ix.get_sequence_coverage_batch(seqs)
    assert np.array_equal(cov, g23["cov_val"]) and np.array_equal(offs, co)
    p0 = int(g23["pos_qidx"][0])
    assert ix.pos(q[p0]) == g23["pos_val"][:int(g23["pos_off"][1])].tolist()
    with pytest.raises(ValueError):
        ix.get_positions("ACGT")
    rid, read = next(iter(ix.iter_reads()))
    assert rid == 0 and read == ix.get_read_by_rid(0)
    assert get_revcomp("ATCGN") == "NCGAT" and hamming_distance("ACGT", "ACNA") == 1
    kmers = list(ix.iter_sequence_kmers(seqs[0]))
    assert len(kmers) == len(seqs[0]) - 22 and kmers[0][1] == int(g23["cov_val"][0])
    # frequency views (aindex.py:594-795): top k-mers by tf, summary statistics
    tf_all = np.fromfile(prefix + ".tf.bin", dtype=np.uint32)
    top = ix.get_top_kmers(25)
    assert [t for _, t in top] == sorted(tf_all.tolist(), reverse=True)[:25]
    assert all(ix[k] == t for k, t in top) and len({k for k, _ in top}) == 25
    assert [t for _, t in ix.iter_kmers_by_frequency(min_tf=int(top[3][1]))] == [t for t in sorted(tf_all.tolist(), reverse=True) if t >= top[3][1]]
    st = ix.get_kmer_frequency_stats()
    nz = tf_all[tf_all > 0]
    assert st["kmer_type"] == "23mer" and st["total_kmers"] == tf_all.size and st["non_zero_kmers"] == nz.size
    assert st["max_tf"] == int(nz.max()) and st["min_tf"] == int(nz.min()) and st["total_tf"] == int(tf_all.sum())
    assert abs(st["avg_tf"] - nz.mean()) < 1e-9


@pytest.fixture(scope="module")
def tf13_file(g13, tmp_path_factory):
    d = tmp_path_factory.mktemp("idx13")
    tf = np.zeros(1 << 26, dtype=np.uint64)
    tf[g13["plain_ids"]] = g13["plain_counts"]
    p = str(d / "r13.tf.bin")
    tf.tofile(p)
    return p


def test_wrapper_13mer_mode(cpp, pf13, tf13_file, g13, tmp_path):
    w = cpp.AindexWrapper()
    w.load_13mer_index(pf13, tf13_file)
    q = _q({"recs": g13["q_recs"], "lens": g13["q_lens"]})
    assert w.get_tf_values(q) == g13["q_tf"].tolist()
    assert w.get_tf_values_13mer(q) == g13["q_tf"].tolist()
    ok = [q[i] for i in g13["q_ok"]]
    assert w.get_total_tf_values_13mer(ok) == g13["q_total"].tolist()
    assert [list(p) for p in w.get_tf_both_directions_13mer_batch(ok)] == g13["q_both"].tolist()
    assert w.get_total_tf_value_13mer(ok[0]) == int(g13["q_total"][0])
    st = w.get_13mer_statistics()
    assert st["total_kmers"] == 1 << 26 and st["non_zero_kmers"] == g13["plain_ids"].size
    assert st["total_count"] == int(g13["plain_counts"].sum())
    assert w.get_tf_by_index_13mer(int(g13["plain_ids"][0])) == int(g13["plain_counts"][0])
    assert w.get_hash_size() == 1 << 26 and "Mode: 13-mer" in w.get_index_info()
    # frequency iterator, 13-mer mode: the direct-address array names the right k-mers
    direct = np.asarray(w.get_13mer_tf_array_direct())
    assert direct.dtype == np.uint64 and direct.size == 1 << 26 and int(direct.sum()) == int(g13["plain_counts"].sum())
    from aindex_b200.core.aindex import AIndex
    a13 = AIndex()
    a13._wrapper, a13._loaded, a13.k = w, True, 13
    top = a13.get_top_kmers(40)
    assert [t for _, t in top] == sorted(g13["plain_counts"].tolist(), reverse=True)[:40]
    assert all(w.get_tf_value(k) == t for k, t in top) and len({k for k, _ in top}) == 40
    st13 = a13.get_kmer_frequency_stats("13mer")
    assert st13["non_zero_kmers"] == g13["plain_ids"].size and st13["total_tf"] == int(g13["plain_counts"].sum())
    # positions: build with the GPU (compute_aindex13 semantics), load back, query
    reads = tmp_path / "r13.reads"
    reads.write_bytes(g13["plain_data"].tobytes())
    w.build_positions(str(reads), str(tmp_path / "r13.index.bin"), str(tmp_path / "r13.indices.bin"), 13)
    pos = np.fromfile(tmp_path / "r13.index.bin", dtype=np.uint64)
    assert np.array_equal(pos, g13["pos13_positions"])
    assert hashlib.md5((tmp_path / "r13.indices.bin").read_bytes()).hexdigest() == str(g13["pos13_indices_md5"])
    w.load_13mer_aindex(str(tmp_path / "r13.index.bin"), str(tmp_path / "r13.indices.bin"))
    data = g13["plain_data"].tobytes().decode("latin-1")
    line = next(l for l in data.split("\n") if len(l) >= 13 and not l[:13].strip("ACGT"))
    got = w.get_positions_13mer(line[:13])
    assert got and all(data[p:p + 13] == line[:13] for p in got) and len(got) == w.get_tf_value(line[:13])
    assert w.get_positions(line[:13]) == got
    # count_kmers13 through the module: file identical to the reference's output
    stats = w.count_kmers13(str(reads), pf13, str(tmp_path / "c.tf.bin"))
    assert [stats["sequences"], stats["total_kmers"], stats["valid_kmers"], stats["invalid_kmers"]] == g13["plain_stats"].tolist()
    assert hashlib.md5((tmp_path / "c.tf.bin").read_bytes()).hexdigest() == str(g13["plain_md5"])


def test_cli_tools(pf13, g13, golden_dir, tmp_path):
    bindir = os.path.join(ROOT, "aindex_b200", "bin")
    for name in ("fasta", "fastq"):
        src = tmp_path / f"in.{name}"
        src.write_bytes(g13[f"{name}_data"].tobytes())
        out = tmp_path / f"{name}.tf.bin"
        r = subprocess.run([os.path.join(bindir, "count_kmers13"), str(src), pf13, str(out), "4"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert hashlib.md5(out.read_bytes()).hexdigest() == str(g13[f"{name}_md5"])
        assert f"Valid k-mers: {int(g13[f'{name}_stats'][2])}" in r.stdout
    p = os.path.join(golden_dir, "idx23")
    r = subprocess.run([os.path.join(bindir, "compute_aindex"), p + ".reads", p + ".pf", str(tmp_path / "o"), "4", "23",
                        p + ".tf.bin", p + ".kmers.bin", "none"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.index.bin").read_bytes() == open(p + ".index.bin", "rb").read()
    assert (tmp_path / "o.indices.bin").read_bytes() == open(p + ".indices.bin", "rb").read()
    tf = np.zeros(1 << 26, dtype=np.uint64)
    tf[g13["plain_ids"]] = g13["plain_counts"]
    tf.tofile(tmp_path / "p.tf.bin")
    (tmp_path / "p.reads").write_bytes(g13["plain_data"].tobytes())
    r = subprocess.run([os.path.join(bindir, "compute_aindex13"), str(tmp_path / "p.reads"), pf13, str(tmp_path / "p.tf.bin"),
                        str(tmp_path / "p13"), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "p13.index.bin", dtype=np.uint64), g13["pos13_positions"])


def test_build_index_from_reads_roundtrip(cpp, golden_dir, tmp_path, oracle):
    """GPU-built index files are loadable by the reference's loader format and answer like the
    reference-built ones (the k-mer ids differ, the answers do not)."""
    w = cpp.AindexWrapper()
    prefix = str(tmp_path / "gpu23")
    n = w.build_index_from_reads(os.path.join(golden_dir, "idx23.reads"), prefix)
    ref = oracle.Index23.load_prefix(os.path.join(golden_dir, "idx23"))
    assert n == ref.n
    mine = oracle.Index23.load_prefix(prefix)  # oracle = the reference's lookup on OUR files
    assert sorted(mine.checker.tolist()) == sorted(ref.checker.tolist())
    km = np.array([list(oracle.bitset_dna23(int(x)).encode()) for x in ref.checker[:2000]], dtype=np.uint8)
    assert np.array_equal(mine.batch(km), ref.batch(km))
    w.load_from_prefix_23mer(prefix)
    assert w.get_tf_values(km).tolist() == ref.batch(km).tolist()


def test_wrapper_is_safe_from_python_threads(cpp, golden_dir):
    """The reference holds the GIL for every call, so one wrapper may be shared by threads.  This module releases the
    GIL around batch calls and therefore guards its aix_ctx with a mutex: batch and single calls from several
    threads on ONE wrapper neither race nor deadlock."""
    import threading
    p = os.path.join(golden_dir, "idx23")
    w = cpp.AindexWrapper()
    w.load(p + ".pf", p + ".tf.bin", p + ".kmers.bin", "")
    tf = np.fromfile(p + ".tf.bin", dtype=np.uint32)
    ids = list(range(0, tf.size, 3))
    strs = [w.get_kmer_by_kid(i) for i in ids]
    want = [int(tf[i]) for i in ids]
    errs = []

    def batch():
        for _ in range(20):
            if w.get_tf_values(strs) != want:
                errs.append("batch")

    def single():
        for k, t in list(zip(strs, want))[:2000]:
            if w.get_tf_value(k) != t:
                errs.append("single")

    ts = [threading.Thread(target=batch), threading.Thread(target=single), threading.Thread(target=batch)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs and not any(t.is_alive() for t in ts)


def test_multi_gpu_count_and_queries():
    """2 ranks on 2 GPUs (skipped on a single-GPU box): NCCL reduce-scatter path == single GPU."""
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _multi(capi, n, dev_ids=None):
    import ctypes as C
    lib = capi.lib()
    h = C.c_void_p()
    ids = (C.c_int * n)(*dev_ids) if dev_ids else None
    rc = lib.aix_multi_create(n, ids, C.byref(h))
    assert rc == 0, lib.aix_multi_last_error(None)
    return h


@pytest.mark.parametrize("ranks", [2, 4])
def test_count13_multi_single_process(pf13, golden_dir, tmp_path, ranks):
    """aix_count13_multi: N contexts / host threads in ONE process, sharded input, exchange over peer pointers, permute.
    On a 1-GPU box the N contexts share device 0 (the peer pointers are then ordinary device pointers), which exercises
    the same sharding, exchange and permutation code; with >= N GPUs it runs one context per GPU.  Results must equal the
    single-context path and the reference's count_kmers13 goldens, for plain text, FASTQ and FASTA."""
    import ctypes as C
    import torch
    from aindex_b200 import capi
    lib = capi.lib()
    n_dev = torch.cuda.device_count()
    ids = list(range(ranks)) if n_dev >= ranks else [0] * ranks
    mg = _multi(capi, ranks, ids)
    try:
        assert lib.aix_multi_size(mg) == ranks
        ctx0 = capi.Context.__new__(capi.Context)
        ctx0._h = C.c_void_p(lib.aix_multi_ctx(mg, 0))
        m13 = capi.Mphf.load(ctx0, pf13)
        single = capi.Context(0)
        m13s = capi.Mphf.load(single, pf13)
        rng = np.random.default_rng(17)
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        genome = rng.choice(acgt, size=30_000)
        reads = []
        for i in range(4000):
            st = int(rng.integers(0, genome.size - 120))
            r = genome[st:st + int(rng.integers(5, 120))].tobytes()
            if i % 37 == 0:
                r = r[:3] + b"N" + r[4:]
            reads.append(r)
        plain = b"\n".join(reads) + b"\n"
        fastq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
        fasta = b"".join(b">s%d desc\n%s\n" % (i, b"\n".join(r[j:j + 60] for j in range(0, len(r), 60))) for i, r in enumerate(reads))
        long_line = b"\n".join([genome.tobytes()] * 3) + b"\n"  # lines far longer than a shard: cuts fall on the few newlines
        for name, img, fmt in (("plain", plain, capi.FMT_PLAIN), ("fastq", fastq, capi.FMT_DETECT), ("fasta", fasta, capi.FMT_DETECT),
                               ("long", long_line, capi.FMT_PLAIN), ("tiny", b"ACGTACGTACGTACGTA\n", capi.FMT_PLAIN), ("empty", b"", capi.FMT_PLAIN)):
            a = np.frombuffer(img, dtype=np.uint8)
            tf = np.zeros(1 << 26, dtype=np.uint64)
            st = capi.CountStats()
            rc = lib.aix_count13_multi(mg, m13._h, a.ctypes.data if a.size else None, a.size, fmt, tf.ctypes.data, C.byref(st))
            assert rc == 0, (name, lib.aix_multi_last_error(mg))
            want, wst = single.count13(m13s, a, fmt)
            assert np.array_equal(tf, want), name
            assert st.as_dict() == wst, (name, st.as_dict(), wst)
        # the reference's own fixture
        g = np.load(os.path.join(golden_dir, "golden13.npz"))
        if "plain_reads" in g.files:
            a = g["plain_reads"]
            tf = np.zeros(1 << 26, dtype=np.uint64)
            st = capi.CountStats()
            assert lib.aix_count13_multi(mg, m13._h, a.ctypes.data, a.size, capi.FMT_PLAIN, tf.ctypes.data, C.byref(st)) == 0
            assert np.array_equal(tf, single.count13(m13s, a, capi.FMT_PLAIN)[0])
        m13s.close()
        single.close()
    finally:
        # ctx0 is borrowed from the aix_multi: never let its wrapper destroy it
        try:
            if m13._h:
                lib.aix_mphf_destroy(ctx0._h, m13._h)
                m13._h = C.c_void_p()
        except NameError:
            pass
        try:
            ctx0._h = C.c_void_p()
        except NameError:
            pass
        lib.aix_multi_destroy(mg)


def test_count_kmers13_tool_uses_every_gpu(pf13, g13, tmp_path):
    """the drop-in binary: same file as the single-GPU run whatever the thread (= GPU) argument"""
    binp = os.path.join(ROOT, "aindex_b200", "bin", "count_kmers13")
    rng = np.random.default_rng(23)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    img = b"".join(rng.choice(acgt, size=int(rng.integers(20, 200))).tobytes() + b"\n" for _ in range(3000))
    inp = tmp_path / "reads.txt"
    inp.write_bytes(img)
    outs = []
    for th in ("1", "16"):
        out = tmp_path / f"o{th}.tf.bin"
        r = subprocess.run([binp, str(inp), pf13, str(out), th], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "GPUs:" in r.stdout
        outs.append(np.fromfile(out, dtype=np.uint64))
    assert np.array_equal(outs[0], outs[1]) and int(outs[0].sum()) > 0
