"""Pins the CPU oracle (oracle/aindex_oracle.c) against answers of the UNMODIFIED reference.

Known-answer values: SURVEY.md 8(c) (generated from the compiled reference); fixtures:
tests/golden/*.npz + idx23.* (tests/golden/make_golden.py, reference binaries/module only).
"""
import hashlib
import os

import numpy as np
import pytest

H = lambda *xs: tuple(int(x, 16) for x in xs)


def test_jenkins_kat(oracle):
    # jenkins64_hasher(0x0123456789abcdef)
    seed = 0x0123456789ABCDEF
    assert oracle.jenkins64(seed, "ACGTACGTACGTACGTACGTACG") == H(
        "a255fe88523b0313", "e39eec7d367108e1", "65c05b62c7c537ed")
    assert oracle.jenkins64(seed, "A" * 23) == H(
        "c03f7b71db29087e", "08859acb98c3562b", "abcdc855339a45ae")
    assert oracle.jenkins64(seed, "GATTACAGATTAC") == H(
        "a9bbf3a3e2b9e232", "4279e9d1c59724be", "b311211806d19a79")
    assert oracle.jenkins64(seed, "A") == H(
        "e41d18c97a8aebe2", "fd64f43ff4ab61b1", "398f6067f4847ee7")


def test_codec_kat(oracle, golden_dir):
    assert oracle.dna23_bitset("ACGTACGTACGTACGTACGTACG") == 7450808207046
    assert oracle.reverse_dna23(7450808207046) == 29803232828187
    assert oracle.dna13_bitset("GATTACAGATTAC") == 37505265
    assert oracle.reverse_dna13(37505265) == 46365453
    # test_kmer_conversion.py:73-103 index <-> 13-mer pairs
    for idx, km in [(0, "AAAAAAAAAAAAA"), (1, "AAAAAAAAAAAAC"), (2, "AAAAAAAAAAAAG"),
                    (3, "AAAAAAAAAAAAT"), (4, "AAAAAAAAAAACA"), (4 ** 13 - 1, "TTTTTTTTTTTTT")]:
        assert oracle.bitset_dna13(idx) == km
        assert oracle.dna13_bitset(km) == idx
    g = np.load(os.path.join(golden_dir, "golden_kat.npz"))
    for s, r in zip(g["rc23_in"], g["rc23_out"]):
        assert oracle.bitset_dna23(oracle.reverse_dna23(oracle.dna23_bitset(s))) == r.decode()
    rng = np.random.default_rng(5)
    for _ in range(200):
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=23).tobytes()
        x = oracle.dna23_bitset(s)
        assert oracle.bitset_dna23(x) == s.decode()
        assert oracle.reverse_dna23(oracle.reverse_dna23(x)) == x
        comp = s.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]
        assert oracle.bitset_dna23(oracle.reverse_dna23(x)) == comp.decode()
        p = oracle.dna_bitset_pack(s)
        assert oracle.dna_bitset_ukmer(p, 0, 23) == x
        assert oracle.dna_bitset_ukmer(p, 5, 13) == oracle.dna13_bitset(s[5:18])


def test_pf13_header_and_ids(oracle, pf13):
    assert hashlib.md5(open(pf13, "rb").read()).hexdigest() == "5fadfc861de1b04045926a24b32e456a"
    m = oracle.Mphf.load(pf13)
    assert (m.n, m.hash_domain, m.seed, m.bv_size) == (67108864, 27514635, 0xF9E51456553305F9, 82543905)
    kat = {"AAAAAAAAAAAAA": (H("ffb6ab0fc30c7a19", "f574d0270d84bb41", "5ec00b3e598a102f"), 51399613),
           "AAAAAAAAAAAAC": (H("e9c3e44f04620901", "ef8901b15b49ce2b", "f9ddfb163929477c"), 20651245),
           "ACGTACGTACGTA": (H("2eb82aba9bd93b22", "02264ab39eedb95d", "84746214b2a4aa24"), 11618410),
           "TTTTTTTTTTTTT": (H("2ce9b5285df46e9a", "fd747456b15eae12", "bcbc8cb70eda95bb"), 16974388),
           "GATTACAGATTAC": (H("d0b1c294b30fdf11", "e1d2199467608fa3", "90bf598d5449d78b"), 34020858)}
    for s, (h, idx) in kat.items():
        assert oracle.jenkins64(m.seed, s) == h
        assert m.lookup(s) == idx


def test_count13_reference_fixtures(oracle, pf13):
    """count_kmers13 on the reference's own tests/data (md5s from SURVEY 8(c)); the files are tiny,
    their bytes are restated here so the test does not read /root/reference."""
    m = oracle.Mphf.load(pf13)
    fx = {
        "test_reads.txt": (b"ATCGATCGATCGATCG\nGCTAGCTAGCTAGCTA\nTTTTAAAACCCCGGGG\nNNNNNNNNNNNNNNNN\n",
                           (4, 16, 12, 4), "f02208bc8a20909ddadd8500ecbf608a"),
    }
    ids = {1750752, 12882478, 29390612, 29436569, 32901741, 33917353, 36648121, 36930144,
           43948230, 52843835, 59073752, 59254560}
    for name, (data, st, md5) in fx.items():
        c, s = oracle.count13(m, data)
        assert (s["sequences"], s["windows"], s["valid"], s["invalid"]) == st
        assert hashlib.md5(c.tobytes()).hexdigest() == md5
        assert set(np.nonzero(c)[0].tolist()) == ids and c.max() == 1


def test_count13_golden(oracle, pf13, golden_dir):
    g = np.load(os.path.join(golden_dir, "golden13.npz"))
    m = oracle.Mphf.load(pf13)
    for name in ("plain", "plain_nonl", "fastq", "fasta", "crlf"):
        data = g[f"{name}_data"]
        c, s = oracle.count13(m, data)
        assert [s["sequences"], s["windows"], s["valid"], s["invalid"]] == g[f"{name}_stats"].tolist()
        nz = np.nonzero(c)[0]
        assert np.array_equal(nz.astype(np.uint32), g[f"{name}_ids"])
        assert np.array_equal(c[nz], g[f"{name}_counts"])
        assert hashlib.md5(c.tobytes()).hexdigest() == str(g[f"{name}_md5"])
        # direct-address histogram == MPHF-ordered counts re-indexed (SURVEY 8(c) semantics)
        hst, s2 = oracle.count13_direct(data)
        assert s2 == s
        v = np.nonzero(hst)[0]
        recs = np.concatenate([oracle.all_13mers_block(int(x), 1) for x in v]) if v.size else np.zeros((0, 13), np.uint8)
        pid = m.lookup_batch(recs)
        assert np.array_equal(c[pid], hst[v]) and hst.sum() == c.sum()


def _idx13(oracle, pf13, g):
    m = oracle.Mphf.load(pf13)
    tf = np.zeros(1 << 26, dtype=np.uint64)
    tf[g["plain_ids"]] = g["plain_counts"]
    return oracle.Index13(m, tf)


def test_queries13_golden(oracle, pf13, golden_dir):
    g = np.load(os.path.join(golden_dir, "golden13.npz"))
    ix = _idx13(oracle, pf13, g)
    recs, lens = g["q_recs"], g["q_lens"]
    assert np.array_equal(ix.batch(recs, lens, oracle.MODE_TF), g["q_tf"])
    ok = g["q_ok"]
    assert np.array_equal(ix.batch(recs[ok], lens[ok], oracle.MODE_TOTAL), g["q_total"])
    assert np.array_equal(ix.batch(recs[ok], lens[ok], oracle.MODE_BOTH), g["q_both"])


def test_positions13_golden(oracle, pf13, golden_dir):
    g = np.load(os.path.join(golden_dir, "golden13.npz"))
    ix = _idx13(oracle, pf13, g)
    indices, positions = ix.positions_build(g["plain_data"])
    assert np.array_equal(positions, g["pos13_positions"])
    assert hashlib.md5(indices.tobytes()).hexdigest() == str(g["pos13_indices_md5"])


@pytest.fixture(scope="module")
def idx23(oracle, golden_dir):
    return oracle.Index23.load_prefix(os.path.join(golden_dir, "idx23"))


def test_queries23_golden(oracle, idx23, golden_dir):
    g = np.load(os.path.join(golden_dir, "golden23.npz"))
    recs, lens = g["recs"], g["lens"]
    assert idx23.n == int(g["n_kmers"][0])
    assert np.array_equal(g["tf"], g["tf_single"])
    # queries shorter than 22 chars are undefined in the reference; the golden set has none
    assert lens.min() >= 22
    assert np.array_equal(idx23.batch(recs, lens, oracle.MODE_TF), g["tf"])
    assert np.array_equal(idx23.batch(recs, lens, oracle.MODE_TOTAL), g["total"])
    assert np.array_equal(idx23.batch(recs, lens, oracle.MODE_BOTH), g["both"])
    assert np.array_equal(idx23.batch(recs, lens, oracle.MODE_KID), g["kid"])
    assert np.array_equal(idx23.batch(recs, lens, oracle.MODE_STRAND), g["strand"])
    assert np.array_equal(idx23.mphf.lookup_batch(recs, lens), g["hash"])
    # kid -> kmer info (python_wrapper.cpp:744-755)
    for kid, tf, km, rk in zip(g["info_kid"], g["info_tf"], g["info_kmer"], g["info_rkmer"]):
        u = int(idx23.checker[int(kid)])
        assert oracle.bitset_dna23(u) == km.decode()
        assert oracle.bitset_dna23(oracle.reverse_dna23(u)) == rk.decode()
        assert int(idx23.tf[int(kid)]) == int(tf)


def test_positions23_golden(oracle, idx23, golden_dir):
    g = np.load(os.path.join(golden_dir, "golden23.npz"))
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    indices, positions = idx23.positions_build(reads)
    assert np.array_equal(indices, np.fromfile(os.path.join(golden_dir, "idx23.indices.bin"), dtype=np.uint64))
    assert np.array_equal(positions, np.fromfile(os.path.join(golden_dir, "idx23.index.bin"), dtype=np.uint64))
    recs, lens = g["recs"], g["lens"]
    for j, qi in enumerate(g["pos_qidx"]):
        s = recs[qi, :lens[qi]].tobytes()
        want = g["pos_val"][int(g["pos_off"][j]):int(g["pos_off"][j + 1])]
        got = idx23.positions_query(indices, positions, s)
        assert np.array_equal(got, want)
    # absent k-mer: the reference aborts; defined as empty
    assert idx23.positions_query(indices, positions, b"A" * 23).size == 0 or idx23.query([b"A" * 23])[0] > 0


def test_coverage23_golden(oracle, idx23, golden_dir):
    g = np.load(os.path.join(golden_dir, "golden23.npz"))
    so, co = g["cov_seq_off"], g["cov_off"]
    for j in range(len(so) - 1):
        seq = g["cov_seq"][so[j]:so[j + 1]]
        want = g["cov_val"][co[j]:co[j + 1]]
        assert np.array_equal(idx23.coverage(seq, 0), want)
        for cutoff in (2, 5):
            assert np.array_equal(idx23.coverage(seq, cutoff), np.where(want >= cutoff, want, 0))


def test_get_freq_matches_tf(oracle, idx23):
    rng = np.random.default_rng(1)
    for i in rng.choice(idx23.n, size=50, replace=False):
        u = int(idx23.checker[i])
        assert idx23.get_freq(u) == int(idx23.tf[i])
        assert idx23.get_freq(oracle.reverse_dna23(u)) == int(idx23.tf[i])


def test_oracle_canonical23_table_matches_the_golden_index(oracle, golden_dir, tmp_path):
    """orc_canonical23_count (rolling, threaded) against the table the golden index was built from: the brute-force
    Python definition of tests/analyze_kmers.py (tests/golden/make_golden.py::canonical_counts) -> compute_index,
    i.e. the reference-built .kmers.bin / .tf.bin hold exactly these (k-mer, count) pairs."""
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    checker = np.fromfile(os.path.join(golden_dir, "idx23.kmers.bin"), dtype=np.uint64)
    tf = np.fromfile(os.path.join(golden_dir, "idx23.tf.bin"), dtype=np.uint32)
    order = np.argsort(checker)
    for threads in (1, 3, 8):
        k, c = oracle.canonical23_count(reads, threads=threads)
        assert np.array_equal(k, checker[order]) and np.array_equal(c, tf[order])
    # windows broken by 'N', '~', lower case and line ends; inputs shorter than k
    odd = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTA\nACGTNACGTACGTACGTACGTACGTACGTACG~TTTTTTTTTTTTTTTTTTTTTTTTa\nAC", dtype=np.uint8)
    k, c = oracle.canonical23_count(odd, threads=2)
    want = {}
    b = odd.tobytes()
    L = oracle.lib()
    for i in range(len(b) - 22):
        w = b[i:i + 23]
        if w.strip(b"ACGT"):
            continue
        u = L.orc_dna23_bitset(w, 23)
        m = min(u, L.orc_reverse_dna23(u))
        want[m] = want.get(m, 0) + 1
    assert dict(zip(k.tolist(), c.tolist())) == want
    assert oracle.canonical23_count(b"ACGT")[0].size == 0
    # the text files: "KMER\tCOUNT\n" / "KMER\n" in table order
    dat, keys = str(tmp_path / "t.dat"), str(tmp_path / "t.kmers")
    oracle.write_dat(k, c, dat, keys)
    lines = open(dat, "rb").read().split(b"\n")[:-1]
    assert len(lines) == k.size and open(keys, "rb").read() == b"".join(l.split(b"\t")[0] + b"\n" for l in lines)
    for ln, kv, cv in zip(lines, k, c):
        s, n = ln.split(b"\t")
        assert L.orc_dna23_bitset(s, 23) == int(kv) and int(n) == int(cv) and len(s) == 23
