"""GPU parity tests: the CUDA path (through the C-ABI, aindex_b200.capi) against
  (a) the golden fixtures produced by the unmodified reference (tests/golden/), and
  (b) the CPU oracle (oracle/) on seeded inputs.
Everything is integer work: the bar is bit-exact equality.
"""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def rc(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


@pytest.fixture(scope="module")
def capi():
    from aindex_b200 import capi as c
    return c


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def g23(golden_dir):
    return np.load(os.path.join(golden_dir, "golden23.npz"))


@pytest.fixture(scope="module")
def g13(golden_dir):
    return np.load(os.path.join(golden_dir, "golden13.npz"))


@pytest.fixture(scope="module")
def idx23(capi, ctx, golden_dir):
    return capi.Index23.load_prefix(ctx, os.path.join(golden_dir, "idx23"))


@pytest.fixture(scope="module")
def oidx23(oracle, golden_dir):
    return oracle.Index23.load_prefix(os.path.join(golden_dir, "idx23"))


def _queries(recs, lens):
    return [recs[i, :lens[i]].tobytes() for i in range(recs.shape[0])]


# ---------------------------------------------------------------------------- codec / hash
def test_jenkins_and_codec(oracle, ctx):
    rng = np.random.default_rng(7)
    strs = [rng.integers(0, 256, size=int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 90, size=300)]
    strs += [b"A" * n for n in (0, 1, 7, 8, 9, 15, 16, 17, 22, 23, 24, 25, 47, 48, 49, 71, 72, 73)]
    seed = 0x0123456789ABCDEF
    got = ctx.jenkins64(seed, strs)
    for s, g in zip(strs, got):
        assert tuple(int(x) for x in g) == oracle.jenkins64(seed, s)
    kmers = [rng.choice(ACGT, size=23).tobytes() for _ in range(200)]
    kmers += [b"ACGTNACGTacgtACGT~ACGTAC", b"", b"ACG", b"N" * 23]
    for k, enc, dec, rev in ((23, oracle.dna23_bitset, oracle.bitset_dna23, oracle.reverse_dna23),
                             (13, oracle.dna13_bitset, oracle.bitset_dna13, oracle.reverse_dna13)):
        vals = ctx.encode(kmers, k)
        assert [int(v) for v in vals] == [enc(s) for s in kmers]
        assert [d.tobytes().decode() for d in ctx.decode(vals, k)] == [dec(int(v)) for v in vals]
        assert [int(v) for v in ctx.revcomp(vals, k)] == [rev(int(v)) for v in vals]
    seq = rng.choice(np.frombuffer(b"ACGTNacgt\n~", dtype=np.uint8), size=1001).tobytes()
    assert np.array_equal(ctx.pack_2bit(seq), oracle.dna_bitset_pack(seq))
    for k, enc, rev in ((23, oracle.dna23_bitset, oracle.reverse_dna23), (13, oracle.dna13_bitset, oracle.reverse_dna13)):
        fwd, rcv, valid = ctx.rolling_kmers(seq, k)
        for i in range(0, len(seq) - k + 1, 7):
            w = seq[i:i + k]
            assert int(fwd[i]) == enc(w) and int(rcv[i]) == rev(enc(w))
            assert int(valid[i]) == (1 if not w.strip(b"ACGT") else 0)


@pytest.mark.parametrize("n", [0, 1, 3, 15, 16, 17, 511, 512, 513, 100_003])
def test_pack2bit_vector_kernel_and_ukmer(oracle, ctx, n):
    """dna_bitset ctor (vector kernel over the aligned bulk + generic tail) and dna_bitset::ukmer against the oracle
    (dna_bitseq.hpp:22-61, 124-151), ragged lengths around the 16-byte vector and the 512-byte warp tile"""
    rng = np.random.default_rng(n + 1)
    seq = rng.choice(np.frombuffer(b"ACGTNacgt\n~", dtype=np.uint8), size=n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .01, .01]).tobytes()
    packed = ctx.pack_2bit(seq)
    assert np.array_equal(packed, oracle.dna_bitset_pack(seq))
    if n >= 40:
        for k in (1, 13, 23, 31, 32):
            pos = np.unique(np.concatenate([rng.integers(0, n - k + 1, size=200), [0, n - k]])).astype(np.uint64)
            got = ctx.ukmers(packed, n, pos, k)
            want = np.array([oracle.dna_bitset_ukmer(packed, int(p), k) for p in pos], dtype=np.uint64)
            assert np.array_equal(got, want)
        # past the end of the sequence: missing bases read as A
        got = ctx.ukmers(packed, n, np.array([n - 5, n - 1, n, n + 7], dtype=np.uint64), 23)
        full = oracle.dna_bitset_pack(seq + b"A" * 64)
        assert np.array_equal(got, np.array([oracle.dna_bitset_ukmer(full, p, 23) for p in (n - 5, n - 1, n, n + 7)], dtype=np.uint64))


@pytest.mark.parametrize("k", [13, 23])
def test_rolling_vector_kernel_every_window(oracle, ctx, k, monkeypatch):
    """rolling forward / reverse-complement k-mers (128-bit loads + shuffles + staged stores): EVERY window of a buffer
    that crosses warp tiles (512 windows), with separators and lower case, single call and forced chunks"""
    rng = np.random.default_rng(k)
    n = 20_000 + k
    seq = rng.choice(np.frombuffer(b"ACGTNa\n~", dtype=np.uint8), size=n, p=[.24, .24, .24, .24, .01, .01, .01, .01])
    enc = oracle.dna23_bitset if k == 23 else oracle.dna13_bitset
    rev = oracle.reverse_dna23 if k == 23 else oracle.reverse_dna13
    b = seq.tobytes()
    want_f = np.array([enc(b[i:i + k]) for i in range(n - k + 1)], dtype=np.uint64)
    want_r = np.array([rev(int(v)) for v in want_f], dtype=np.uint64)
    want_v = np.array([0 if b[i:i + k].strip(b"ACGT") else 1 for i in range(n - k + 1)], dtype=np.uint8)
    for chunk in (None, "4096", "528"):
        if chunk:
            monkeypatch.setenv("AIX_ROLLING_CHUNK", chunk)
        fwd, rcv, valid = ctx.rolling_kmers(seq, k)
        assert np.array_equal(fwd, want_f) and np.array_equal(rcv, want_r) and np.array_equal(valid, want_v)
    monkeypatch.delenv("AIX_ROLLING_CHUNK")
    for m in (k - 1, k, k + 1, 15 + k, 16 + k, 512 + k - 1, 512 + k):
        fwd, rcv, valid = ctx.rolling_kmers(seq[:m], k)
        w = max(0, m - k + 1)
        assert np.array_equal(fwd, want_f[:w]) and np.array_equal(rcv, want_r[:w]) and np.array_equal(valid, want_v[:w])


def test_mphf_lookup_golden(capi, ctx, idx23, g23, pf13, oracle):
    recs, lens = g23["recs"], g23["lens"]
    assert np.array_equal(idx23.mphf.lookup(_queries(recs, lens)), g23["hash"])
    m = capi.Mphf.load(ctx, pf13)
    kat = {b"AAAAAAAAAAAAA": 51399613, b"AAAAAAAAAAAAC": 20651245, b"ACGTACGTACGTA": 11618410,
           b"TTTTTTTTTTTTT": 16974388, b"GATTACAGATTAC": 34020858}
    assert m.lookup(list(kat)).tolist() == list(kat.values())
    # the full permutation, uint32 LE, is pinned by md5 and is a bijection
    perm = m.perm13()
    assert hashlib.md5(perm.tobytes()).hexdigest() == "3864cfb768e7d5182b9b88df1b23b575"
    assert np.array_equal(np.sort(perm), np.arange(1 << 26, dtype=np.uint32))
    # .pf round trip is byte-identical
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.pf")
        m.save(p)
        assert hashlib.md5(open(p, "rb").read()).hexdigest() == "5fadfc861de1b04045926a24b32e456a"


# ---------------------------------------------------------------------------- 23-mer queries
def test_tf23_golden_all_modes(capi, idx23, g23):
    q = _queries(g23["recs"], g23["lens"])
    assert idx23.info == {"n": int(g23["n_kmers"][0]), "canonical_only": True}
    assert np.array_equal(idx23.query(q, capi.Q_TF), g23["tf"])
    assert np.array_equal(idx23.query(q, capi.Q_TOTAL), g23["total"])
    assert np.array_equal(idx23.query(q, capi.Q_BOTH), g23["both"])
    assert np.array_equal(idx23.query(q, capi.Q_KID), g23["kid"])
    assert np.array_equal(idx23.query(q, capi.Q_STRAND), g23["strand"])
    # single-query calls agree with the batch (test_aindex_functionality.py:911-913)
    for i in range(0, len(q), 97):
        assert idx23.query([q[i]], capi.Q_TF)[0] == g23["tf"][i]


def _mixed_queries(rng, oidx, n):
    """hits on both strands, misses, N / lower case / junk, odd lengths."""
    from oracle import oracle as O
    out = []
    chk = oidx.checker
    for i in rng.choice(chk.size, size=n, replace=True):
        km = O.bitset_dna23(int(chk[i])).encode()
        r = rng.random()
        if r < 0.3:
            out.append(km)
        elif r < 0.6:
            out.append(rc(km))
        elif r < 0.8:
            out.append(rng.choice(ACGT, size=23).tobytes())
        elif r < 0.9:
            b = bytearray(km if rng.random() < 0.5 else rc(km))
            b[int(rng.integers(0, 23))] = int(rng.choice(np.frombuffer(b"NnacgtX~\n\x00\xff", dtype=np.uint8)))
            out.append(bytes(b))
        else:
            ln = int(rng.integers(0, 60))
            base = (km + rc(km) + km)[:ln]
            out.append(base)
    return out


def test_tf23_vs_oracle_mixed(capi, oracle, idx23, oidx23):
    rng = np.random.default_rng(11)
    q = _mixed_queries(rng, oidx23, 4000)
    recs, lens = oracle.pack_queries(q, stride=64)
    for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH),
                        (capi.Q_PFID, oracle.MODE_PFID), (capi.Q_STRAND, oracle.MODE_STRAND), (capi.Q_KID, oracle.MODE_KID)):
        got = idx23.query(q, mode)
        want = oidx23.batch(recs, lens, omode)
        assert np.array_equal(got, want), f"mode {mode}"


def test_tf23_fixed_stride_fast_path(capi, oracle, idx23, oidx23):
    """uint8[q,23] records: the shared-memory staged kernel, incl. a ragged last CTA."""
    rng = np.random.default_rng(12)
    for nq in (1, 255, 256, 257, 10007):
        q = [x for x in _mixed_queries(rng, oidx23, 2 * nq) if len(x) == 23][:nq]
        recs = np.frombuffer(b"".join(q), dtype=np.uint8).reshape(len(q), 23).copy()
        for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_BOTH, oracle.MODE_BOTH), (capi.Q_PFID, oracle.MODE_PFID),
                            (capi.Q_STRAND, oracle.MODE_STRAND)):
            assert np.array_equal(idx23.query(recs, mode), oidx23.batch(recs, None, omode))
    assert np.array_equal(idx23.get_freq(oidx23.checker[:500]), oidx23.tf[:500])
    # PHASH_MAP::get_freq(uint64_t): stored k-mers, their reverse complements, random values, values with bits above 46
    pick = oidx23.checker[rng.integers(0, oidx23.n, size=4000)]
    rcs = np.array([oracle.reverse_dna23(int(x)) for x in pick[:1000]], dtype=np.uint64)
    rnd = rng.integers(0, 1 << 46, size=2000, dtype=np.uint64)
    high = pick[:500] | (np.uint64(1) << np.uint64(50))
    high_rc = rcs[:500] | (np.uint64(3) << np.uint64(46))
    u = np.concatenate([pick, rcs, rnd, high, high_rc])
    assert np.array_equal(idx23.get_freq(u), np.array([oidx23.get_freq(int(x)) for x in u], dtype=np.uint32))
    assert idx23.query(np.zeros((0, 23), dtype=np.uint8)).size == 0


def test_tf23_non_canonical_index_two_probe_path(capi, oracle, ctx, oidx23):
    """An index that stores some k-mers in their non-canonical orientation (what the reference's
    broken kmer_counter produces) must take the exact forward-then-reverse path."""
    rng = np.random.default_rng(13)
    kmers = oidx23.checker.copy()
    flip = rng.random(kmers.size) < 0.4
    kmers[flip] = np.array([oracle.reverse_dna23(int(x)) for x in kmers[flip]], dtype=np.uint64)
    counts = rng.integers(1, 1000, size=kmers.size).astype(np.uint32)
    m = capi.Mphf.build(ctx, kmers, 23)
    words, ranks = m.arrays()
    om = oracle.Mphf.from_arrays(m.info["n"], m.info["hash_domain"], m.info["seed"], words, ranks)
    ids = om.lookup_batch(np.array([list(oracle.bitset_dna23(int(x)).encode()) for x in kmers], dtype=np.uint8))
    assert np.array_equal(np.sort(ids), np.arange(kmers.size, dtype=np.uint64))  # minimal + perfect
    checker = np.zeros_like(kmers)
    tf = np.zeros_like(counts)
    checker[ids.astype(np.int64)] = kmers
    tf[ids.astype(np.int64)] = counts
    ix = capi.Index23.upload(ctx, m, checker, tf)
    assert ix.info["canonical_only"] is False
    oix = oracle.Index23(om, checker, tf)
    q = _mixed_queries(rng, oix, 3000)
    recs, lens = oracle.pack_queries(q, stride=64)
    for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH),
                        (capi.Q_PFID, oracle.MODE_PFID), (capi.Q_STRAND, oracle.MODE_STRAND), (capi.Q_KID, oracle.MODE_KID)):
        assert np.array_equal(ix.query(q, mode), oix.batch(recs, lens, omode)), f"mode {mode}"
    q23 = [x for x in q if len(x) == 23]
    r23 = np.frombuffer(b"".join(q23), dtype=np.uint8).reshape(len(q23), 23).copy()
    assert np.array_equal(ix.query(r23, capi.Q_TF), oix.batch(r23, None, oracle.MODE_TF))
    # the probe / verify split used by the hash-range-sharded index (dist.ShardedIndex23), here with one "rank"
    assert np.array_equal(_probe_verify(capi, ctx, m, ix, recs, lens), oix.batch(recs, lens, oracle.MODE_TF))


def _probe_verify(capi, ctx, mphf, ix, recs, lens):
    """aix_tf23_probes_dev -> aix_probe23_dev -> first hit of the two probes (what ShardedIndex23.query does
    around its all-to-alls), on one GPU that owns every id."""
    import torch
    lib = capi.lib()
    q, stride = recs.shape
    r_t = torch.from_numpy(np.ascontiguousarray(recs)).cuda()
    l_t = torch.from_numpy(np.ascontiguousarray(lens)).cuda() if lens is not None else None
    probes = torch.empty((2 * q, 2), dtype=torch.int64, device="cuda")
    res = torch.empty(2 * q, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    info = ix.info
    ctx.check(lib.aix_tf23_probes_dev(ctx.handle, mphf._h, info["n"], int(info["canonical_only"]), r_t.data_ptr(), stride,
                                      l_t.data_ptr() if l_t is not None else None, q, probes.data_ptr()))
    ctx.check(lib.aix_probe23_dev(ctx.handle, ix._h, probes.data_ptr(), 2 * q, res.data_ptr()))
    ctx.sync()
    r = res.cpu().numpy().reshape(q, 2)
    first = (r[:, 0] >> 32) != 0
    return (np.where(first, r[:, 0], r[:, 1]) & 0xFFFFFFFF).astype(np.uint32)


def test_probes_bucket_is_a_counting_sort(capi, ctx):
    """aix_probes_bucket_dev: every live probe lands in its owner's bucket with its local id, k-mer and tag."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(83)
    for world, n in ((1, 1000), (3, 70001), (8, 400003), (16, 5)):
        n_total = 1_000_003
        bounds = [(n_total * r) // world for r in range(world)] + [n_total]
        probes = np.empty((n, 2), dtype=np.uint64)
        probes[:, 0] = rng.integers(0, n_total, size=n, dtype=np.uint64)
        probes[:, 1] = rng.integers(0, 1 << 46, size=n, dtype=np.uint64)
        dead = rng.random(n) < 0.3
        probes[dead, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
        p_t = torch.from_numpy(probes.view(np.int64)).cuda()
        counts = torch.empty(world, dtype=torch.int64, device="cuda")
        send = torch.full((n, 2), -7, dtype=torch.int64, device="cuda")
        tag = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.check(capi.lib().aix_probes_bucket_dev(ctx.handle, p_t.data_ptr(), n, (C.c_uint64 * (world + 1))(*bounds), world,
                                                   counts.data_ptr(), send.data_ptr(), tag.data_ptr()))
        ctx.sync()
        cnt = counts.cpu().numpy()
        live = np.flatnonzero(~dead)
        owner = np.searchsorted(np.array(bounds[1:-1], dtype=np.uint64), probes[live, 0], side="right") if world > 1 else np.zeros(live.size, int)
        assert np.array_equal(cnt, np.bincount(owner, minlength=world))
        total = int(cnt.sum())
        s_np, t_np = send.cpu().numpy().view(np.uint64)[:total], tag.cpu().numpy()[:total].astype(np.int64)
        assert np.array_equal(np.sort(t_np), live)                       # every live probe exactly once
        seg = np.repeat(np.arange(world), cnt)                           # bucket of every slot
        assert np.array_equal(s_np[:, 0] + np.array(bounds, dtype=np.uint64)[seg], probes[t_np, 0])
        assert np.array_equal(s_np[:, 1], probes[t_np, 1])
        assert np.array_equal(seg, np.searchsorted(np.array(bounds[1:-1], dtype=np.uint64), probes[t_np, 0], side="right") if world > 1 else seg)


def test_probe_verify_split_canonical_index(capi, oracle, ctx, idx23, oidx23):
    rng = np.random.default_rng(77)
    q = _mixed_queries(rng, oidx23, 5000)
    recs, lens = oracle.pack_queries(q, stride=64)
    assert np.array_equal(_probe_verify(capi, ctx, idx23.mphf, idx23, recs, lens), oidx23.batch(recs, lens, oracle.MODE_TF))
    q23 = [x for x in q if len(x) == 23]
    r23 = np.frombuffer(b"".join(q23), dtype=np.uint8).reshape(len(q23), 23).copy()
    assert np.array_equal(_probe_verify(capi, ctx, idx23.mphf, idx23, r23, None), oidx23.batch(r23, None, oracle.MODE_TF))


# ---------------------------------------------------------------------------- MPHF build
def test_mphf_build_two_keys_is_unbuildable(capi, ctx):
    """n = 2 gives hash_domain 1 (mphf.hpp:27): both hyperedges are (0,1,2) for every seed, so the
    reference's trial loop (mphf.hpp:47-51) never ends; here it is a clean error."""
    with pytest.raises(capi.AixError) as e:
        capi.Mphf.build(ctx, np.array([5, 9], dtype=np.uint64), 23)
    assert e.value.code == -6
    m = capi.Mphf.build(ctx, np.zeros(0, dtype=np.uint64), 23)
    assert m.info["n"] == 0


@pytest.mark.parametrize("n", [1, 3, 17, 1000, 200_000])
def test_mphf_build_is_minimal_perfect(capi, oracle, ctx, n):
    rng = np.random.default_rng(100 + n)
    kmers = np.unique(rng.integers(0, 1 << 46, size=n + n // 8 + 4, dtype=np.uint64))[:n]
    rng.shuffle(kmers)
    m = capi.Mphf.build(ctx, kmers, 23)
    info = m.info
    assert info["n"] == n and info["hash_domain"] == (int(np.ceil(n * 1.23)) + 2) // 3
    words, ranks = m.arrays()
    om = oracle.Mphf.from_arrays(info["n"], info["hash_domain"], info["seed"], words, ranks)
    recs = ctx.decode(kmers, 23)
    ids_gpu = m.lookup(recs)
    ids_cpu = om.lookup_batch(recs, threads=8)
    assert np.array_equal(ids_gpu, ids_cpu)
    assert np.array_equal(np.sort(ids_gpu), np.arange(n, dtype=np.uint64))
    # the block ranks follow ranked_bitpair_vector.hpp:17-31
    nz = np.array([bin((int(w) | (int(w) >> 1)) & 0x5555555555555555).count("1") for w in words[:4096]], dtype=np.uint64)
    nz = np.concatenate([nz, np.zeros((-nz.size) % 16, dtype=np.uint64)])
    blocks = min(ranks.size, nz.size // 16)
    want = np.concatenate([[0], np.cumsum(nz.reshape(-1, 16).sum(axis=1))])[:blocks]
    assert np.array_equal(ranks[:blocks], want.astype(np.uint64))
    assert int(nz.sum()) == n or words.size > 4096


def test_index_build_pipeline_vs_bruteforce(capi, oracle, ctx):
    """reads -> canonical 23-mer table -> MPHF -> checker/tf fill -> queries, against a brute-force
    dictionary (tests/analyze_kmers.py definition) and the oracle on the built index."""
    rng = np.random.default_rng(21)
    genome = rng.choice(ACGT, size=30000).tobytes()
    lines = []
    for i in range(1500):
        ln = int(rng.integers(20, 160))
        st = int(rng.integers(0, len(genome) - ln))
        r = genome[st:st + ln]
        if rng.random() < 0.5:
            r = rc(r)
        if i % 50 == 7:
            r = r[:10] + b"N" + r[11:]
        if i % 31 == 3:
            r = r + b"~" + rc(genome[st:st + 40])
        lines.append(r)
    reads = b"\n".join(lines) + b"\n"
    want = {}
    for line in reads.split(b"\n"):
        for i in range(len(line) - 22):
            km = line[i:i + 23]
            if km.strip(b"ACGT"):
                continue
            c = min(km, rc(km))
            want[c] = want.get(c, 0) + 1
    kmers, counts = ctx.canonical23_count(reads)
    assert kmers.size == len(want)
    assert np.all(kmers[1:] > kmers[:-1])
    dec = ctx.decode(kmers, 23)
    for i in range(0, kmers.size, 37):
        assert want[dec[i].tobytes()] == int(counts[i])
    assert int(counts.sum()) == sum(want.values())
    m = capi.Mphf.build(ctx, kmers, 23)
    checker = np.zeros(kmers.size, dtype=np.uint64)
    tf = np.zeros(kmers.size, dtype=np.uint32)
    ctx.check(capi.lib().aix_index23_fill(ctx.handle, m._h, kmers.ctypes.data, counts.ctypes.data, kmers.size,
                                          checker.ctypes.data, tf.ctypes.data))
    ix = capi.Index23.upload(ctx, m, checker, tf)
    assert ix.info["canonical_only"] is True
    words, ranks = m.arrays()
    oix = oracle.Index23(oracle.Mphf.from_arrays(m.info["n"], m.info["hash_domain"], m.info["seed"], words, ranks), checker, tf)
    q = _mixed_queries(rng, oix, 5000)
    recs, lens = oracle.pack_queries(q, stride=64)
    got = ix.query(q, capi.Q_TF)
    assert np.array_equal(got, oix.batch(recs, lens, oracle.MODE_TF))
    for s, t in zip(q, got):
        if len(s) == 23 and not s.strip(b"ACGT"):
            assert int(t) == want.get(min(s, rc(s)), 0)
    # positions index on the same reads: every bucket full, ascending, pointing at the k-mer
    indices, positions = ix.positions_build(reads)
    oi, op = oix.positions_build(reads)
    assert np.array_equal(indices, oi) and np.array_equal(positions, op)
    assert (positions > 0).all()


# ---------------------------------------------------------------------------- 13-mers
@pytest.fixture(scope="module")
def m13(capi, ctx, pf13):
    return capi.Mphf.load(ctx, pf13)


def test_count13_reference_fixtures(capi, ctx, m13):
    """count_kmers13 on the reference's tests/data files: md5 of the 512 MiB output."""
    data = b"ATCGATCGATCGATCG\nGCTAGCTAGCTAGCTA\nTTTTAAAACCCCGGGG\nNNNNNNNNNNNNNNNN\n"
    tf, st = ctx.count13(m13, data)
    assert st == {"sequences": 4, "windows": 16, "valid": 12, "invalid": 4}
    assert hashlib.md5(tf.tobytes()).hexdigest() == "f02208bc8a20909ddadd8500ecbf608a"
    tf, st = ctx.count13(m13, b"")
    assert st["windows"] == 0 and not tf.any()
    tf, st = ctx.count13(m13, b"ACGTACGTACGT\n")  # 12 characters: skipped before it counts as a sequence
    assert st == {"sequences": 0, "windows": 0, "valid": 0, "invalid": 0}


@pytest.mark.parametrize("name", ["plain", "plain_nonl", "fastq", "fasta", "crlf"])
def test_count13_golden(capi, ctx, m13, g13, name):
    tf, st = ctx.count13(m13, g13[f"{name}_data"])
    assert [st["sequences"], st["windows"], st["valid"], st["invalid"]] == g13[f"{name}_stats"].tolist()
    nz = np.nonzero(tf)[0]
    assert np.array_equal(nz.astype(np.uint32), g13[f"{name}_ids"])
    assert np.array_equal(tf[nz], g13[f"{name}_counts"])
    assert hashlib.md5(tf.tobytes()).hexdigest() == str(g13[f"{name}_md5"])


def _random_reads(rng, n_reads, read_len, alphabet=b"ACGT", n_rate=0.0):
    a = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=(n_reads, read_len + 1))
    if n_rate:
        a[rng.random(a.shape) < n_rate] = ord("N")
    a[:, read_len] = ord("\n")
    return a.reshape(-1)


@pytest.mark.parametrize("variant", ["0", "1", "2", "3"])
def test_count13_vs_oracle_and_chunking(capi, oracle, ctx, m13, variant, monkeypatch):
    """2 MB of reads against the oracle; the same input through the multi-chunk streaming path
    (forced small chunks) and through every atomics variant must give the same histogram."""
    import aindex_b200.capi as c
    rng = np.random.default_rng(31)
    data = _random_reads(rng, 13000, 150, b"ACGTacgt", n_rate=0.002)
    want, wst = oracle.count13_direct(data)
    perm = m13.perm13()
    lib = c.lib()
    monkeypatch.setenv("AIX_COUNT13_CHUNK", "65536")
    monkeypatch.setenv("AIX_COUNT13_VARIANT", variant)
    tf = np.zeros(1 << 26, dtype=np.uint64)
    st = c.CountStats()
    ctx.check(lib.aix_count13_begin(ctx.handle))
    half = (data.size // 2 // 151) * 151
    ctx.check(lib.aix_count13_add(ctx.handle, data[:half].ctypes.data, half, c.FMT_PLAIN))
    ctx.check(lib.aix_count13_add(ctx.handle, data[half:].ctypes.data, data.size - half, c.FMT_PLAIN))
    ctx.check(lib.aix_count13_finish(ctx.handle, m13._h, 0, 1 << 26, tf.ctypes.data, st))
    ctx.check(lib.aix_count13_end(ctx.handle))
    assert st.as_dict() == wst
    assert np.array_equal(tf[perm.astype(np.int64)], want)
    assert int(tf.sum()) == wst["valid"]


def _random_fasta(rng, n_records, alphabet=b"ACGTacgtN"):
    """multi-line records of ragged widths, empty lines, empty records, adjacent headers, CRLF, '>' inside lines"""
    letters = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for i in range(n_records):
        out.append(b">rec%d some description > with a bracket" % i)
        if rng.random() < 0.05:
            continue  # empty record
        for _ in range(int(rng.integers(1, 12))):
            ln = rng.choice(letters, size=int(rng.integers(0, 90))).tobytes()
            if rng.random() < 0.03:
                ln = ln[:5] + b">" + ln[5:]      # '>' not at a line start is an ordinary (invalid) character
            if rng.random() < 0.03:
                ln += b"\r"
            out.append(ln)
    return b"\n".join(out)


@pytest.mark.parametrize("trailing_newline", [True, False])
def test_count13_fasta_device_concatenation(capi, oracle, ctx, m13, trailing_newline):
    """FASTA records are concatenated on the device (header state scan + stream compaction): windows span
    the line breaks of a record and never a header; compared with the oracle reader (count_kmers13.cpp:211-235)."""
    rng = np.random.default_rng(37 + int(trailing_newline))
    for n_rec in (1, 3, 40, 2500):   # the last one spans ~250 4-KiB tiles
        data = _random_fasta(rng, n_rec) + (b"\n" if trailing_newline else b"")
        arr = np.frombuffer(data, dtype=np.uint8)
        want, wst = oracle.count13_direct(arr, oracle.FMT_FASTA)
        tf, st = ctx.count13(m13, arr, capi.FMT_FASTA)
        assert st == wst, (n_rec, st, wst)
        assert np.array_equal(tf[m13.perm13().astype(np.int64)], want)
        tf2, st2 = ctx.count13(m13, arr, capi.FMT_DETECT)
        assert st2 == wst and np.array_equal(tf2, tf)
    # an image already in HBM at an odd address takes the aligned staging copy
    import torch
    buf = torch.zeros(arr.size + 3, dtype=torch.uint8, device="cuda:0")
    buf[3:] = torch.from_numpy(arr.copy()).cuda()
    torch.cuda.synchronize()
    lib = capi.lib()
    ctx.check(lib.aix_count13_begin(ctx.handle))
    ctx.check(lib.aix_count13_add_dev(ctx.handle, buf.data_ptr() + 3, arr.size, capi.FMT_FASTA))
    st3 = capi.CountStats()
    tf3 = np.zeros(1 << 26, dtype=np.uint64)
    ctx.check(lib.aix_count13_finish(ctx.handle, m13._h, 0, 1 << 26, tf3.ctypes.data, st3))
    ctx.check(lib.aix_count13_end(ctx.handle))
    assert st3.as_dict() == wst and np.array_equal(tf3, tf)


@pytest.mark.parametrize("passes_log2", ["0", "2"])
def test_count13_low_complexity_merge(capi, oracle, ctx, m13, monkeypatch, passes_log2):
    """Period-1 / period-2 windows (poly-A, (CA)n ...) are merged across the warp before they reach L2: tracts of
    every phase and length, broken by N / lower case / newlines, next to random sequence, against the oracle."""
    monkeypatch.setenv("AIX_COUNT13_PASSES_LOG2", passes_log2)
    rng = np.random.default_rng(59)
    lines = []
    units = [b"A", b"C", b"G", b"T", b"AC", b"CA", b"AG", b"GT", b"TA", b"AT", b"CG", b"GC", b"TG", b"ct", b"ACG", b"AAC", b"ACAG"]
    for i in range(6000):
        parts = []
        for _ in range(int(rng.integers(1, 5))):
            r = rng.random()
            if r < 0.6:
                u = units[int(rng.integers(0, len(units)))]
                parts.append((u * 200)[int(rng.integers(0, 3)):][:int(rng.integers(1, 180))])
            else:
                parts.append(rng.choice(ACGT, size=int(rng.integers(1, 60))).tobytes())
            if rng.random() < 0.1:
                parts.append(b"N")
        lines.append(b"".join(parts))
    lines += [b"A" * 13, b"A" * 12, b"AC" * 6 + b"A", b"AC" * 6, b"A" * 5000, b"TG" * 3000, b""]
    data = np.frombuffer(b"\n".join(lines) + b"\n", dtype=np.uint8)
    want, wst = oracle.count13_direct(data, oracle.FMT_PLAIN)
    tf, st = ctx.count13(m13, data, capi.FMT_PLAIN)
    assert st == wst
    assert np.array_equal(tf[m13.perm13().astype(np.int64)], want)
    assert int(want.max()) > 10000  # the hot counters really are hot in this input


def test_count13_device_inplace_ragged_tail(capi, oracle, ctx, m13):
    """A resident plain-text image whose length is not a multiple of 16: whole vectors are counted in place, the
    ragged tail through the staged path with lookback -- windows across the seam are counted exactly once."""
    import torch
    rng = np.random.default_rng(53)
    data = _random_reads(rng, 3000, 150, b"ACGT", n_rate=0.001)
    lib = capi.lib()
    for cut in (data.size, data.size - 1, data.size - 7, data.size - 151 - 9, 16 * 1000 + 3, 15, 40):
        part = np.ascontiguousarray(data[:cut])
        want, wst = oracle.count13_direct(part, oracle.FMT_PLAIN)
        buf = torch.from_numpy(part.copy()).cuda()   # exact size: nothing readable is promised past the end
        torch.cuda.synchronize()
        ctx.check(lib.aix_count13_begin(ctx.handle))
        ctx.check(lib.aix_count13_add_dev(ctx.handle, buf.data_ptr(), part.size, capi.FMT_PLAIN))
        st = capi.CountStats()
        tf = np.zeros(1 << 26, dtype=np.uint64)
        ctx.check(lib.aix_count13_finish(ctx.handle, m13._h, 0, 1 << 26, tf.ctypes.data, st))
        ctx.check(lib.aix_count13_end(ctx.handle))
        assert st.as_dict() == wst, (cut, st.as_dict(), wst)
        assert np.array_equal(tf[m13.perm13().astype(np.int64)], want)


def test_tf13_golden(capi, ctx, m13, g13):
    tf = np.zeros(1 << 26, dtype=np.uint64)
    tf[g13["plain_ids"]] = g13["plain_counts"]
    ix = capi.Index13.upload(ctx, m13, tf)
    q = _queries(g13["q_recs"], g13["q_lens"])
    assert np.array_equal(ix.query(q, capi.Q_TF), g13["q_tf"])
    ok = g13["q_ok"]
    qok = [q[i] for i in ok]
    assert np.array_equal(ix.query(qok, capi.Q_TOTAL), g13["q_total"])
    assert np.array_equal(ix.query(qok, capi.Q_BOTH), g13["q_both"])
    # positions index: compute_aindex13 (1 thread) output
    indices, positions = ix.positions_build(g13["plain_data"])
    assert np.array_equal(positions, g13["pos13_positions"])
    assert hashlib.md5(indices.tobytes()).hexdigest() == str(g13["pos13_indices_md5"])


def test_tf13_vs_oracle_odd_queries(capi, oracle, ctx, m13, g13):
    tf = np.zeros(1 << 26, dtype=np.uint64)
    tf[g13["plain_ids"]] = g13["plain_counts"]
    ix = capi.Index13.upload(ctx, m13, tf)
    oix = oracle.Index13(oracle.Mphf.load(oracle.PF13_PATH), tf)
    rng = np.random.default_rng(41)
    q = [rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), size=int(n)).tobytes()
         for n in rng.choice([0, 5, 12, 13, 13, 13, 14, 16], size=1500)]
    q += _queries(g13["q_recs"], g13["q_lens"])
    recs, lens = oracle.pack_queries(q, stride=16)
    for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH)):
        assert np.array_equal(ix.query(q, mode), oix.batch(recs, lens, omode))
    # uint8[q, 13] batches take the TMA-ring kernel: ragged sizes around the 32-query tile, odd bytes included
    for nq in (1, 31, 32, 33, 4095, 4096, 4097, 10007):
        r13 = rng.choice(ACGT, size=(nq, 13))
        hit = rng.random(nq) < 0.5
        pd = g13["plain_data"]
        st = rng.integers(0, pd.size - 13, size=int(hit.sum()))
        r13[hit] = pd[st[:, None] + np.arange(13)[None, :]]  # windows of the counted reads (some span a newline)
        odd = rng.random(nq) < 0.1
        r13[odd, rng.integers(0, 13, size=int(odd.sum()))] = rng.choice(np.frombuffer(b"NnacgtX~\x00\xff", dtype=np.uint8), size=int(odd.sum()))
        for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH)):
            assert np.array_equal(ix.query(r13, mode), oix.batch(r13, None, omode)), f"13-mer fixed mode {mode} nq {nq}"
    # coverage, k = 13 (aindex.py:314-322 over get_tf_value_13mer)
    seq = g13["plain_data"][:3000].tobytes().replace(b"\n", b"")
    for cutoff in (0, 2):
        assert np.array_equal(ix.coverage(seq, cutoff=cutoff), oix.coverage(seq, cutoff))


def test_all_4p13_query_is_perm(capi, ctx, m13):
    """config 1: tf query of all 4^13 13-mers in numeric order == tf re-indexed by perm13."""
    rng = np.random.default_rng(43)
    tf = rng.integers(0, 1 << 40, size=1 << 26, dtype=np.uint64)
    ix = capi.Index13.upload(ctx, m13, tf)
    perm = m13.perm13().astype(np.int64)
    block = 1 << 22
    for start in (0, 17 * block // 4, (1 << 26) - block):
        v = np.arange(start, start + block, dtype=np.uint64)
        recs = ctx.decode(v, 13)
        got = ix.query(recs, capi.Q_TF)
        assert np.array_equal(got, tf[perm[start:start + block]].astype(np.uint32))


# ---------------------------------------------------------------------------- coverage / positions
def test_coverage23_golden(idx23, g23):
    so, co = g23["cov_seq_off"], g23["cov_off"]
    got = idx23.coverage(g23["cov_seq"], so, cutoff=0)
    assert np.array_equal(got, g23["cov_val"])
    for j in range(len(so) - 1):
        one = idx23.coverage(g23["cov_seq"][so[j]:so[j + 1]], cutoff=3)
        want = g23["cov_val"][co[j]:co[j + 1]]
        assert np.array_equal(one, np.where(want >= 3, want, 0))


def test_coverage23_vs_oracle_long(capi, oracle, idx23, oidx23, golden_dir):
    rng = np.random.default_rng(51)
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    seqs, offs = [], [0]
    for _ in range(40):
        st = int(rng.integers(0, reads.size - 2000))
        s = reads[st:st + int(rng.integers(1, 1500))].copy()
        seqs.append(s)
        offs.append(offs[-1] + s.size)
    cat = np.concatenate(seqs)
    got = idx23.coverage(cat, np.array(offs, dtype=np.int64), cutoff=2)
    want = np.concatenate([oidx23.coverage(s, 2) for s in seqs])
    assert np.array_equal(got, want)


def test_positions23_golden(capi, ctx, idx23, g23, golden_dir):
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    indices, positions = idx23.positions_build(reads)
    assert np.array_equal(indices, np.fromfile(os.path.join(golden_dir, "idx23.indices.bin"), dtype=np.uint64))
    assert np.array_equal(positions, np.fromfile(os.path.join(golden_dir, "idx23.index.bin"), dtype=np.uint64))
    pos = capi.Positions(ctx, indices, positions)
    q = _queries(g23["recs"], g23["lens"])
    qi = g23["pos_qidx"]
    offs, vals = pos.query(idx23, [q[i] for i in qi], 23)
    assert np.array_equal(offs, g23["pos_off"]) and np.array_equal(vals, g23["pos_val"])
    # len(get_positions) == tf (test_aindex_functionality.py:376-380); absent / odd queries -> []
    tfs = idx23.query([q[i] for i in qi])
    assert np.array_equal(np.diff(offs).astype(np.uint32), tfs)
    offs, vals = pos.query(idx23, [b"A" * 23, b"ACGT", b"N" * 23], 23)
    assert vals.size == 0


def test_positions23_overflow_and_underflow(capi, oracle, ctx, oidx23, golden_dir):
    """tf larger / smaller than the true occurrence count: zero tail / first-tf-kept."""
    rng = np.random.default_rng(61)
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    tf = oidx23.tf.copy()
    sel = rng.random(tf.size) < 0.3
    tf[sel] = np.maximum(1, tf[sel] // 2)
    sel2 = rng.random(tf.size) < 0.2
    tf[sel2] += 3
    m = capi.Mphf.from_arrays(ctx, oidx23.mphf.n, oidx23.mphf.hash_domain, oidx23.mphf.seed, oidx23.mphf.words,
                              oidx23.mphf.block_ranks)
    ix = capi.Index23.upload(ctx, m, oidx23.checker, tf)
    oix = oracle.Index23(oidx23.mphf, oidx23.checker, tf)
    indices, positions = ix.positions_build(reads)
    oi, op = oix.positions_build(reads)
    assert np.array_equal(indices, oi) and np.array_equal(positions, op)


def test_get_freq_packed6_equals_string_path(capi, oracle, ctx, idx23, oidx23):
    """the 6-byte dna_bitset record form (aix_get_freq23_packed): same answers as the 23-byte string path and as
    PHASH_MAP::get_freq(uint64_t) restated in the oracle; record layout = aix_pack_2bit of the 23 characters"""
    rng = np.random.default_rng(77)
    n = 300_001  # not a multiple of the block size
    km = ctx.decode(oidx23.checker[rng.integers(0, oidx23.n, size=n)], 23)
    km[::3] = rng.choice(ACGT, size=(len(km[::3]), 23))
    flip = rng.random(n) < 0.5
    km[flip] = ctx.decode(ctx.revcomp(ctx.encode(km[flip], 23), 23), 23)
    p6 = capi.pack23(km)
    assert p6.shape == (n, 6)
    for i in (0, 1, 77):
        assert np.array_equal(p6[i], ctx.pack_2bit(km[i]))  # dna_bitset ctor layout
    got = idx23.get_freq_packed(p6)
    assert np.array_equal(got, idx23.query(km))
    u = ctx.encode(km, 23)
    assert np.array_equal(got[:4000], np.array([oidx23.get_freq(int(x)) for x in u[:4000]], dtype=np.uint32))
    assert (got > 0).sum() > n // 2
    assert idx23.get_freq_packed(p6[:0]).size == 0
    # a non-canonical index takes the two-probe order
    sel = np.arange(0, oidx23.n, 2)
    chk = oidx23.checker.copy()
    chk[sel] = ctx.revcomp(chk[sel], 23)
    m = capi.Mphf.from_arrays(ctx, oidx23.mphf.n, oidx23.mphf.hash_domain, oidx23.mphf.seed, oidx23.mphf.words, oidx23.mphf.block_ranks)
    ix2 = capi.Index23.upload(ctx, m, chk, oidx23.tf)
    assert not ix2.info["canonical_only"]
    o2 = oracle.Index23(oidx23.mphf, chk, oidx23.tf)
    assert np.array_equal(ix2.get_freq_packed(p6[:4000]), np.array([o2.get_freq(int(x)) for x in u[:4000]], dtype=np.uint32))
    assert np.array_equal(ix2.get_freq_packed(p6), ix2.get_freq(u))


def test_single_query_mailbox(capi, oracle, ctx, idx23, oidx23, m13, g13, monkeypatch):
    """q = 1 TF calls go through the resident one-thread mailbox kernel (two PCIe traversals, no launch): same answers
    as the batch path for hits on both strands, misses, odd strings and lengths; survives idling out (relaunch), a
    second index taking the mailbox over, batch calls in between, and being switched off."""
    import time
    rng = np.random.default_rng(123)
    q = _mixed_queries(rng, oidx23, 1500)
    want = idx23.query(q)
    got = np.array([int(idx23.query([x])[0]) for x in q], dtype=np.uint32)
    assert np.array_equal(got, want)
    time.sleep(0.02)  # longer than the idle timeout: the next call relaunches the kernel
    assert int(idx23.query([q[0]])[0]) == int(want[0])
    assert np.array_equal(idx23.query(q[:100]), want[:100])            # a batch call in between
    # a 13-mer index takes the mailbox over, then the 23-mer index takes it back
    tf = np.arange(1 << 26, dtype=np.uint64) % np.uint64(1000)
    ix13 = capi.Index13.upload(ctx, m13, tf)
    k13 = [rng.choice(ACGT, size=13).tobytes() for _ in range(300)] + [b"ACGTNACGTACGT", b"ACG", b"acgtacgtacgta", b"A" * 14]
    want13 = ix13.query(k13)
    got13 = np.array([int(ix13.query([x])[0]) for x in k13], dtype=np.uint32)
    assert np.array_equal(got13, want13)
    assert int(idx23.query([q[5]])[0]) == int(want[5])
    ix13.close()
    assert int(idx23.query([q[6]])[0]) == int(want[6])
    t0 = time.perf_counter()
    n = 5000
    for i in range(n):
        idx23.query([q[i % len(q)]])
    rate_on = n / (time.perf_counter() - t0)
    print(f"[single-call path through ctypes] {rate_on:.0f} calls/s")
    # the same calls timed from C (no interpreter in the loop), echo requests first: same answers
    k23 = np.frombuffer(b"".join(x for x in q if len(x) == 23), dtype=np.uint8).reshape(-1, 23)
    lat = idx23.single_call_latency(k23)
    assert np.array_equal(lat["tf"], idx23.query(k23))
    assert 0 < lat["echo_ns"] <= lat["query_ns"] * 1.5
    print(f"[single call from C] echo {lat['echo_ns']:.0f} ns, lookup {lat['query_ns']:.0f} ns")


# ---------------------------------------------------------------------------- size-independent properties
def test_large_batch_properties(capi, ctx, idx23, oidx23):
    """2 M queries through the chunked host pipeline: revcomp invariance, total == 2 x tf,
    batch == concatenation of sub-batches."""
    rng = np.random.default_rng(71)
    n = 2_000_000
    pick = rng.integers(0, oidx23.n, size=n)
    km = ctx.decode(oidx23.checker[pick], 23)
    miss = rng.random(n) < 0.5
    km[miss] = rng.choice(ACGT, size=(int(miss.sum()), 23))
    tf = idx23.query(km)
    rcv = ctx.decode(ctx.revcomp(ctx.encode(km, 23), 23), 23)
    assert np.array_equal(idx23.query(rcv), tf)
    assert np.array_equal(tf[~miss], oidx23.tf[pick[~miss]])
    assert np.array_equal(idx23.query(km, capi.Q_TOTAL), 2 * tf.astype(np.uint64))
    parts = np.concatenate([idx23.query(km[i:i + 333_333]) for i in range(0, n, 333_333)])
    assert np.array_equal(parts, tf)


# ---------------------------------------------------------------------------- device-resident paths
def test_positions_dev_build_and_query_match_host_path(capi, ctx, idx23, g23, golden_dir):
    """aix_positions_build23_dev / aix_positions_query_dev (everything stays in HBM) against the
    golden .indices.bin / .index.bin and the host-buffer query path."""
    import torch
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    buf = torch.full((reads.size + 64,), 10, dtype=torch.uint8, device="cuda:0")
    buf[:reads.size] = torch.from_numpy(reads).cuda()
    torch.cuda.synchronize()
    pos = capi.Positions.build_dev(idx23, buf.data_ptr(), reads.size, 23)
    indices, positions = pos.download()
    assert np.array_equal(indices, np.fromfile(os.path.join(golden_dir, "idx23.indices.bin"), dtype=np.uint64))
    assert np.array_equal(positions, np.fromfile(os.path.join(golden_dir, "idx23.index.bin"), dtype=np.uint64))
    assert pos.info == {"n_indices": indices.size, "n_positions": positions.size}
    q = _queries(g23["recs"], g23["lens"])
    offs, vals = pos.query(idx23, [q[i] for i in g23["pos_qidx"]], 23)
    assert np.array_equal(offs, g23["pos_off"]) and np.array_equal(vals, g23["pos_val"])
    # a leading run of separator-only windows (first_start prologue, hash.cpp:973-988)
    lead = np.frombuffer(b"\n~\n" + b"ACGT\n" * 3, dtype=np.uint8)
    reads2 = np.concatenate([lead, reads])
    buf2 = torch.full((reads2.size + 64,), 10, dtype=torch.uint8, device="cuda:0")
    buf2[:reads2.size] = torch.from_numpy(reads2).cuda()
    torch.cuda.synchronize()
    i2, p2 = capi.Positions.build_dev(idx23, buf2.data_ptr(), reads2.size, 23).download()
    hi, hp = idx23.positions_build(reads2)
    assert np.array_equal(i2, hi) and np.array_equal(p2, hp)
    assert np.array_equal(p2[p2 > 0], positions[positions > 0] + lead.size)


def test_canonical23_multi_pass_equals_single_pass(capi, ctx, monkeypatch):
    """Inputs too large for one sort are counted in passes over k-mer ranges; forcing tiny passes on
    a small input must give the same sorted table."""
    rng = np.random.default_rng(97)
    genome = rng.choice(ACGT, size=40000).tobytes()
    lines = []
    for i in range(3000):
        st = int(rng.integers(0, len(genome) - 150))
        r = genome[st:st + int(rng.integers(10, 150))]
        if rng.random() < 0.5:
            r = rc(r)
        if i % 40 == 3:
            r = r[:7] + b"N" + r[8:]
        lines.append(r)
    reads = b"\n".join(lines) + b"\n"
    k1, c1 = ctx.canonical23_count(reads)
    monkeypatch.setenv("AIX_CANONICAL23_PASS_KEYS", "20000")
    k2, c2 = ctx.canonical23_count(reads)
    monkeypatch.delenv("AIX_CANONICAL23_PASS_KEYS")
    assert k1.size > 30000  # > one pass: at least two ranges were needed
    assert np.array_equal(k1, k2) and np.array_equal(c1, c2)
    assert np.all(k2[1:] > k2[:-1])


@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("fp_bits", [0, 4, 8])
@pytest.mark.parametrize("kernel", [0, 1])
def test_tf23_layout_and_kernel_variants(capi, oracle, ctx, oidx23, monkeypatch, wide, fp_bits, kernel):
    """Every HBM/L2 layout (compact / wide MPHF records, no / 4-bit / 8-bit fingerprint tier) and both
    fixed-stride kernels (TMA-ring streaming kernel, one-CTA-per-256 kernel) give the oracle's
    answers in every mode, including ragged batch sizes around the 32-query tile."""
    monkeypatch.setenv("AIX_MPHF_WIDE", str(wide))
    monkeypatch.setenv("AIX_FP_TIER_BITS", str(fp_bits))
    monkeypatch.setenv("AIX_TF23_KERNEL", str(kernel))
    m = capi.Mphf.from_arrays(ctx, oidx23.mphf.n, oidx23.mphf.hash_domain, oidx23.mphf.seed, oidx23.mphf.words,
                              oidx23.mphf.block_ranks)
    ix = capi.Index23.upload(ctx, m, oidx23.checker, oidx23.tf)
    lay = ix.layout
    assert lay["fp_bits"] == fp_bits and lay["mphf_compact"] == (wide == 0) and lay["records"] != "fused"
    rng = np.random.default_rng(100 + 10 * wide + fp_bits + kernel)
    # raw mphf ids (keys and non-keys) must not depend on the record shape
    q = _mixed_queries(rng, oidx23, 3000)
    recs, lens = oracle.pack_queries(q, stride=64)
    assert np.array_equal(m.lookup(q), oidx23.mphf.lookup_batch(recs, lens))
    for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH),
                        (capi.Q_PFID, oracle.MODE_PFID), (capi.Q_STRAND, oracle.MODE_STRAND), (capi.Q_KID, oracle.MODE_KID)):
        assert np.array_equal(ix.query(q, mode), oidx23.batch(recs, lens, omode)), f"generic mode {mode}"
    for nq in (1, 31, 32, 33, 8191, 8192, 20011):
        qq = [x for x in _mixed_queries(rng, oidx23, 2 * nq) if len(x) == 23][:nq]
        r23 = np.frombuffer(b"".join(qq), dtype=np.uint8).reshape(len(qq), 23).copy()
        for mode, omode in ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH),
                            (capi.Q_PFID, oracle.MODE_PFID), (capi.Q_STRAND, oracle.MODE_STRAND), (capi.Q_KID, oracle.MODE_KID)):
            assert np.array_equal(ix.query(r23, mode), oidx23.batch(r23, None, omode)), f"fixed mode {mode} nq {nq}"
    cov_in = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "idx23.reads"), dtype=np.uint8)[:3000]
    assert np.array_equal(ix.coverage(cov_in), oidx23.coverage(cov_in))
    ix.close()
    m.close()


@pytest.mark.parametrize("bits", ["8", "3", "64"])
def test_tf23_front_filter(capi, oracle, ctx, oidx23, monkeypatch, bits):
    """The Bloom filter in front of the MPHF (batches of absent k-mers: one 8-byte request instead of the lookup): forced
    on, forced off and left to the launcher, the fixed-stride TF path gives the oracle's answers for stored k-mers on
    both strands, absent k-mers, strings with lower-case / N / control bytes, at batch sizes around the tile, the
    queue and the CTA; a sparse filter (3 bits per key: most queries pass) and a wide one (64 bits per key) as well."""
    monkeypatch.setenv("AIX_BLOOM_BITS", bits)
    m = capi.Mphf.from_arrays(ctx, oidx23.mphf.n, oidx23.mphf.hash_domain, oidx23.mphf.seed, oidx23.mphf.words,
                              oidx23.mphf.block_ranks)
    ix = capi.Index23.upload(ctx, m, oidx23.checker, oidx23.tf)
    assert ix.filter_stats["filter_bytes"] == (oidx23.checker.size * int(bits) + 63) // 64 * 8
    rng = np.random.default_rng(77 + int(bits))
    km = ctx.decode(oidx23.checker, 23)
    for nq, p_hit in ((4096, 0.5), (4127, 0.0), (65_536, 0.02), (100_003, 0.9), (300_000, 0.0), (262_144, 1.0)):
        r23 = rng.choice(ACGT, size=(nq, 23))
        hit = rng.random(nq) < p_hit
        pick = rng.integers(0, oidx23.checker.size, size=int(hit.sum()))
        r23[hit] = km[pick]
        flip = hit & (rng.random(nq) < 0.5)
        r23[flip] = ctx.decode(ctx.revcomp(ctx.encode(r23[flip], 23), 23), 23)
        odd = rng.random(nq) < 0.01
        r23[odd, rng.integers(0, 23, size=int(odd.sum()))] = rng.choice(np.frombuffer(b"Nnacgt\n\x00~", dtype=np.uint8), size=int(odd.sum()))
        want = oidx23.batch(r23, None, oracle.MODE_TF)
        for mode, kern in (("on", "1"), ("on", "2"), ("on", "3"), ("off", "1"), ("auto", "1"), ("auto", "3")):
            monkeypatch.setenv("AIX_FILTER_KERNEL", kern)  # one query per lane and iteration / two
            ix.set_filter(mode)
            assert np.array_equal(ix.query(r23), want), f"filter {mode}, kernel {kern}, {nq} queries, hit fraction {p_hit}"
    st = ix.filter_stats
    assert st["batches_filter"] > 0 and st["batches_direct"] > 0 and st["queries_counted"] > 0
    # device buffers: the kernel with two queries per lane needs an 8-byte aligned result array, a 4-byte aligned one falls
    # back to the kernel with one; neither writes outside [0, nq)
    import torch
    ix.set_filter("on")
    for nq, p_hit in ((100_003, 0.05), (8191, 0.5), (4096 + 63, 0.0)):
        r23 = rng.choice(ACGT, size=(nq, 23))
        hit = rng.random(nq) < p_hit
        r23[hit] = km[rng.integers(0, oidx23.checker.size, size=int(hit.sum()))]
        r23[-1] = km[0]   # the last query of the batch is a stored k-mer (the queue's careful load)
        want = oidx23.batch(r23, None, oracle.MODE_TF)
        d = torch.from_numpy(r23).cuda()
        for off, kern in ((0, "1"), (0, "2"), (1, "2"), (0, "3"), (1, "3")):
            monkeypatch.setenv("AIX_FILTER_KERNEL", kern)
            o = torch.full((nq + 3,), -7, dtype=torch.int32, device="cuda")
            ix.query_dev(d.data_ptr(), 23, None, nq, capi.Q_TF, o.data_ptr() + 4 * off)
            ctx.sync()
            got = o.cpu().numpy()
            assert np.array_equal(got[off:off + nq].view(np.uint32), want), f"device buffers, offset {off}, {nq} queries"
            assert (got[:off] == -7).all() and (got[off + nq:] == -7).all()
    # every stored k-mer passes its own filter: no false negatives on either strand
    ix.set_filter("on")
    reps = -(-8192 // km.shape[0])
    allk = np.tile(km, (reps, 1))
    assert np.array_equal(ix.query(allk), np.tile(oidx23.tf, reps))
    assert np.array_equal(ix.query(ctx.decode(ctx.revcomp(ctx.encode(allk, 23), 23), 23)), np.tile(oidx23.tf, reps))
    ix.close()
    m.close()


@pytest.mark.parametrize("kernel", [0, 1])
def test_tf23_fused_layout(capi, oracle, ctx, oidx23, monkeypatch, kernel, golden_dir):
    """The fused layout (16 pair values + 16 x 4-bit fingerprints + rank per 16-byte MPHF record; the default when the
    records fit L2): every query mode, ragged batches, odd strings whose raw bytes are hashed (their probes must skip the
    node's fingerprint), values with bits above 46 through get_freq, coverage, positions build and positions query --
    all equal to the oracle and to the separate-tier layout."""
    monkeypatch.setenv("AIX_TF23_KERNEL", str(kernel))
    monkeypatch.setenv("AIX_INDEX23_LAYOUT", "fused")
    m = capi.Mphf.from_arrays(ctx, oidx23.mphf.n, oidx23.mphf.hash_domain, oidx23.mphf.seed, oidx23.mphf.words,
                              oidx23.mphf.block_ranks)
    ix = capi.Index23.upload(ctx, m, oidx23.checker, oidx23.tf)
    lay = ix.layout
    assert lay["fp_bits"] == 4 and lay["fp_bytes"] == 0 and lay["records"] == "fused"
    rng = np.random.default_rng(300 + kernel)
    q = _mixed_queries(rng, oidx23, 6000)
    recs, lens = oracle.pack_queries(q, stride=64)
    modes = ((capi.Q_TF, oracle.MODE_TF), (capi.Q_TOTAL, oracle.MODE_TOTAL), (capi.Q_BOTH, oracle.MODE_BOTH),
             (capi.Q_PFID, oracle.MODE_PFID), (capi.Q_STRAND, oracle.MODE_STRAND), (capi.Q_KID, oracle.MODE_KID))
    for mode, omode in modes:
        assert np.array_equal(ix.query(q, mode), oidx23.batch(recs, lens, omode)), f"generic mode {mode}"
    for nq in (1, 31, 33, 8192, 50_021):
        qq = [x for x in _mixed_queries(rng, oidx23, 2 * nq) if len(x) == 23][:nq]
        r23 = np.frombuffer(b"".join(qq), dtype=np.uint8).reshape(len(qq), 23).copy()
        for mode, omode in modes:
            assert np.array_equal(ix.query(r23, mode), oidx23.batch(r23, None, omode)), f"fixed mode {mode} nq {nq}"
    # every stored k-mer is found on both strands (no false negative from a node's fingerprint)
    km = ctx.decode(oidx23.checker, 23)
    assert np.array_equal(ix.query(km), oidx23.tf)
    assert np.array_equal(ix.query(ctx.decode(ctx.revcomp(oidx23.checker, 23), 23)), oidx23.tf)
    assert np.array_equal(ix.get_freq(oidx23.checker), oidx23.tf)
    assert np.array_equal(ix.get_freq_packed(capi.pack23(km)), oidx23.tf)
    # stored values with bits above 46 (only get_freq(uint64_t) can name them): own fingerprints, still exact
    chk = oidx23.checker.copy()
    chk[::5] |= np.uint64(1) << np.uint64(50)
    ix2 = capi.Index23.upload(ctx, m, chk, oidx23.tf)
    o2 = oracle.Index23(oidx23.mphf, chk, oidx23.tf)
    probe = np.concatenate([chk[:400], oidx23.checker[:400]])
    assert np.array_equal(ix2.get_freq(probe), np.array([o2.get_freq(int(x)) for x in probe], dtype=np.uint32))
    ix2.close()
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    assert np.array_equal(ix.coverage(reads[:5000]), oidx23.coverage(reads[:5000]))
    noisy = reads[:5000].copy()
    noisy[::97] = ord("N")
    noisy[50::131] = ord("a")
    assert np.array_equal(ix.coverage(noisy), oidx23.coverage(noisy))
    gi, gp = ix.positions_build(reads)
    assert np.array_equal(gi, np.fromfile(os.path.join(golden_dir, "idx23.indices.bin"), dtype=np.uint64))
    assert np.array_equal(gp, np.fromfile(os.path.join(golden_dir, "idx23.index.bin"), dtype=np.uint64))
    ix.close()
    m.close()
