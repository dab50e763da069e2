"""GPU tests of the hand-written radix sort / run-length kernels (csrc/radix_sort.cu) through the C-ABI,
against numpy's stable sort and np.unique, and of the sort-based positions build in its three regimes
(every bucket exactly full / fewer occurrences than tf / more occurrences than tf) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from aindex_b200 import capi as c
    return c


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


def _sort(capi, ctx, keys, b0, b1):
    import torch
    k = torch.from_numpy(keys.view(np.int64).copy()).cuda()
    alt = torch.full_like(k, -1)
    torch.cuda.synchronize()
    in_alt = C.c_int(-1)
    ctx.check(capi.lib().aix_sort_u64_dev(ctx.handle, k.data_ptr(), alt.data_ptr(), keys.size, b0, b1, C.byref(in_alt)))
    ctx.sync()
    return (alt if in_alt.value else k).cpu().numpy().view(np.uint64)


def _want(keys, b0, b1):
    width = b1 - b0
    digit = (keys >> np.uint64(b0)) & np.uint64((1 << width) - 1 if width < 64 else 0xFFFFFFFFFFFFFFFF)
    return keys[np.argsort(digit, kind="stable")]


@pytest.mark.parametrize("n", [0, 1, 2, 31, 100, 8191, 8192, 8193, 70_001, 1_000_003])
@pytest.mark.parametrize("bits", [(0, 64), (0, 47), (33, 61), (5, 12), (20, 29), (63, 64)])
def test_radix_sort_matches_stable_numpy_sort(capi, ctx, n, bits):
    rng = np.random.default_rng(n * 131 + bits[0])
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    got = _sort(capi, ctx, keys, *bits)
    assert np.array_equal(got, _want(keys, *bits))


def test_radix_sort_skewed_and_large(capi, ctx):
    rng = np.random.default_rng(5)
    n = 20_000_000
    # packed (bucket << 33 | position) keys in position order, as the positions build emits them:
    # a few huge buckets (repeats) + many small ones; stability must keep positions ascending per bucket
    bucket = np.where(rng.random(n) < 0.3, rng.integers(0, 4, size=n), rng.integers(0, 1 << 27, size=n)).astype(np.uint64)
    keys = (bucket << np.uint64(33)) | np.arange(1, n + 1, dtype=np.uint64)
    got = _sort(capi, ctx, keys, 33, 33 + 27)
    assert np.array_equal(got, np.sort(keys))  # positions ascending inside a bucket <=> full-key order here
    same = np.full(100_000, 0x123456789ABCDEF0, dtype=np.uint64)
    assert np.array_equal(_sort(capi, ctx, same, 0, 64), same)


@pytest.mark.parametrize("n", [0, 1, 5, 2048, 2049, 500_000])
def test_run_length_matches_numpy_unique(capi, ctx, n):
    import torch
    rng = np.random.default_rng(n + 9)
    keys = np.sort(rng.integers(0, max(1, n // 3), size=n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15 >> 20))
    k = torch.from_numpy(keys.view(np.int64).copy()).cuda()
    uniq = torch.full((max(n, 1),), -1, dtype=torch.int64, device="cuda")
    cnt = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    runs = C.c_uint64(0)
    ctx.check(capi.lib().aix_rle_u64_dev(ctx.handle, k.data_ptr(), n, uniq.data_ptr(), cnt.data_ptr(), C.byref(runs)))
    ctx.sync()
    wu, wc = np.unique(keys, return_counts=True)
    assert runs.value == wu.size
    assert np.array_equal(uniq.cpu().numpy().view(np.uint64)[:wu.size], wu)
    assert np.array_equal(cnt.cpu().numpy().view(np.uint32)[:wu.size], wc.astype(np.uint32))


@pytest.mark.parametrize("regime", ["exact", "tf_larger", "tf_smaller", "tf_zero_somewhere"])
def test_positions_build_regimes_vs_oracle(capi, oracle, ctx, golden_dir, regime):
    """the sort-based build: buckets exactly full (sorted low words = positions), tf above the occurrence count
    (zero tail), tf below it (first tf kept; more keys than sum(tf): the emit pass runs again with room for all)"""
    oidx = oracle.Index23.load_prefix(os.path.join(golden_dir, "idx23"))
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    tf = oidx.tf.copy()
    if regime == "tf_larger":
        tf += 2
    elif regime == "tf_smaller":
        tf = np.maximum(1, tf // 2)
    elif regime == "tf_zero_somewhere":
        tf[::3] = 0
    m = capi.Mphf.from_arrays(ctx, oidx.mphf.n, oidx.mphf.hash_domain, oidx.mphf.seed, oidx.mphf.words, oidx.mphf.block_ranks)
    ix = capi.Index23.upload(ctx, m, oidx.checker, tf)
    oix = oracle.Index23(oidx.mphf, oidx.checker, tf)
    indices, positions = ix.positions_build(reads)
    oi, op = oix.positions_build(reads)
    assert np.array_equal(indices, oi) and np.array_equal(positions, op)
    # ragged inputs: no trailing newline, separators only, shorter than k
    for img in (reads[:-1], np.frombuffer(b"\n~\nNNNN\n", dtype=np.uint8), reads[:10], reads[:0]):
        gi, gp = ix.positions_build(img)
        wi, wp = oix.positions_build(img)
        assert np.array_equal(gi, wi) and np.array_equal(gp, wp)


def test_write_dat_equals_the_oracle_writer(capi, oracle, ctx, golden_dir, tmp_path):
    """aix_canonical23_count + aix_write_dat = the counting stage's text output; byte-equal to the oracle's"""
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    k, c = ctx.canonical23_count(reads)
    ok, oc = oracle.canonical23_count(reads)
    assert np.array_equal(k, ok) and np.array_equal(c, oc)
    c = c.copy()
    c[:3] = [1, 4294967295, 1000000]  # every digit count
    gd, gk, od, okp = (str(tmp_path / n) for n in ("g.dat", "g.kmers", "o.dat", "o.kmers"))
    ctx.write_dat(k, c, gd, gk)
    oracle.write_dat(k, c, od, okp)
    assert open(gd, "rb").read() == open(od, "rb").read() and open(gk, "rb").read() == open(okp, "rb").read()
    ctx.write_dat(k[:0], c[:0], gd, None)
    assert os.path.getsize(gd) == 0


@pytest.mark.parametrize("n", [0, 1, 4095, 4096, 4097, 300_001])
@pytest.mark.parametrize("n_ranges", [1, 2, 3, 8, 16])
def test_partition_by_range_is_stable(capi, ctx, n, n_ranges):
    """the routing pass of the multi-GPU positions build: keys grouped by key range, input order kept inside a range"""
    import torch
    rng = np.random.default_rng(n + n_ranges)
    keys = rng.integers(0, 1 << 40, size=n, dtype=np.uint64)
    bounds = np.sort(rng.integers(0, 1 << 40, size=n_ranges, dtype=np.uint64))
    bounds[0] = 0
    if n_ranges > 2:
        bounds[2] = bounds[1]  # an empty range
    k = torch.from_numpy(keys.view(np.int64).copy()).cuda()
    out = torch.full_like(k, -1)
    counts = np.zeros(n_ranges, dtype=np.uint64)
    torch.cuda.synchronize()
    ctx.check(capi.lib().aix_partition_u64_dev(ctx.handle, k.data_ptr(), out.data_ptr(), n, bounds.ctypes.data, n_ranges, counts.ctypes.data))
    owner = np.searchsorted(bounds, keys, side="right") - 1
    # ranges with equal bounds: the key belongs to the LAST range whose bound is <= key
    want = keys[np.argsort(owner, kind="stable")]
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    assert np.array_equal(counts, np.bincount(owner, minlength=n_ranges).astype(np.uint64))


@pytest.mark.parametrize("ranks", [2, 3, 4])
@pytest.mark.parametrize("regime", ["exact", "tf_smaller", "tf_larger"])
def test_positions_build_multi_equals_single(capi, oracle, golden_dir, ranks, regime):
    """aix_positions_build23_multi (reads sharded by byte range, keys routed to the owner of their bucket range, owner
    sorts): bit-identical to the single-GPU build and to the reference's files.  On a 1-GPU box the contexts share
    device 0 (peer copies become device copies); the sharding, routing, ordering and slicing logic is the same."""
    import ctypes as C
    import torch
    lib = capi.lib()
    n_dev = torch.cuda.device_count()
    ids = list(range(ranks)) if n_dev >= ranks else [0] * ranks
    mg = C.c_void_p()
    assert lib.aix_multi_create(ranks, (C.c_int * ranks)(*ids), C.byref(mg)) == 0
    oidx = oracle.Index23.load_prefix(os.path.join(golden_dir, "idx23"))
    reads = np.fromfile(os.path.join(golden_dir, "idx23.reads"), dtype=np.uint8)
    tf = oidx.tf.copy()
    if regime == "tf_smaller":
        tf = np.maximum(1, tf // 2)
    elif regime == "tf_larger":
        tf[::2] += 3
    ctxs, ms, ixs = [], [], []
    try:
        for r in range(ranks):
            c = capi.Context.__new__(capi.Context)
            c._h = C.c_void_p(lib.aix_multi_ctx(mg, r))
            ctxs.append(c)
            m = capi.Mphf.from_arrays(c, oidx.mphf.n, oidx.mphf.hash_domain, oidx.mphf.seed, oidx.mphf.words, oidx.mphf.block_ranks)
            ms.append(m)
            ixs.append(capi.Index23.upload(c, m, oidx.checker, tf))
        handles = (C.c_void_p * ranks)(*[ix._h for ix in ixs])
        want_i, want_p = ixs[0].positions_build(reads)
        if regime == "exact":
            assert np.array_equal(want_p, np.fromfile(os.path.join(golden_dir, "idx23.index.bin"), dtype=np.uint64))
        for img in (reads, reads[:-1], reads[:5000], reads[:30], np.frombuffer(b"\n~\n" + reads[:4000].tobytes(), dtype=np.uint8)):
            wi, wp = ixs[0].positions_build(img)
            gi = np.zeros(oidx.n + 1, dtype=np.uint64)
            gp = np.full(max(1, wp.size), 0xDEAD, dtype=np.uint64)
            st = capi.MultiBuildStats()
            rc = lib.aix_positions_build23_multi(mg, handles, img.ctypes.data if img.size else None, img.size, gi.ctypes.data, gp.ctypes.data, C.byref(st))
            assert rc == 0, lib.aix_multi_last_error(mg)
            assert np.array_equal(gi, wi) and np.array_equal(gp[:wp.size], wp), (regime, ranks, img.size)
            assert st.positions == wp.size
    finally:
        for ix in ixs:
            ix.close()
        for m in ms:
            m.close()
        for c in ctxs:
            c._h = C.c_void_p()  # borrowed from the aix_multi
        lib.aix_multi_destroy(mg)
