"""GPU parity at BASELINE.json's full sizes, through size-independent properties (bit-exact integer
work) plus the compiled reference on bounded samples.  The generators are bench.py's, so these are
the very inputs the headline numbers are measured on.

  C2  100 M random + 20 M half-hit 23-mer queries on the 50 M-key index
  C3  one GPU's shard of the 13-mer counting job: 25 M x 150 bp reads (3.45 G windows)
  C4  coverage of 1 M x 10 kb sequences;  C5  positions index over 50 M reads (bench_configs.py)
  C2' the same 50 M-key index built by the UNMODIFIED reference tools from the GPU-counted .dat
"""
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    import torch
    import bench
    from aindex_b200 import capi
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # torch work (data generation, checks) and the library's kernels share ONE stream: the caching
    # allocator hands freed blocks out again in stream order, so a second stream would race with it
    prev = torch.cuda.current_stream(dev)
    torch.cuda.set_stream(stream)
    yield types.SimpleNamespace(torch=torch, bench=bench, capi=capi, dev=dev, ctx=ctx, stream=stream)
    torch.cuda.synchronize()
    torch.cuda.set_stream(prev)
    ctx.close()
    torch.cuda.empty_cache()


def _tmp_root():
    import bench
    return bench._tmp_root()


def _revcomp_records(torch, q):
    comp = torch.zeros(256, device=q.device, dtype=torch.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    return comp[q.long()].flip(1).contiguous()


def test_c2_full_size_properties(env):
    t, capi, ctx = env.torch, env.capi, env.ctx
    reads = env.bench.make_reads(t, env.dev, 50_000_000, 10_000_000, 150, 1, 2)
    mphf, index, checker_t, tf_t, n = env.bench.build_index(t, capi, ctx, reads)
    assert index.info["canonical_only"] is True and 49_000_000 < n < 50_000_001
    # checksum of checksums: querying every stored k-mer returns the tf array, in order
    keys = t.empty((n, 23), device=env.dev, dtype=t.uint8)
    lib = capi.lib()
    # decode on the device through the C-ABI host path would copy 1.1 GB: build ASCII with torch instead
    lut = t.tensor(list(b"ACGT"), device=env.dev, dtype=t.uint8)
    for j in range(23):
        keys[:, j] = lut[(checker_t >> (2 * (22 - j))) & 3]
    out = t.empty(n, device=env.dev, dtype=t.int32)
    index.query_dev(keys.data_ptr(), 23, None, n, capi.Q_TF, out.data_ptr())
    ctx.sync()
    assert t.equal(out, tf_t)
    assert int(tf_t.sum().item()) == 10_000_000 * 128  # every window of every read is counted once
    kid = t.empty(n, device=env.dev, dtype=t.int64)
    index.query_dev(keys.data_ptr(), 23, None, n, capi.Q_PFID, kid.data_ptr())
    ctx.sync()
    assert t.equal(kid, t.arange(n, device=env.dev))  # the MPHF is minimal and perfect on its keys
    del keys, kid, out
    # Q2: half substrings of the reads (either strand), half random
    q2 = t.cat([env.bench.make_hit_queries(t, env.dev, reads, 10_000_000, 4), env.bench.make_queries(t, env.dev, 10_000_000, 5)])
    del reads
    t.cuda.empty_cache()
    o2 = t.empty(q2.shape[0], device=env.dev, dtype=t.int32)
    index.query_dev(q2.data_ptr(), 23, None, q2.shape[0], capi.Q_TF, o2.data_ptr())
    ctx.sync()
    assert bool((o2[:10_000_000] > 0).all().item())          # a window of a read is always in the index
    # Q1: 100 M uniform-random queries; reverse-complement invariance, total == 2 x tf, batch splitting
    q = env.bench.make_queries(t, env.dev, 100_000_000, 3)
    o = t.empty(100_000_000, device=env.dev, dtype=t.int32)
    index.query_dev(q.data_ptr(), 23, None, 100_000_000, capi.Q_TF, o.data_ptr())
    ctx.sync()
    for lo in range(0, 100_000_000, 25_000_000):
        rc = _revcomp_records(t, q[lo:lo + 25_000_000])
        orc = t.empty(25_000_000, device=env.dev, dtype=t.int32)
        index.query_dev(rc.data_ptr(), 23, None, 25_000_000, capi.Q_TF, orc.data_ptr())
        ctx.sync()
        assert t.equal(orc, o[lo:lo + 25_000_000])
        del rc, orc
    tot = t.empty(20_000_000, device=env.dev, dtype=t.int64)
    index.query_dev(q2.data_ptr(), 23, None, 20_000_000, capi.Q_TOTAL, tot.data_ptr())
    ctx.sync()
    assert t.equal(tot, 2 * o2.to(t.int64))
    # an odd split point (not a multiple of the 32-query tile, unaligned sub-batch start) gives the same answers
    cut = 33_333_331
    o_b = t.empty(100_000_000 - cut, device=env.dev, dtype=t.int32)
    index.query_dev(q.data_ptr() + cut * 23, 23, None, 100_000_000 - cut, capi.Q_TF, o_b.data_ptr())
    ctx.sync()
    assert t.equal(o_b, o[cut:])
    # the host-buffer C-ABI call and the unmodified reference on a sample
    qs = q[:3_000_000].cpu().numpy()
    assert np.array_equal(index.query(qs), o[:3_000_000].cpu().numpy().view(np.uint32))
    import tempfile, shutil
    tmp = tempfile.mkdtemp(prefix="aix_t_", dir=_tmp_root())
    try:
        prefix = env.bench.write_index_files(tmp, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
        mix = np.concatenate([qs[:1_000_000], q2[:1_000_000].cpu().numpy(), q2[-1_000_000:].cpu().numpy()])
        kind, secs, res = env.bench.cpu_query_runs(prefix, mix, os.cpu_count() or 1, 1)
        assert np.array_equal(res, index.query(mix)), f"differs from the CPU {kind}"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_c3_full_size_shard_properties(env):
    t, capi, ctx = env.torch, env.capi, env.ctx
    lib = capi.lib()
    n_reads = 25_000_000
    reads = env.bench.make_reads(t, env.dev, 100_000_000, n_reads, 150, 11, 12)
    flat = reads.reshape(-1)

    def count(lo, hi):
        ctx.check(lib.aix_count13_begin(ctx.handle))
        ctx.check(lib.aix_count13_add_dev(ctx.handle, flat.data_ptr() + lo * 151, (hi - lo) * 151, capi.FMT_PLAIN))
        ctx.check(lib.aix_count13_flush(ctx.handle))
        st = capi.CountStats()
        ctx.check(lib.aix_count13_stats(ctx.handle, st))
        h = env.bench._wrap_device_i64(t, lib.aix_count13_hist_dev(ctx.handle), 1 << 26, env.dev).clone()
        return h, st.as_dict()

    whole, st = count(0, n_reads)
    assert st == {"sequences": n_reads, "windows": n_reads * 138, "valid": n_reads * 138, "invalid": 0}
    assert int(whole.sum().item()) == n_reads * 138
    # linearity: the histogram of the shard is the sum of the histograms of its parts (any cut at a line start)
    cut = 9_999_937
    a, sa = count(0, cut)
    b, sb = count(cut, n_reads)
    assert t.equal(a + b, whole) and sa["valid"] + sb["valid"] == st["valid"]
    # reverse-complementing every read reverse-complements the histogram: hist'[rc(v)] == hist[v]
    v = t.arange(1 << 26, device=env.dev, dtype=t.int64)
    rcv = t.zeros_like(v)
    for j in range(13):
        rcv |= (3 - ((v >> (2 * j)) & 3)) << (2 * (12 - j))
    sub = reads[:2_000_000]
    comp = t.zeros(256, device=env.dev, dtype=t.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        comp[x] = y
    rsub = sub.clone()
    rsub[:, :150] = comp[sub[:, :150].long()].flip(1)
    ctx.check(lib.aix_count13_begin(ctx.handle))
    ctx.check(lib.aix_count13_add_dev(ctx.handle, rsub.data_ptr(), rsub.numel(), capi.FMT_PLAIN))
    ctx.check(lib.aix_count13_flush(ctx.handle))
    hr = env.bench._wrap_device_i64(t, lib.aix_count13_hist_dev(ctx.handle), 1 << 26, env.dev).clone()
    hs, _ = count(0, 2_000_000)
    assert t.equal(hr[rcv], hs)
    ctx.check(lib.aix_count13_end(ctx.handle))
    # the reference binary on a 300 k-read sample (MPHF-order tf.bin, 512 MiB): byte-equal
    pf = os.path.join(ROOT, "oracle", "_ref", "data", "all_13mers.pf")
    if os.path.exists(pf) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bin", "count_kmers13")):
        import tempfile, shutil
        tmp = tempfile.mkdtemp(prefix="aix_t_", dir=_tmp_root())
        try:
            sample = reads[:300_000].cpu().numpy()
            kind, secs, nk, tf_ref, _ = env.bench.cpu_count_run(sample, os.cpu_count() or 1, tmp)
            m13 = capi.Mphf.load(ctx, pf)
            tf_gpu, stats = ctx.count13(m13, sample.reshape(-1), capi.FMT_PLAIN)
            assert np.array_equal(tf_ref, tf_gpu) and stats["valid"] == nk
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


def test_c2_reference_built_index(env):
    """north_star: "MPHF construction reuses the reference's emphf output, so index files are interchangeable".
    The C2 table counted on the GPU is written as the reference's text files (aix_write_dat), the UNMODIFIED
    compute_mphf_seq + compute_index build {pf, kmers.bin, tf.bin} from them (cached in /dev/shm, ~2-3 min once),
    the files are loaded with aix_index23_load_prefix, and 100 M Q1 + 20 M Q2 answers must equal the GPU-built
    index's; 3 M of them are compared with the reference's own get_freq on the reference-built files."""
    import json
    import shutil
    import time
    from oracle import oracle as O
    t, capi, ctx, bench = env.torch, env.capi, env.ctx, env.bench
    binp = os.path.join(ROOT, "oracle", "_ref", "bin")
    if not (os.path.exists(os.path.join(binp, "compute_mphf_seq")) and os.path.exists(os.path.join(binp, "compute_index"))):
        pytest.skip("reference tools not compiled under oracle/_ref")
    reads = bench.make_reads(t, env.dev, 50_000_000, 10_000_000, 150, 1, 2)
    mphf, index, checker_t, tf_t, n = bench.build_index(t, capi, ctx, reads)
    cache = os.path.join(_tmp_root() or "/tmp", "aix_test_cache_c2_gpucounted")
    prefix = os.path.join(cache, "c2.23")
    if not os.path.exists(os.path.join(cache, "meta.json")):
        os.makedirs(cache, exist_ok=True)
        lib = capi.lib()
        kmers, counts = ctx.canonical23_count(reads.cpu().numpy().reshape(-1))
        # the GPU table equals the CPU definition (tests/analyze_kmers.py) on a 200 k-read slice, and its totals at full size
        ok, oc = O.canonical23_count(reads[:200_000].cpu().numpy().reshape(-1))
        gk, gc = ctx.canonical23_count(reads[:200_000].cpu().numpy().reshape(-1))
        assert np.array_equal(ok, gk) and np.array_equal(oc, gc)
        assert kmers.size == n and int(counts.sum()) == 10_000_000 * 128 and bool(np.all(kmers[1:] > kmers[:-1]))
        t0 = time.perf_counter()
        ctx.write_dat(kmers, counts, prefix + ".dat", prefix + ".kmers")
        t1 = time.perf_counter()
        import subprocess
        subprocess.check_call([os.path.join(binp, "compute_mphf_seq"), prefix + ".kmers", prefix + ".pf"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t2 = time.perf_counter()
        subprocess.check_call([os.path.join(binp, "compute_index"), prefix + ".dat", prefix + ".pf", prefix, str(os.cpu_count() or 1), "0"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t3 = time.perf_counter()
        os.unlink(prefix + ".dat")
        os.unlink(prefix + ".kmers")
        with open(os.path.join(cache, "meta.json"), "w") as f:
            json.dump({"index_keys": int(kmers.size), "write_dat_s": t1 - t0, "compute_mphf_seq_s": t2 - t1, "compute_index_s": t3 - t2}, f)
        del kmers, counts
    meta = json.load(open(os.path.join(cache, "meta.json")))
    assert meta["index_keys"] == n
    ref_ix = capi.Index23.load_prefix(ctx, prefix)
    assert ref_ix.info["canonical_only"] is True and ref_ix.info["n"] == n
    # the reference's tf / checker arrays are the GPU fill's, re-indexed by the other MPHF: same multiset of (kmer, tf)
    rk = np.fromfile(prefix + ".kmers.bin", dtype=np.uint64)
    rt = np.fromfile(prefix + ".tf.bin", dtype=np.uint32)
    o1, o2 = np.argsort(rk), np.argsort(checker_t.cpu().numpy().view(np.uint64))
    assert np.array_equal(rk[o1], checker_t.cpu().numpy().view(np.uint64)[o2]) and np.array_equal(rt[o1], tf_t.cpu().numpy().view(np.uint32)[o2])
    del rk, rt, o1, o2
    q2 = t.cat([bench.make_hit_queries(t, env.dev, reads, 10_000_000, 4), bench.make_queries(t, env.dev, 10_000_000, 5)])
    del reads
    t.cuda.empty_cache()
    q1 = bench.make_queries(t, env.dev, 100_000_000, 3)
    rates = {}
    for name, q in (("q1", q1), ("q2", q2)):
        nq = q.shape[0]
        a, b = t.empty(nq, device=env.dev, dtype=t.int32), t.empty(nq, device=env.dev, dtype=t.int32)
        index.query_dev(q.data_ptr(), 23, None, nq, capi.Q_TF, a.data_ptr())
        ref_ix.query_dev(q.data_ptr(), 23, None, nq, capi.Q_TF, b.data_ptr())
        ctx.sync()
        assert t.equal(a, b), f"{name}: reference-built index answers differ from the GPU-built index's"
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        for _ in range(5):
            ref_ix.query_dev(q.data_ptr(), 23, None, nq, capi.Q_TF, b.data_ptr())
        e1.record(env.stream)
        ctx.sync()
        rates[name] = nq / (e0.elapsed_time(e1) / 5 / 1e3)
        if name == "q2":
            assert bool((b[:10_000_000] > 0).all().item())
    print(f"[reference-built C2 index] {n} keys, Q1 {rates['q1'] / 1e9:.1f} G q/s, Q2 {rates['q2'] / 1e9:.1f} G q/s; "
          f"compute_mphf_seq {meta['compute_mphf_seq_s']:.0f} s, compute_index {meta['compute_index_s']:.0f} s")
    mix = np.concatenate([q1[:1_000_000].cpu().numpy(), q2[:1_000_000].cpu().numpy(), q2[-1_000_000:].cpu().numpy()])
    kind, secs, res = bench.cpu_query_runs(prefix, mix, os.cpu_count() or 1, 1)
    assert kind == "reference" and np.array_equal(res, ref_ix.query(mix))


@pytest.mark.parametrize("config", ["c4", "c5"])
def test_c4_c5_full_size(env, config):
    """bench_configs.py at scale 1.0: every full-size property and every comparison with the
    oracle / the compiled reference must hold (the timings it also takes are not asserted)."""
    import bench_configs
    env.torch.cuda.empty_cache()
    args = types.SimpleNamespace(scale=1.0, checks=True, e2e=True, cpu=True)
    line = {"c4": bench_configs.run_c4, "c5": bench_configs.run_c5}[config](env.ctx, env.stream, env.dev, args)
    assert all(line["checks"].values()), line["checks"]
    cb = line["cpu_baseline"]
    if cb is not None:
        assert all(v for k, v in cb.items() if k.endswith("equal") or k.startswith("results_equal") or "_equal" in k), cb
    if line.get("e2e"):
        assert line["e2e"]["matches_device_path"]
    env.torch.cuda.empty_cache()
