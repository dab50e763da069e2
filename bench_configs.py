#!/usr/bin/env python3
"""Measurement of the BASELINE.json configs beside bench.py's headline; bench.py puts their
results under extra.c1_tf13_all / extra.c4_coverage / extra.c5_positions of its JSON line, and
tests/test_gpu_fullsize.py asserts their full-size property checks.

  C1  tf query of all 4^13 13-mers in numeric order (25 B / lookup)
  C4  sequence coverage, 1 M x 10 kb sequences on the C2 index (17 B / position)
  C5  positions index build over 50 M x 150 bp reads + 10 M position queries (29 B / occurrence)

Every config runs at BASELINE size on one B200: `value` with all buffers resident in HBM (CUDA events on the
library stream), `e2e` through the host-buffer C-ABI call (pinned host memory, copies inside the timed region),
`roofline` against the measured HBM copy peak, `cpu_baseline` = the compiled reference (oracle/_ref) on a bounded
sample of the same workload with its results compared to the GPU's.

  python bench_configs.py --configs c1,c4,c5 [--scale 1.0] [--out gpurun_out/configs.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
import types

import numpy as np

import bench_common as bc

ROOT = bc.ROOT
REF_BIN = bc.REF_BIN


def timed(torch, ctx, stream, fn, reps=3, warmup=1):
    for _ in range(warmup):
        fn()
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    ctx.sync()
    return a.elapsed_time(b) / reps


GATHER_PEAK_G = 50.1  # random 16-byte gathers/s from a 1 GiB table on this GPU (profiles/r02_atomic_roofline.txt): one DRAM burst each


def roof(units, bytes_per_unit, ms, kernel, traffic_key=None, random_access=False):
    """`frac` is the contract's figure (algorithmic bytes against the HBM copy peak).  With a committed ncu capture,
    `traffic_frac` = the DRAM bytes the kernel really moves per second against the same peak, and for kernels whose reads are
    scattered records (`random_access`) `random_access.frac` = 64-byte DRAM bursts per second against the measured random-gather
    rate -- the ceiling such a kernel can actually reach."""
    peak, src = bc.peak_hbm_gbs()
    ach = units * bytes_per_unit / (ms / 1e3) / 1e9
    r = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
         "peak_source": src, "bytes_per_unit": bytes_per_unit, "units_per_launch": int(units), "kernel_ms": ms,
         "traffic": bc.ncu_traffic(traffic_key, units) if traffic_key else None}
    if r["traffic"]:
        r["traffic_gbs"] = r["traffic"] / (ms / 1e3) / 1e9
        r["traffic_frac"] = r["traffic_gbs"] / peak
        if random_access:
            bursts = r["traffic"] / 64.0 / (ms / 1e3) / 1e9
            r["random_access"] = {"dram_bursts_64B_g_per_s": bursts, "peak_ggathers_s": GATHER_PEAK_G, "frac": bursts / GATHER_PEAK_G,
                                  "note": "all DRAM traffic counted as 64-byte bursts (an upper bound of the scattered share)"}
    return r


def _ns(args, **defaults):
    d = dict(scale=1.0, checks=True, e2e=True, cpu=True)
    d.update(defaults)
    d.update({k: v for k, v in vars(args).items() if v is not None} if args is not None else {})
    return types.SimpleNamespace(**d)


def _env(dev):
    import torch
    from aindex_b200 import capi
    return torch, capi, capi.lib()


# ------------------------------------------------------------------------------------------ C1
def run_c1(ctx, stream, dev, args=None):
    """count13 over synthetic reads -> 13-mer index -> tf query of all 4^13 13-mers (python_wrapper.cpp:482-503, 938-980)."""
    args = _ns(args)
    torch, capi, lib = _env(dev)
    n_reads = max(1000, int(1_000_000 * args.scale))
    reads = bc.make_reads(torch, dev, 5_000_000, n_reads, 150, 41, 42)
    if os.path.exists(bc.PF13):
        m13, pf_kind = capi.Mphf.load(ctx, bc.PF13), "reference all_13mers.pf"
    else:
        allk = torch.arange(1 << 26, device=dev, dtype=torch.int64)
        m13, pf_kind = capi.Mphf.build_dev(ctx, allk.data_ptr(), 1 << 26, 13), "GPU-built MPHF"
        del allk
    tf, stats = ctx.count13(m13, reads.cpu().numpy().reshape(-1), capi.FMT_PLAIN)
    del reads
    ix = capi.Index13.upload(ctx, m13, tf)
    q = 1 << 26
    v = torch.arange(q, device=dev, dtype=torch.int64)
    lut = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    recs = torch.empty((q, 13), device=dev, dtype=torch.uint8)
    for j in range(13):
        recs[:, j] = lut[(v >> (2 * (12 - j))) & 3]
    del v
    out = torch.empty(q, device=dev, dtype=torch.int32)

    def step():
        ctx.check(lib.aix_tf13_batch_dev(ctx.handle, ix._h, recs.data_ptr(), 13, None, q, capi.Q_TF, out.data_ptr()))

    ms = timed(torch, ctx, stream, step, reps=10, warmup=3)
    got = out.cpu().numpy().view(np.uint32)
    checks = {}
    if args.checks:
        perm = m13.perm13()
        checks["query_equals_tf_of_perm13"] = bool(np.array_equal(got.astype(np.uint64), tf[perm]))
        checks["sum_equals_valid_windows"] = int(got.sum()) == int(stats["valid"])
    e2e = None
    if args.e2e:
        r_host = ctx.pinned((q, 13), np.uint8)
        o_host = ctx.pinned((q,), np.uint32)
        torch.from_numpy(r_host).copy_(recs)
        torch.cuda.synchronize()

        def host_step():
            ctx.check(lib.aix_tf13_batch(ctx.handle, ix._h, r_host.ctypes.data, 13, None, q, capi.Q_TF, o_host.ctypes.data))

        host_step()
        t0 = time.perf_counter()
        for _ in range(3):
            host_step()
        e_s = (time.perf_counter() - t0) / 3
        e2e = {"value": q / e_s, "unit": "lookups/s", "ms_per_step": e_s * 1e3, "h2d_bytes_per_step": q * 13,
               "d2h_bytes_per_step": q * 4, "api": "aix_tf13_batch (pinned host buffers)",
               "matches_device_path": bool(np.array_equal(o_host, got))}
        del r_host, o_host
    cpu = None
    if args.cpu:
        cpu = _cpu_tf13(tf, got, pf_kind)
    del recs, out
    return {"config": "C1", "workload": f"count13 over {n_reads} reads -> tf query of all 4^13 13-mers in numeric order ({pf_kind})",
            "metric": "13-mer tf lookups/s", "value": q / (ms / 1e3), "unit": "lookups/s", "ms_per_step": ms,
            "roofline": roof(q, 25, ms, "tf13_stream_kernel<AIX_Q_TF>", "tf13_stream_kernel_all_4p13"), "e2e": e2e,
            "checks": checks, "cpu_baseline": cpu}


def _cpu_tf13(tf, got, pf_kind):
    threads = os.cpu_count() or 1
    h = bc.ref_harness_path()
    count = min(1 << 26, 2_000_000 * threads)
    if h and os.path.exists(bc.PF13) and pf_kind.startswith("reference"):
        tmp = tempfile.mkdtemp(prefix="aix_c1_", dir=bc._tmp_root(2 << 30))
        try:
            tfp, op = os.path.join(tmp, "c1.tf.bin"), os.path.join(tmp, "c1.out.bin")
            tf.tofile(tfp)
            r = subprocess.run([h, "tf13", bc.PF13, tfp, str(threads), op, str(count), "2"], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True)
            if r.returncode == 0:
                secs = [float(l.split()[0].split("=")[1]) for l in r.stdout.splitlines() if l.startswith("seconds=")]
                res = np.fromfile(op, dtype=np.uint32)
                return {"value": count / min(secs), "unit": "lookups/s", "cores": threads, "kind": "reference",
                        "sample": f"first {count} of the 4^13 13-mers, get_tf_value_13mer (validity + HASHER::lookup + tf gather) "
                                  f"from {threads} std::threads, best of {len(secs)}",
                        "seconds": min(secs), "results_equal_gpu": bool(np.array_equal(res, got[:count]))}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if not os.path.exists(bc.PF13):
        return None
    from oracle import oracle as O
    oix = O.Index13(O.Mphf.load(bc.PF13), tf)
    sample = O.all_13mers_block(0, count)
    t0 = time.perf_counter()
    ores = oix.batch(sample, None, O.MODE_TF, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": count / dt, "unit": "lookups/s", "cores": threads, "kind": "port",
            "sample": f"first {count} of the 4^13 13-mers, oracle get_tf_value_13mer, {threads} threads",
            "seconds": dt, "results_equal_gpu": bool(np.array_equal(ores, got[:count]))}


# ------------------------------------------------------------------------------------------ C4
def make_sequences(torch, dev, genome_codes, n_seq, seq_len, seed, sub_rate=0.01):
    """n_seq genome substrings (uniform start, random strand) with 1 % substitutions."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    out = torch.empty((n_seq, seq_len), device=dev, dtype=torch.uint8)
    ar = torch.arange(seq_len, device=dev, dtype=torch.int64)
    chunk = 20_000
    for s in range(0, n_seq, chunk):
        e = min(n_seq, s + chunk)
        start = torch.randint(0, genome_codes.numel() - seq_len, (e - s,), generator=g, device=dev, dtype=torch.int64)
        codes = genome_codes[start[:, None] + ar[None, :]]
        flip = torch.rand((e - s,), generator=g, device=dev) < 0.5
        codes = torch.where(flip[:, None], (3 - codes).flip(1), codes)
        sub = torch.rand((e - s, seq_len), generator=g, device=dev) < sub_rate
        shift = torch.randint(1, 4, (e - s, seq_len), generator=g, device=dev, dtype=torch.uint8)
        codes = torch.where(sub, (codes + shift) & 3, codes)
        out[s:e] = lut[codes.long()]
    return out


def run_c4(ctx, stream, dev, args=None, index_bundle=None):
    """aindex.py:314-322 for 1 M x 10 kb sequences on the C2 index.  index_bundle = (mphf, index, checker_t, tf_t,
    n_keys) of the C2 index when the caller already built it (bench.py), else it is built here."""
    args = _ns(args)
    torch, capi, lib = _env(dev)
    n_reads, genome_len = int(10_000_000 * args.scale), int(50_000_000 * args.scale)
    if index_bundle is None:
        reads = bc.make_reads(torch, dev, genome_len, n_reads, 150, 1, 2)
        index_bundle = bc.build_index(torch, capi, ctx, reads)
        del reads
    mphf, index, checker_t, tf_t, n_keys = index_bundle
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    genome = torch.randint(0, 4, (genome_len,), generator=g, device=dev, dtype=torch.uint8)  # same stream as make_reads
    n_seq, seq_len = int(1_000_000 * args.scale), 10_000
    seqs = make_sequences(torch, dev, genome, n_seq, seq_len, 21)
    del genome
    torch.cuda.empty_cache()
    offs = torch.arange(n_seq + 1, device=dev, dtype=torch.int64) * seq_len
    per = seq_len - 22
    total_out = n_seq * per
    out = torch.empty(total_out, device=dev, dtype=torch.int32)

    def step():
        ctx.check(lib.aix_coverage_dev(ctx.handle, index._h, None, seqs.data_ptr(), offs.data_ptr(), n_seq,
                                       seqs.numel(), total_out, 23, 0, out.data_ptr()))

    ms = timed(torch, ctx, stream, step, reps=3, warmup=1)
    hit = float((out[: 50_000_000] > 0).float().mean().item())
    checks = {}
    if args.checks:
        # property at full size: coverage[s, i] == batch tf query of the window seq[s, i:i+23] (random sample)
        g.manual_seed(5)
        ns = 5_000_000
        si = torch.randint(0, n_seq, (ns,), generator=g, device=dev, dtype=torch.int64)
        oi = torch.randint(0, per, (ns,), generator=g, device=dev, dtype=torch.int64)
        win = seqs.reshape(-1)[(si * seq_len + oi)[:, None] + torch.arange(23, device=dev)[None, :]].contiguous()
        qout = torch.empty(ns, device=dev, dtype=torch.int32)
        index.query_dev(win.data_ptr(), 23, None, ns, capi.Q_TF, qout.data_ptr())
        ctx.sync()
        checks["coverage_equals_batch_tf_on_5M_sampled_windows"] = bool(torch.equal(qout, out[si * per + oi]))
        del win, qout, si, oi
    cpu = None
    if args.cpu or args.checks:
        tmpdir = tempfile.mkdtemp(prefix="aix_c4_", dir=bc._tmp_root())
        try:
            prefix = bc.write_index_files(tmpdir, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
            if args.checks:
                # the reference loop aindex.py:314-322 restated in C (oracle) on whole sequences
                from oracle import oracle as O
                oix = O.Index23.load_prefix(prefix)
                n_o = 20
                sh = seqs[:n_o].cpu().numpy()
                ocov = np.concatenate([oix.coverage(sh[i]) for i in range(n_o)])
                checks["oracle_equal_first_20_sequences"] = bool(np.array_equal(ocov, out[: n_o * per].cpu().numpy().view(np.uint32)))
            if args.cpu:
                # reference C++ (PHASH_MAP::get_freq from all threads) over every window of the first sequences
                threads = os.cpu_count() or 1
                n_c = min(n_seq, 60 * threads)
                wins = seqs[:n_c].unfold(1, 23, 1).reshape(-1, 23).contiguous().cpu().numpy()
                kind, secs, res = bc.cpu_query_runs(prefix, wins, threads, 2)
                cpu = {"value": wins.shape[0] / min(secs), "unit": "positions/s", "cores": threads, "kind": kind,
                       "sample": f"all {wins.shape[0]} windows of the first {n_c} sequences, {threads} std::threads over "
                                 f"PHASH_MAP::get_freq, best of {len(secs)}",
                       "seconds": min(secs),
                       "results_equal_gpu": bool(np.array_equal(res, out[: n_c * per].cpu().numpy().view(np.uint32)))}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
    e2e = None
    if args.e2e:
        # through host buffers (pinned), bounded to 100 k sequences (1 GB in, 4 GB out)
        n_e = min(n_seq, 100_000)
        s_host = ctx.pinned((n_e * seq_len,), np.uint8)
        torch.from_numpy(s_host).copy_(seqs[:n_e].reshape(-1))
        o_host = ctx.pinned((n_e * per,), np.uint32)
        offs_h = (np.arange(n_e + 1, dtype=np.int64) * seq_len)
        torch.cuda.synchronize()

        def e2e_step():
            ctx.check(lib.aix_coverage(ctx.handle, index._h, None, s_host.ctypes.data, offs_h.ctypes.data, n_e, 23, 0,
                                       o_host.ctypes.data))

        e2e_step()
        t0 = time.perf_counter()
        e2e_step()
        e_s = time.perf_counter() - t0
        e2e = {"value": n_e * per / e_s, "unit": "positions/s", "ms_per_step": e_s * 1e3, "sequences": n_e,
               "h2d_bytes_per_step": int(n_e * seq_len), "d2h_bytes_per_step": int(n_e * per * 4),
               "api": "aix_coverage (pinned host buffers, two groups in flight)",
               "matches_device_path": bool(np.array_equal(o_host, out[: n_e * per].cpu().numpy().view(np.uint32)))}
        del s_host, o_host
    del seqs, out
    return {"config": "C4", "workload": f"coverage of {n_seq} x {seq_len} bp sequences (1% substitutions) on the C2 index ({n_keys} keys)",
            "metric": "coverage positions/s", "value": total_out / (ms / 1e3), "unit": "positions/s",
            "sequences_per_s": n_seq / (ms / 1e3), "ms_per_step": ms, "hit_fraction": hit,
            "roofline": roof(total_out, 17, ms, "coverage_kernel<23, canonical>", "coverage_kernel_c4", random_access=True), "e2e": e2e,
            "checks": checks, "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------ C5
def run_c5(ctx, stream, dev, args=None):
    """compute_aindex.cpp:48-115 (positions index build) + python_wrapper.cpp:800-822 (position queries)."""
    args = _ns(args)
    torch, capi, lib = _env(dev)
    n_reads, genome_len = int(50_000_000 * args.scale), int(250_000_000 * args.scale)
    reads = bc.make_reads(torch, dev, genome_len, n_reads, 150, 31, 32)
    n_bytes = reads.numel()
    pad = torch.full((64,), 10, device=dev, dtype=torch.uint8)
    reads = torch.cat([reads.reshape(-1), pad])  # readable past the end (aix_positions_build23_dev contract)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mphf, index, checker_t, tf_t, n_keys = bc.build_index(torch, capi, ctx, reads[:n_bytes])
    ctx.sync()
    index_s = time.perf_counter() - t0
    total_occ = n_reads * 128
    torch.cuda.empty_cache()

    # ---- build, timed with events around the whole call (prefix sum + emit + sort + finalize, allocations included)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pos = capi.Positions.build_dev(index, reads.data_ptr(), n_bytes, 23)  # warm-up
    pos.close()
    ctx.sync()
    a.record(stream)
    t0 = time.perf_counter()
    pos = capi.Positions.build_dev(index, reads.data_ptr(), n_bytes, 23)
    b.record(stream)
    ctx.sync()
    build_wall = time.perf_counter() - t0
    build_ms = a.elapsed_time(b)
    info = pos.info
    ip, pp = pos.device_arrays()
    indices = bc._wrap_device_i64(torch, ip, info["n_indices"], dev)
    positions = bc._wrap_device_i64(torch, pp, info["n_positions"], dev)
    checks = {"total_equals_128_per_read": info["n_positions"] == total_occ}
    g = torch.Generator(device=dev)
    g.manual_seed(33)
    ns = int(10_000_000 * min(1.0, args.scale * 4))
    j = torch.randint(0, info["n_positions"], (ns,), generator=g, device=dev, dtype=torch.int64)
    win = reads[(positions[j] - 1)[:, None] + torch.arange(23, device=dev)[None, :]].contiguous()
    if args.checks:
        # indices == exclusive cumsum of tf (hash.hpp:365-399)
        cs = torch.cumsum(tf_t.to(torch.int64), 0)
        checks["indices_equal_exclusive_cumsum_of_tf"] = bool(indices[0].item() == 0 and torch.equal(indices[1:], cs))
        del cs
        # every slot filled (tf was counted on the same reads), ascending inside every bucket
        chunk = 1 << 28
        nz, asc = True, True
        is_start = torch.zeros(info["n_positions"] + 1, device=dev, dtype=torch.bool)
        is_start[indices] = True
        for s in range(0, info["n_positions"], chunk):
            e = min(info["n_positions"], s + chunk)
            p = positions[s:e]
            nz = nz and bool((p > 0).all().item())
            d_ok = (p[1:] > p[:-1]) | is_start[s + 1:e]
            asc = asc and bool(d_ok.all().item())
            if s > 0:
                asc = asc and bool(positions[s].item() > positions[s - 1].item() or is_start[s].item())
        checks["no_empty_slot"] = nz
        checks["ascending_inside_every_bucket"] = asc
        del is_start
        # the k-mer at every stored position hashes to the bucket that stores it (10 M sampled slots)
        bucket = torch.searchsorted(indices, j, right=True) - 1
        kid = torch.empty(ns, device=dev, dtype=torch.int64)
        index.query_dev(win.data_ptr(), 23, None, ns, capi.Q_PFID, kid.data_ptr())
        ctx.sync()
        checks["kmer_at_position_maps_to_its_bucket"] = bool(torch.equal(kid, bucket))
        del kid, bucket

    # ---- position queries: k-mers sampled from the reads (seed 33), counts pass + fill pass
    nq = ns
    counts = torch.empty(nq, device=dev, dtype=torch.int64)
    offs = torch.zeros(nq + 1, device=dev, dtype=torch.int64)

    def q_counts():
        ctx.check(lib.aix_positions_query_dev(ctx.handle, index._h, None, pos._h, win.data_ptr(), 23, None, nq, 23,
                                              counts.data_ptr(), None, None))

    q_counts()
    ctx.sync()
    offs[1:] = torch.cumsum(counts, 0)
    n_out = int(offs[-1].item())
    pout = torch.empty(n_out, device=dev, dtype=torch.int64)

    def q_both():
        q_counts()
        ctx.check(lib.aix_positions_query_dev(ctx.handle, index._h, None, pos._h, win.data_ptr(), 23, None, nq, 23,
                                              None, offs.data_ptr(), pout.data_ptr()))

    q_ms = timed(torch, ctx, stream, q_both, reps=3, warmup=1)
    # bytes of one query: 23 in + 8 checker + 16 indices + 8 per position returned (read) + 8 (written)
    q_bytes = nq * (23 + 8 + 16 + 8) + 2 * 8 * n_out
    peak, _ = bc.peak_hbm_gbs()
    queries = {"value": nq / (q_ms / 1e3), "unit": "queries/s (counts pass + fill pass)", "ms_per_step": q_ms,
               "positions_returned": n_out, "positions_per_s": n_out / (q_ms / 1e3),
               "roofline": {"bound": "hbm", "kernel": "positions_query_kernel<23>", "achieved": q_bytes / (q_ms / 1e3) / 1e9,
                            "peak": peak, "unit": "GB/s", "frac": q_bytes / (q_ms / 1e3) / 1e9 / peak,
                            "bytes_per_launch": q_bytes}}
    if args.checks:
        tfq = torch.empty(nq, device=dev, dtype=torch.int32)
        index.query_dev(win.data_ptr(), 23, None, nq, capi.Q_TF, tfq.data_ptr())
        ctx.sync()
        checks["len_positions_equals_tf"] = bool(torch.equal(counts, tfq.to(torch.int64)))  # test_aindex_functionality.py:376-380
        own = positions[j] - 1
        seg = torch.repeat_interleave(torch.arange(nq, device=dev), counts)
        found = torch.zeros(nq, device=dev, dtype=torch.bool)
        found[seg[pout == own[seg]]] = True
        checks["query_returns_the_sampled_position"] = bool(found.all().item())
        del seg, found, tfq, own
    del pout, counts, offs, win, j
    pos.close()
    del indices, positions
    ctx.trim()  # the builders' pool keeps ~100 GB of freed blocks for the next build: hand them back before torch needs HBM
    torch.cuda.empty_cache()

    # ---- the reference compute_aindex on a read subsample with its own index (1 thread = parity order), and the
    #      host-buffer C-ABI call (e2e) on the same subsample
    cpu, e2e = None, None
    cbin = os.path.join(REF_BIN, "compute_aindex")
    n_sub = min(n_reads, 200_000)
    sub_t = torch.cat([reads[: n_sub * 151], pad])
    m2, ix2, chk2, tf2, n2 = bc.build_index(torch, capi, ctx, sub_t[: n_sub * 151])
    p2 = capi.Positions.build_dev(ix2, sub_t.data_ptr(), n_sub * 151, 23)
    gi, gp = p2.download()
    p2.close()
    if args.cpu and os.path.exists(cbin):
        tmpdir = tempfile.mkdtemp(prefix="aix_c5_", dir=bc._tmp_root())
        try:
            prefix = bc.write_index_files(tmpdir, m2, chk2.cpu().numpy().view(np.uint64), tf2.cpu().numpy().view(np.uint32))
            reads[: n_sub * 151].cpu().numpy().tofile(prefix + ".reads")
            threads = os.cpu_count() or 1
            res, ri, rp = {}, None, None
            for th in (1, threads):
                t0 = time.perf_counter()
                r = subprocess.run([cbin, prefix + ".reads", prefix + ".pf", prefix, str(th), "23", prefix + ".tf.bin",
                                    prefix + ".kmers.bin", prefix + ".kmers"], stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True)
                res[th] = (time.perf_counter() - t0, r.returncode)
                if th == 1 and r.returncode == 0:
                    ri = np.fromfile(prefix + ".indices.bin", dtype=np.uint64)
                    rp = np.fromfile(prefix + ".index.bin", dtype=np.uint64)
            if res[1][1] == 0:
                cpu = {"value": n_sub * 128 / res[threads][0], "unit": "occurrences/s (wall, incl. index load)", "cores": threads,
                       "kind": "reference", "sample": f"compute_aindex on the first {n_sub} reads with their own index ({n2} keys)",
                       "seconds_1_thread": res[1][0], "seconds_all_threads": res[threads][0],
                       "indices_bin_equal": bool(np.array_equal(ri, gi)), "index_bin_equal_1_thread": bool(np.array_equal(rp, gp))}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
    if args.e2e:
        n_e = min(n_reads, 5_000_000)   # 0.76 GB of reads in, 5.1 GB of positions + indices out
        if n_e != n_sub:
            sub_t = torch.cat([reads[: n_e * 151], pad])
            m2, ix2, chk2, tf2, n2 = bc.build_index(torch, capi, ctx, sub_t[: n_e * 151])
        r_host = ctx.pinned((n_e * 151,), np.uint8)
        torch.from_numpy(r_host).copy_(sub_t[: n_e * 151])
        i_host = ctx.pinned((n2 + 1,), np.uint64)
        p_host = ctx.pinned((n_e * 128,), np.uint64)
        torch.cuda.synchronize()

        def host_build():
            ctx.check(lib.aix_positions_build23(ctx.handle, ix2._h, r_host.ctypes.data, n_e * 151, i_host.ctypes.data, p_host.ctypes.data))

        host_build()
        t0 = time.perf_counter()
        host_build()
        e_s = time.perf_counter() - t0
        p3 = capi.Positions.build_dev(ix2, sub_t.data_ptr(), n_e * 151, 23)
        di, dp = p3.download()
        p3.close()
        e2e = {"value": n_e * 128 / e_s, "unit": "occurrences/s", "ms_per_step": e_s * 1e3, "reads": n_e, "index_keys": n2,
               "h2d_bytes_per_step": int(n_e * 151), "d2h_bytes_per_step": int((n2 + 1) * 8 + n_e * 128 * 8),
               "api": "aix_positions_build23 (pinned host buffers)",
               "matches_device_path": bool(np.array_equal(i_host, di) and np.array_equal(p_host, dp))}
        del r_host, i_host, p_host
    del reads, sub_t
    ctx.trim()
    return {"config": "C5", "workload": f"positions index over {n_reads} x 150 bp reads ({n_keys} keys, {total_occ} occurrences) + {nq} position queries",
            "metric": "positions-index occurrences/s", "value": total_occ / (build_ms / 1e3), "unit": "occurrences/s",
            "ms_per_step": build_ms, "build_wall_s": build_wall, "index_build_s": index_s,
            "roofline": roof(total_occ, 29, build_ms, "positions_emit_kernel<23> + rs_pass_kernel x4 + positions_finalize_kernel (whole build)",
                             "positions_build_c5"),
            "queries": queries, "e2e": e2e, "checks": checks, "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------ K1
def run_k1(ctx, stream, dev, args=None):
    """The encoding kernels on their own: dna_bitset packing (dna_bitseq.hpp:22-61) and the rolling forward /
    reverse-complement 23-mers of a reads buffer (the loop of hash.cpp:1006-1032 without the lookup), input resident in HBM.
    No CPU baseline: the reference never runs these loops on their own (the packing class is unused by its pipeline)."""
    args = _ns(args)
    torch, capi, lib = _env(dev)
    n_reads = max(1600, int(3_500_000 * args.scale) // 16 * 16)       # n_reads * 151 stays a multiple of 16
    reads = bc.make_reads(torch, dev, 50_000_000, n_reads, 150, 51, 52).reshape(-1)
    n = reads.numel()
    packed = torch.empty((n + 3) // 4 + 16, device=dev, dtype=torch.uint8)
    fwd = torch.empty(n, device=dev, dtype=torch.int64)
    rcv = torch.empty(n, device=dev, dtype=torch.int64)
    valid = torch.empty(n + 16, device=dev, dtype=torch.uint8)

    def pack_step():
        ctx.check(lib.aix_pack_2bit_dev(ctx.handle, reads.data_ptr(), n, packed.data_ptr()))

    def roll_step():
        ctx.check(lib.aix_rolling_kmers_dev(ctx.handle, reads.data_ptr(), n, 23, fwd.data_ptr(), rcv.data_ptr(), valid.data_ptr()))

    pack_ms = timed(torch, ctx, stream, pack_step, reps=10, warmup=3)
    roll_ms = timed(torch, ctx, stream, roll_step, reps=5, warmup=2)
    checks = {}
    if args.checks:
        # independent arithmetic on a sample: 4 bases per byte, first base in the top bits, anything but ACGT -> 0;
        # window value = sum of code << 2 (22 - j), valid iff 23 upper-case ACGT letters
        rng = np.random.default_rng(5)
        code = np.zeros(256, dtype=np.uint64)
        code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
        starts = np.concatenate([rng.integers(0, n - 64, size=4000), np.arange(0, 400)]).astype(np.int64) // 4 * 4
        idx = torch.from_numpy(starts[:, None] + np.arange(28)[None, :]).to(dev)
        win = reads[idx].cpu().numpy()
        c = code[win]
        want_p = (c[:, 0:28:4] << np.uint64(6)) | (c[:, 1:28:4] << np.uint64(4)) | (c[:, 2:28:4] << np.uint64(2)) | c[:, 3:28:4]
        got_p = packed[torch.from_numpy(starts[:, None] // 4 + np.arange(7)[None, :]).to(dev)].cpu().numpy()
        checks["pack_equals_numpy"] = bool(np.array_equal(got_p.astype(np.uint64), want_p))
        w23 = win[:, :23]
        ok = np.isin(w23, np.frombuffer(b"ACGT", dtype=np.uint8)).all(axis=1)
        sh = (np.uint64(2) * (np.uint64(22) - np.arange(23, dtype=np.uint64)))[None, :]
        want_f = (c[:, :23] << sh).sum(axis=1, dtype=np.uint64)
        want_r = ((np.uint64(3) - c[:, :23]) << (np.uint64(2) * np.arange(23, dtype=np.uint64))[None, :]).sum(axis=1, dtype=np.uint64)
        st = torch.from_numpy(starts).to(dev)
        got_v = valid[st].cpu().numpy().astype(bool)
        got_f = fwd[st].cpu().numpy().view(np.uint64)
        got_r = rcv[st].cpu().numpy().view(np.uint64)
        checks["valid_equals_numpy"] = bool(np.array_equal(got_v, ok))
        checks["rolling_equals_numpy_on_valid_windows"] = bool(np.array_equal(got_f[ok], want_f[ok]) and np.array_equal(got_r[ok], want_r[ok]))
        checks["valid_windows_in_sample"] = int(ok.sum())
    windows = n - 22
    del packed, fwd, rcv, valid, reads
    return {"config": "K1", "workload": f"{n_reads} x 150 bp reads ({n} bytes, plain format) resident in HBM: 2-bit packing; rolling 23-mers (forward, "
                                        "reverse complement, validity) of every window",
            "pack": {"metric": "bases packed/s", "value": n / (pack_ms / 1e3), "unit": "bases/s", "ms_per_step": pack_ms,
                     "roofline": roof(n, 1.25, pack_ms, "pack2bit_vec_kernel")},
            "rolling": {"metric": "windows/s", "value": windows / (roll_ms / 1e3), "unit": "windows/s", "ms_per_step": roll_ms,
                        "roofline": roof(windows, 18, roll_ms, "rolling_vec_kernel<23> (1 B in, 8 + 8 + 1 B out per window)")},
            "checks": checks, "cpu_baseline": None}


RUNNERS = {"c1": run_c1, "c4": run_c4, "c5": run_c5, "k1": run_k1}


def main():
    import torch
    from aindex_b200 import capi
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c4,c5")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the BASELINE sizes (smoke runs)")
    ap.add_argument("--no-checks", dest="checks", action="store_false", default=True)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.set_stream(stream)  # one stream for torch and the library: allocator reuse stays ordered
    lines = []
    for name in args.configs.split(","):
        t0 = time.perf_counter()
        line = RUNNERS[name.strip().lower()](ctx, stream, dev, types.SimpleNamespace(scale=args.scale, checks=args.checks))
        line["wall_s"] = time.perf_counter() - t0
        line["scale"] = args.scale
        print(json.dumps(line), flush=True)
        lines.append(line)
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
