#!/usr/bin/env python3
"""Measurement of the BASELINE.json configs that are not bench.py's headline line:

  C1  tf query of all 4^13 13-mers in numeric order (25 B / lookup)
  C4  sequence coverage, 1 M x 10 kb sequences on the C2 index (17 B / position)
  C5  positions index build over 50 M x 150 bp reads + 10 M position queries (29 B / occurrence)

Every config is run at BASELINE size on one B200 with all buffers resident in HBM, timed with
CUDA events on the library stream, checked through size-independent properties at full size and
against the CPU oracle / the compiled reference on a bounded sample, and printed as one JSON line.
Test infrastructure (it lives under tests/ because it checks against oracle/ and the compiled reference).

  python tests/bench_configs.py --configs c1,c4,c5 [--scale 1.0] [--out gpurun_out/configs.json]
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

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402  (synthetic data generators shared with the headline benchmark)
from aindex_b200 import capi  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
_tmp_root = bench._tmp_root


def timed(ctx, stream, fn, reps=3, warmup=1):
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


def roof(units, bytes_per_unit, ms):
    ach = units * bytes_per_unit / (ms / 1e3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": PEAK, "unit": "GB/s", "frac": ach / PEAK,
            "bytes_per_unit": bytes_per_unit, "units_per_launch": int(units), "kernel_ms": ms}


# ------------------------------------------------------------------------------------------ C1
def run_c1(ctx, stream, dev, args):
    """count13 over synthetic reads -> 13-mer index -> tf query of all 4^13 13-mers."""
    lib = capi.lib()
    n_reads = max(1000, int(1_000_000 * args.scale))
    reads = bench.make_reads(torch, dev, 5_000_000, n_reads, 150, 41, 42)
    pf = os.path.join(ROOT, "oracle", "_ref", "data", "all_13mers.pf")
    if os.path.exists(pf):
        m13, pf_kind = capi.Mphf.load(ctx, pf), "reference all_13mers.pf"
    else:
        allk = torch.arange(1 << 26, device=dev, dtype=torch.int64)
        m13, pf_kind = capi.Mphf.build_dev(ctx, allk.data_ptr(), 1 << 26, 13), "GPU-built MPHF"
        del allk
    tf, stats = ctx.count13(m13, reads.cpu().numpy().reshape(-1), capi.FMT_PLAIN)
    ix = capi.Index13.upload(ctx, m13, tf)
    # all 4^13 13-mers, numeric order, as 13-byte records
    v = torch.arange(1 << 26, device=dev, dtype=torch.int64)
    lut = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    recs = torch.empty((1 << 26, 13), device=dev, dtype=torch.uint8)
    for j in range(13):
        recs[:, j] = lut[(v >> (2 * (12 - j))) & 3]
    out = torch.empty(1 << 26, device=dev, dtype=torch.int32)
    q = 1 << 26

    def step():
        ctx.check(lib.aix_tf13_batch_dev(ctx.handle, ix._h, recs.data_ptr(), 13, None, q, capi.Q_TF, out.data_ptr()))

    ms = timed(ctx, stream, step, reps=5, warmup=3)
    got = out.cpu().numpy().view(np.uint32)
    perm = m13.perm13()
    ok_perm = bool(np.array_equal(got.astype(np.uint64), tf[perm]))  # query(v) == tf[mphf(v)]
    ok_sum = int(got.sum()) == int(stats["valid"])
    # CPU: the oracle port on a 2 M sample (1 thread), reference semantics python_wrapper.cpp:482-503
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    sample = recs[: 8_000_000].cpu().numpy()
    cpu = None
    if os.path.exists(pf):
        oix = O.Index13(O.Mphf.load(pf), tf)
        t0 = time.perf_counter()
        ores = oix.batch(sample, None, O.MODE_TF, threads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": sample.shape[0] / dt, "unit": "lookups/s", "cores": threads, "kind": "port",
               "sample": f"first {sample.shape[0]} of the 4^13 13-mers, oracle get_tf_value_13mer (MPHF lookup + tf gather), {threads} threads",
               "results_equal_gpu": bool(np.array_equal(ores, got[: sample.shape[0]]))}
    return {"config": "C1", "workload": f"count13 over {n_reads} reads -> tf query of all 4^13 13-mers ({pf_kind})",
            "metric": "13-mer tf lookups/s", "value": q / (ms / 1e3), "unit": "lookups/s", "ms_per_step": ms,
            "roofline": roof(q, 25, ms), "checks": {"query_equals_tf_of_perm13": ok_perm, "sum_equals_valid_windows": ok_sum},
            "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------ C4
def make_sequences(dev, genome_codes, n_seq, seq_len, seed, sub_rate=0.01):
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


def run_c4(ctx, stream, dev, args):
    lib = capi.lib()
    n_reads, genome_len = int(10_000_000 * args.scale), int(50_000_000 * args.scale)
    reads = bench.make_reads(torch, dev, genome_len, n_reads, 150, 1, 2)
    mphf, index, checker_t, tf_t, n_keys = bench.build_index(torch, capi, ctx, reads)
    del reads
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    genome = torch.randint(0, 4, (genome_len,), generator=g, device=dev, dtype=torch.uint8)  # same stream as make_reads
    n_seq, seq_len = int(1_000_000 * args.scale), 10_000
    seqs = make_sequences(dev, genome, n_seq, seq_len, 21)
    del genome
    torch.cuda.empty_cache()
    offs = torch.arange(n_seq + 1, device=dev, dtype=torch.int64) * seq_len
    per = seq_len - 22
    total_out = n_seq * per
    out = torch.empty(total_out, device=dev, dtype=torch.int32)

    def step():
        ctx.check(lib.aix_coverage_dev(ctx.handle, index._h, None, seqs.data_ptr(), offs.data_ptr(), n_seq,
                                       seqs.numel(), total_out, 23, 0, out.data_ptr()))

    ms = timed(ctx, stream, step, reps=3, warmup=1)
    hit = float((out[: 50_000_000] > 0).float().mean().item())
    # property at full size: coverage[s, i] == batch tf query of the window seq[s, i:i+23] (random sample)
    g.manual_seed(5)
    ns = 5_000_000
    si = torch.randint(0, n_seq, (ns,), generator=g, device=dev, dtype=torch.int64)
    oi = torch.randint(0, per, (ns,), generator=g, device=dev, dtype=torch.int64)
    win = seqs.reshape(-1)[(si * seq_len + oi)[:, None] + torch.arange(23, device=dev)[None, :]].contiguous()
    qout = torch.empty(ns, device=dev, dtype=torch.int32)
    index.query_dev(win.data_ptr(), 23, None, ns, capi.Q_TF, qout.data_ptr())
    ctx.sync()
    ok_prop = bool(torch.equal(qout, out[si * per + oi]))
    # oracle on whole sequences (bounded): the reference loop aindex.py:314-322 restated in C
    from oracle import oracle as O
    tmpdir = tempfile.mkdtemp(prefix="aix_c4_", dir=_tmp_root())
    cpu = None
    try:
        prefix = bench.write_index_files(tmpdir, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
        oix = O.Index23.load_prefix(prefix)
        n_o = 20
        sh = seqs[:n_o].cpu().numpy()
        ocov = np.concatenate([oix.coverage(sh[i]) for i in range(n_o)])
        ok_oracle = bool(np.array_equal(ocov, out[: n_o * per].cpu().numpy().view(np.uint32)))
        # reference C++ (PHASH_MAP::get_freq from all threads) over every window of the first 1000 sequences
        threads = os.cpu_count() or 1
        n_c = min(n_seq, 1000)
        wins = seqs[:n_c].unfold(1, 23, 1).reshape(-1, 23).contiguous().cpu().numpy()
        kind, secs, res = bench.cpu_query_runs(prefix, wins, threads, 2)
        cpu = {"value": wins.shape[0] / min(secs), "unit": "positions/s", "cores": threads, "kind": kind,
               "sample": f"all {wins.shape[0]} windows of the first {n_c} sequences, {threads} threads over PHASH_MAP::get_freq",
               "results_equal_gpu": bool(np.array_equal(res, out[: n_c * per].cpu().numpy().view(np.uint32)))}
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    # e2e through host buffers (pinned), bounded to 100 k sequences (1 GB in, 4 GB out)
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
    ok_e2e = bool(np.array_equal(o_host, out[: n_e * per].cpu().numpy().view(np.uint32)))
    return {"config": "C4", "workload": f"coverage of {n_seq} x {seq_len} bp sequences (1% substitutions) on the C2 index ({n_keys} keys)",
            "metric": "coverage positions/s", "value": total_out / (ms / 1e3), "unit": "positions/s",
            "sequences_per_s": n_seq / (ms / 1e3), "ms_per_step": ms, "hit_fraction": hit,
            "roofline": roof(total_out, 17, ms),
            "e2e": {"value": n_e * per / e_s, "unit": "positions/s", "sequences": n_e, "h2d_bytes_per_step": int(n_e * seq_len),
                    "d2h_bytes_per_step": int(n_e * per * 4), "matches_device_path": ok_e2e},
            "checks": {"coverage_equals_batch_tf_on_5M_sampled_windows": ok_prop, "oracle_equal_first_20_sequences": ok_oracle},
            "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------ C5
def run_c5(ctx, stream, dev, args):
    lib = capi.lib()
    n_reads, genome_len = int(50_000_000 * args.scale), int(250_000_000 * args.scale)
    reads = bench.make_reads(torch, dev, genome_len, n_reads, 150, 31, 32)
    n_bytes = reads.numel()
    pad = torch.full((64,), 10, device=dev, dtype=torch.uint8)
    reads = torch.cat([reads.reshape(-1), pad])  # readable past the end (aix_positions_build23_dev contract)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mphf, index, checker_t, tf_t, n_keys = bench.build_index(torch, capi, ctx, reads[:n_bytes])
    ctx.sync()
    index_s = time.perf_counter() - t0
    total_occ = n_reads * 128
    torch.cuda.empty_cache()

    # ---- build, timed with events around the whole call (prefix sum + count + scatter + sort)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pos = capi.Positions.build_dev(index, reads.data_ptr(), n_bytes, 23)  # warm-up (allocations)
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
    indices = bench._wrap_device_i64(torch, ip, info["n_indices"], dev)
    positions = bench._wrap_device_i64(torch, pp, info["n_positions"], dev)
    checks = {}
    checks["total_equals_128_per_read"] = info["n_positions"] == total_occ
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
    g = torch.Generator(device=dev)
    g.manual_seed(33)
    ns = int(10_000_000 * min(1.0, args.scale * 4))
    j = torch.randint(0, info["n_positions"], (ns,), generator=g, device=dev, dtype=torch.int64)
    bucket = torch.searchsorted(indices, j, right=True) - 1
    win = reads[(positions[j] - 1)[:, None] + torch.arange(23, device=dev)[None, :]].contiguous()
    kid = torch.empty(ns, device=dev, dtype=torch.int64)
    index.query_dev(win.data_ptr(), 23, None, ns, capi.Q_PFID, kid.data_ptr())
    ctx.sync()
    checks["kmer_at_position_maps_to_its_bucket"] = bool(torch.equal(kid, bucket))

    # ---- position queries: k-mers sampled from the reads (seed 33), counts pass + fill pass
    nq = int(10_000_000 * min(1.0, args.scale * 4))
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

    q_ms = timed(ctx, stream, q_both, reps=3, warmup=1)
    tfq = torch.empty(nq, device=dev, dtype=torch.int32)
    index.query_dev(win.data_ptr(), 23, None, nq, capi.Q_TF, tfq.data_ptr())
    ctx.sync()
    checks["len_positions_equals_tf"] = bool(torch.equal(counts, tfq.to(torch.int64)))  # test_aindex_functionality.py:376-380
    own = positions[j] - 1
    seg = torch.repeat_interleave(torch.arange(nq, device=dev), counts)
    found = torch.zeros(nq, device=dev, dtype=torch.bool)
    found[seg[pout == own[seg]]] = True
    checks["query_returns_the_sampled_position"] = bool(found.all().item())
    del seg, found, pout

    # ---- reference compute_aindex on a read subsample (same index files), 1 thread = parity order
    cpu = None
    cbin = os.path.join(REF_BIN, "compute_aindex")
    if os.path.exists(cbin):
        n_sub = min(n_reads, 200_000)
        tmpdir = tempfile.mkdtemp(prefix="aix_c5_", dir=_tmp_root())
        try:
            sub = reads[: n_sub * 151].cpu().numpy()
            # index of the subsample (tf must be counted on the same reads)
            sub_t = torch.cat([reads[: n_sub * 151], pad])
            m2, ix2, chk2, tf2, n2 = bench.build_index(torch, capi, ctx, sub_t[: n_sub * 151])
            prefix = bench.write_index_files(tmpdir, m2, chk2.cpu().numpy().view(np.uint64), tf2.cpu().numpy().view(np.uint32))
            sub.tofile(prefix + ".reads")
            threads = os.cpu_count() or 1
            res = {}
            for th in (1, threads):
                t0 = time.perf_counter()
                r = subprocess.run([cbin, prefix + ".reads", prefix + ".pf", prefix, str(th), "23", prefix + ".tf.bin",
                                    prefix + ".kmers.bin", prefix + ".kmers"], stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True)
                res[th] = (time.perf_counter() - t0, r.returncode)
                if th == 1 and r.returncode == 0:
                    ri = np.fromfile(prefix + ".indices.bin", dtype=np.uint64)
                    rp = np.fromfile(prefix + ".index.bin", dtype=np.uint64)
            p2 = capi.Positions.build_dev(ix2, sub_t.data_ptr(), n_sub * 151, 23)
            gi, gp = p2.download()
            p2.close()
            if res[1][1] == 0:
                cpu = {"value": n_sub * 128 / res[threads][0], "unit": "occurrences/s (wall, incl. index load)", "cores": threads,
                       "kind": "reference", "sample": f"compute_aindex on the first {n_sub} reads with their own index ({n2} keys)",
                       "seconds_1_thread": res[1][0], "seconds_all_threads": res[threads][0],
                       "indices_bin_equal": bool(np.array_equal(ri, gi)), "index_bin_equal_1_thread": bool(np.array_equal(rp, gp))}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
    return {"config": "C5", "workload": f"positions index over {n_reads} x 150 bp reads ({n_keys} keys, {info['n_positions']} occurrences) + {nq} position queries",
            "metric": "positions-index occurrences/s", "value": total_occ / (build_ms / 1e3), "unit": "occurrences/s",
            "ms_per_step": build_ms, "build_wall_s": build_wall, "index_build_s": index_s, "roofline": roof(total_occ, 29, build_ms),
            "queries": {"value": nq / (q_ms / 1e3), "unit": "queries/s (counts pass + fill pass)", "ms_per_step": q_ms,
                        "positions_returned": n_out},
            "checks": checks, "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c4,c5")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the BASELINE sizes (smoke runs)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.set_stream(stream)  # one stream for torch and the library: allocator reuse stays ordered
    lines = []
    for name in args.configs.split(","):
        fn = {"c1": run_c1, "c4": run_c4, "c5": run_c5}[name.strip().lower()]
        t0 = time.perf_counter()
        line = fn(ctx, stream, dev, args)
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
