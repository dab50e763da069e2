#!/usr/bin/env python3
"""bench.py -- headline benchmark of the aindex hot path on B200.

Metric (BASELINE.json): 23-mer batch tf queries/s on config C2 -- a 23-mer index over
10 M synthetic 150 bp reads (50 Mbp random genome), 100 M uniform-random 23-mer queries
(~100 % misses: the workload the reference's own "2.3 M q/s" stress test measures).  A step
is one pass of the batch-lookup path over the 100 M-query batch.  13-mer counting
(k-mers counted/s, config C3's per-GPU shard) is measured in the same run and reported
under "extra".

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference [...]                        the reference's CPU path
  torchrun ... bench.py --gpus N ...                             one rank per GPU (weak scaling)

Our arm:   value = whole-job queries/s, inputs resident in HBM (CUDA events on the launch
           stream, max over ranks); e2e = the same through the host-buffer C-ABI call
           (pinned host memory, H2D + kernel + D2H inside the timed region).
Reference: PHASH_MAP::get_freq (src/hash.hpp:123-140) from all host threads through
           oracle/_ref/bin/ref_harness (unmodified reference code), or the C oracle port when
           the reference was not compiled; a bounded sample of the same queries per step.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

Q1_BYTES_PER_QUERY = 43       # 23 B in + 4 B out + 2 x 8 B checker
Q1_BYTES_ONE_PROBE = 35       # what the canonical-first kernel really needs: 23 + 4 + 8
C3_BYTES_PER_KMER = 9.09      # 151/138 B in + 4 B read + 4 B write
# measured on this pool's B200 by profiles/atomic_roofline.cu (profiles/r01_atomic_roofline.txt)
GATHER_PEAK_G = 50.1          # random 16-byte gathers/s (1 GiB table), G/s
RED_PEAK_256M_G = 32.0        # random RED.ADD.U32 into a 256 MiB table (the north star's atomic roofline), G/s
RED_PEAK_L2_G = 210.0         # the same into a 64 MiB (L2-resident) slice, G/s
# request-only kernels in the shape of one lookup (profiles/atomic_roofline.cu `lookup_shape_stream`, profiles/r02_atomic_roofline.txt):
# 23 B streamed in through the product kernel's TMA ring, 4 B out, 3 scattered 16-byte loads from an L2-resident table, nothing else
LOOKUP_SHAPE_16MIB_G = 65.0   # 16 MiB record table
LOOKUP_SHAPE_64MIB_G = 53.4   # 64 MiB record table (the fused C2 structure is 61.5 MB)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="C2 reads (150 bp)")
    ap.add_argument("--genome", type=int, default=50_000_000, help="C2 genome length")
    ap.add_argument("--queries", type=int, default=100_000_000, help="queries per GPU per step")
    ap.add_argument("--count-reads", type=int, default=25_000_000, help="C3 reads per GPU (0 = skip)")
    ap.add_argument("--count-genome", type=int, default=100_000_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-legs", dest="multi_legs", action="store_false", default=True,
                    help="skip the single-process (C-ABI) multi-GPU legs at N > 1")
    ap.add_argument("--multi-positions-reads", type=int, default=50_000_000, help="reads of the N-GPU positions build (C5 = 50 M)")
    ap.add_argument("--configs", default="c4,c1,c5,k1", help="other BASELINE configs (and k1 = the encoding kernels) measured at N=1 ('' = none)")
    ap.add_argument("--config-scale", type=float, default=1.0, help="fraction of the BASELINE sizes for --configs (smoke runs)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


from bench_common import PF13 as PF13_PATH  # noqa: E402
from bench_common import (ClockSampler, _tmp_root, _wrap_device_i64, build_index, cpu_count_run, cpu_query_runs,  # noqa: E402,F401
                          make_hit_queries, make_queries, make_reads, ncu_traffic, peak_hbm_gbs, ref_harness_path,
                          write_index_files)


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def c2_config(args, world, n_keys):
    """The `config` object of the JSON line: identical in both arms (the reference arm times the same 100 M-query
    batch per step on the same index; it has one host, so its rate does not depend on `world`)."""
    return {"workload": "C2: 23-mer index over 10M synthetic 150bp reads; 100M random batch tf queries (Q1, ~100% miss) per GPU",
            "reads": args.reads, "genome_bp": args.genome, "queries_per_gpu": args.queries, "index_keys": n_keys,
            "parallelism": f"replicated index, queries sharded x{world}",
            "l2": "inputs larger than L2 (2.3 GB of queries, 0.8 GB index per pass); no flush needed"}


def _guard(fn):
    """A config run must not take the headline line down with it: report the error in its block instead."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        return {"error": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------
# the C2 index built by the UNMODIFIED reference tools, cached per box
# ------------------------------------------------------------------------------------------
def ref_index_cache_dir(args):
    root = _tmp_root(4 << 30) or tempfile.gettempdir()
    return os.path.join(root, f"aix_bench_cache_c2_{args.reads}_{args.genome}")


def build_reference_index_c2(args, reads_np, cache):
    """reads (uint8 host array) -> canonical 23-mer table (oracle, CPU; definition tests/analyze_kmers.py:25-33) ->
    compute_mphf_seq + compute_index (unmodified reference binaries) -> {cache}/c2.23.{pf,kmers.bin,tf.bin} + meta.json.
    No product code is involved.  Returns the meta dict, or None when the reference was not compiled."""
    from oracle import oracle as O
    meta_p = os.path.join(cache, "meta.json")
    if os.path.exists(meta_p):
        return json.load(open(meta_p))
    os.makedirs(cache, exist_ok=True)
    t0 = time.perf_counter()
    kmers, counts = O.canonical23_count(reads_np)
    t1 = time.perf_counter()
    tool_s = O.build_reference_index23(kmers, counts, os.path.join(cache, "c2.23"))
    if tool_s is None:
        return None
    meta = {"index_keys": int(kmers.size), "canonical_count_s": t1 - t0, **tool_s,
            "built_by": "oracle canonical23_count + oracle/_ref/bin/compute_mphf_seq + compute_index"}
    with open(meta_p + ".tmp", "w") as f:
        json.dump(meta, f)
    os.replace(meta_p + ".tmp", meta_p)
    return meta


def reference_built_index_check(torch, capi, ctx, stream, args, q_dev, out_dev, n_keys):
    """north_star: "MPHF construction reuses the reference's emphf output".  When the reference arm (run first by the
    driver) has left its reference-built C2 index in the box's cache, load THOSE files and answer the same 100 M
    queries with them: answers must equal the GPU-built index's, and the rate is reported next to the headline."""
    cache = ref_index_cache_dir(args)
    prefix = os.path.join(cache, "c2.23")
    if not os.path.exists(os.path.join(cache, "meta.json")):
        return {"available": False, "note": "no reference-built index cached on this box (the --impl reference arm builds it; "
                                            "tests/test_gpu_fullsize.py::test_c2_reference_built_index builds and checks it as well)"}
    try:
        meta = json.load(open(os.path.join(cache, "meta.json")))
        ix = capi.Index23.load_prefix(ctx, prefix)
        out2 = torch.empty_like(out_dev)

        def step():
            ix.query_dev(q_dev.data_ptr(), 23, None, args.queries, capi.Q_TF, out2.data_ptr())

        for _ in range(3):
            step()
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5):
            step()
        b.record(stream)
        ctx.sync()
        ms = a.elapsed_time(b) / 5
        return {"available": True, "index_keys": meta["index_keys"], "same_key_count_as_gpu_built": meta["index_keys"] == n_keys,
                "answers_equal_gpu_built_index": bool(torch.equal(out2, out_dev)), "queries": args.queries,
                "value": args.queries / (ms / 1e3), "unit": "queries/s (device resident, reference-built .pf/.kmers.bin/.tf.bin)",
                "ms_per_step": ms, "layout": ix.layout if hasattr(ix, "layout") else None, "built_by": meta.get("built_by")}
    except Exception as e:  # noqa: BLE001
        return {"available": False, "error": f"{type(e).__name__}: {e}"}


def sharded_index_check(torch, dist, capi, ctx, stream, dev, args, rank, world, mphf, checker_t, tf_t, n_keys, q_dev, out_dev,
                        barrier, max_over_ranks):
    """The index split by hash-id range over the ranks (north_star: "split by hash range when it exceeds HBM"):
    one `query` collective per rank batch; answers must equal the replicated index's, and on rank 0 the
    unmodified reference's on the first 1 M queries."""
    from aindex_b200 import dist as D
    nq = min(args.queries, 20_000_000)
    shared = None
    try:
        # every rank must split THE SAME index: the GPU MPHF builder's peeling order (hence the ids) differs from run to
        # run, so rank 0 writes its files once and every rank loads its slice of them
        pbox = [None]
        if rank == 0:
            shared = tempfile.mkdtemp(prefix="aix_shard_", dir=_tmp_root())
            pbox[0] = write_index_files(shared, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
        dist.broadcast_object_list(pbox, src=0)
        prefix0 = pbox[0]
        m_sh = capi.Mphf.load(ctx, prefix0 + ".pf")
        chk = np.memmap(prefix0 + ".kmers.bin", dtype=np.uint64, mode="r")
        tfs = np.memmap(prefix0 + ".tf.bin", dtype=np.uint32, mode="r")
        sh = D.ShardedIndex23(n_keys).attach(ctx, m_sh, chk, tfs, stream)
        recs = q_dev[:nq]
        res = sh.query(recs)  # warm-up (NCCL all-to-all buffers)
        ctx.sync()
        barrier()
        times = []  # five batches, each bracketed by a barrier; the median (one all-to-all in five hit a 10x outlier at N = 4)
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            res = sh.query(recs)
            b.record(stream)
            ctx.sync()
            barrier()
            times.append(max_over_ranks(a.elapsed_time(b)))
        ms = sorted(times)[len(times) // 2]
        diff = torch.nonzero(res.to(torch.int32) != out_dev[:nq]).reshape(-1)
        if diff.numel():  # diagnosis on stderr: which queries, what the two paths say
            d = diff[:8]
            sys.stderr.write(f"[bench] rank {rank}: sharded index differs from the replicated one on {diff.numel()} of {nq} queries; "
                             f"first {d.tolist()}: sharded {res[d].tolist()} replicated {out_dev[:nq][d].tolist()} "
                             f"(replicated hits in the batch: {int((out_dev[:nq] > 0).sum())}, sharded hits: {int((res > 0).sum())})\n")
        same = torch.tensor([1 if diff.numel() == 0 else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out = {"equal": bool(int(same.item())), "queries_per_rank": nq, "value": world * nq / (ms / 1e3), "unit": "queries/s (aggregate)",
               "ms_per_step": ms, "ms_all_batches": times, "records_per_rank": sh.hi - sh.lo, "equals_reference_1M": None}
        if rank == 0 and ref_harness_path():
            n1 = min(nq, 1_000_000)
            kind, secs, ref = cpu_query_runs(prefix0, recs[:n1].cpu().numpy(), os.cpu_count() or 1, 1)
            out["equals_reference_1M"] = bool(kind == "reference" and np.array_equal(ref, res[:n1].cpu().numpy().astype(np.uint32)))
        return out
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        return {"equal": None, "error": f"{type(e).__name__}: {e}"}
    finally:
        barrier()  # every rank is done with rank 0's files
        if shared:
            shutil.rmtree(shared, ignore_errors=True)


def count13_multi_gpu_vs_reference(torch, dist, capi, ctx, stream, dev, rank, world, creads, peer, hist_tensor, rs_out):
    """Multi-GPU counting against the UNMODIFIED reference binary: every rank counts the first reads of its shard,
    the histograms are combined over k-mer ranges exactly as in the timed step (fused peer-memory combine, else NCCL
    reduce-scatter), the slices are gathered on rank 0 and permuted into .tf.bin order; rank 0 runs count_kmers13 on
    the concatenation of the samples.  Also checks rank 0's own sample alone (single-GPU path on the same data)."""
    lib = capi.lib()
    from bench_common import PF13
    n_s = max(50_000, 2_000_000 // world)
    n_s = min(n_s, creads.shape[0])
    sample = creads[:n_s].contiguous()
    out = {"reads_per_rank": n_s, "combined_equal": None, "rank0_alone_equal": None}
    try:
        ctx.check(lib.aix_count13_begin(ctx.handle))
        ctx.check(lib.aix_count13_add_dev(ctx.handle, sample.data_ptr(), sample.numel(), capi.FMT_PLAIN))
        if peer is not None:
            res = peer.reduce(stream)
            with torch.cuda.stream(stream):
                mine = res.clone()
        else:
            ctx.check(lib.aix_count13_flush(ctx.handle))
            with torch.cuda.stream(stream):
                dist.reduce_scatter_tensor(rs_out, hist_tensor, op=dist.ReduceOp.SUM)
                mine = rs_out.clone()
        ctx.sync()
        with torch.cuda.stream(stream):
            parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            reads_parts = [torch.empty_like(sample) for _ in range(world)] if rank == 0 else None
            dist.gather(sample, reads_parts, dst=0)
        ctx.sync()
        ok = torch.ones(2, device=dev, dtype=torch.int32)
        if rank == 0:
            have_ref = os.path.exists(PF13) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bin", "count_kmers13"))
            if have_ref:
                m13 = capi.Mphf.load(ctx, PF13)
                # combined direct-address histogram -> the library's buffer -> MPHF order (.tf.bin layout)
                ctx.check(lib.aix_count13_flush(ctx.handle))
                with torch.cuda.stream(stream):
                    hist_tensor.copy_(torch.cat(parts))
                    tf_dev = torch.zeros(1 << 26, device=dev, dtype=torch.int64)
                ctx.check(lib.aix_count13_finish_dev(ctx.handle, m13._h, 0, 1 << 26, tf_dev.data_ptr()))
                ctx.sync()
                tmpdir = tempfile.mkdtemp(prefix="aix_mg_", dir=_tmp_root())
                try:
                    all_reads = torch.cat(reads_parts).cpu().numpy()
                    cr = cpu_count_run(all_reads, os.cpu_count() or 1, tmpdir)
                    if cr:
                        ok[0] = 1 if np.array_equal(cr[3], tf_dev.cpu().numpy().view(np.uint64)) else 0
                        out["reference_seconds"] = cr[1]
                        cr1 = cpu_count_run(all_reads[:min(n_s, 250_000)], os.cpu_count() or 1, tmpdir)
                        tf_q, _ = ctx.count13(m13, all_reads[:min(n_s, 250_000)].reshape(-1), capi.FMT_PLAIN)
                        ok[1] = 1 if (cr1 and np.array_equal(cr1[3], tf_q)) else 0
                finally:
                    shutil.rmtree(tmpdir, ignore_errors=True)
                del m13
            else:
                ok[:] = -1
        with torch.cuda.stream(stream):
            dist.broadcast(ok, src=0)
        ctx.sync()
        v = [int(x) for x in ok.cpu()]
        out["combined_equal"] = None if v[0] < 0 else bool(v[0])
        out["rank0_alone_equal"] = None if v[1] < 0 else bool(v[1])
        out["total_reads"] = n_s * world
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        out["error"] = f"{type(e).__name__}: {e}"
    return out



def single_process_multi_legs(torch, capi, args, world, extra):
    """aix_count13_multi_dev and aix_positions_build23_multi on all `world` GPUs from this one process (rank 0), while the
    other ranks idle: (1) 13-mer counting, the same shards as the torchrun step (count + exchange, device resident), next
    to the torchrun number; (2) the C5 positions index built by all GPUs -- bit-equality with the single-GPU build is
    checked on a 5 M-read prefix (host arrays), the full 50 M-read build is timed with its phase times and the bytes that
    crossed NVLink."""
    import ctypes as C
    lib = capi.lib()
    out = {}
    mg = C.c_void_p()
    if lib.aix_multi_create(world, None, C.byref(mg)) != 0:
        raise RuntimeError((lib.aix_multi_last_error(None) or b"").decode())
    ctxs = []
    try:
        out["gpus"] = int(lib.aix_multi_size(mg))
        out["peer_access"] = bool(lib.aix_multi_peer_access(mg))
        for r in range(world):
            c = capi.Context.__new__(capi.Context)
            c._h = C.c_void_p(lib.aix_multi_ctx(mg, r))
            ctxs.append(c)
        # ---- (1) counting: one 25 M-read shard per GPU, generated on that GPU with the torchrun ranks' seeds
        if args.count_reads > 0 and os.path.exists(PF13_PATH):
            shards = []
            for r in range(world):
                d = torch.device("cuda", r)
                with torch.cuda.device(d):
                    shards.append(make_reads(torch, d, args.count_genome, args.count_reads, 150, 11, 12 + r).reshape(-1))
                    torch.cuda.synchronize(d)
            m13 = capi.Mphf.load(ctxs[0], PF13_PATH)
            ptrs = (C.c_void_p * world)(*[t.data_ptr() for t in shards])
            lens = (C.c_uint64 * world)(*[t.numel() for t in shards])
            n_kmers = args.count_reads * 138 * world

            def step(tf=None, st=None):
                rc = lib.aix_count13_multi_dev(mg, m13._h, ptrs, lens, capi.FMT_PLAIN, tf, st)
                if rc != 0:
                    raise RuntimeError((lib.aix_multi_last_error(mg) or b"").decode())

            for _ in range(2):
                step()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            dt = (time.perf_counter() - t0) / args.steps
            st = capi.CountStats()
            tf = np.zeros(1 << 26, dtype=np.uint64)
            t0 = time.perf_counter()
            step(tf.ctypes.data, C.byref(st))
            full_s = time.perf_counter() - t0
            torchrun = extra.get("count13", {}).get("value")
            out["count13"] = {"value": n_kmers / dt, "unit": "k-mers/s", "ms_per_step": dt * 1e3,
                              "what": "aix_count13_multi_dev: count + exchange over peer pointers, host wall clock around the call",
                              "torchrun_value": torchrun, "ratio_to_torchrun": (n_kmers / dt) / torchrun if torchrun else None,
                              "stats_ok": bool(st.valid == n_kmers and st.sequences == args.count_reads * world),
                              "tf_sum_ok": bool(int(tf.sum()) == n_kmers),
                              "with_permutation_and_download_ms": full_s * 1e3}
            del shards, tf
            lib.aix_mphf_destroy(ctxs[0]._h, m13._h)
            m13._h = C.c_void_p()
            for r in range(world):
                with torch.cuda.device(r):
                    torch.cuda.empty_cache()
        # ---- (2) positions index over all GPUs
        d0 = torch.device("cuda", 0)
        n_reads = args.multi_positions_reads
        with torch.cuda.device(d0):
            reads = make_reads(torch, d0, 250_000_000 if n_reads >= 10_000_000 else 5 * n_reads, n_reads, 150, 31, 32)
            n_bytes = reads.numel()
            pad = torch.full((64,), 10, device=d0, dtype=torch.uint8)
            reads = torch.cat([reads.reshape(-1), pad])
            stream0 = torch.cuda.ExternalStream(ctxs[0].stream, device=d0)
            torch.cuda.synchronize(d0)
            res = {}
            for tag, nr in (("check_5M_reads", min(n_reads, 5_000_000)), ("full", n_reads)):
                nb = nr * 151
                mphf, index, checker_t, tf_t, n_keys = build_index(torch, capi, ctxs[0], reads[:nb])
                ctxs[0].trim()  # the index build leaves ~100 GB of freed blocks in this ctx's pool: the exchange buffers need the room
                # replicate the index: the host arrays go to every other GPU
                info = mphf.info
                words, ranks = mphf.arrays()
                chk_h, tf_h = checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32)
                ms, ixs = [mphf], [index]
                for r in range(1, world):
                    m = capi.Mphf.from_arrays(ctxs[r], info["n"], info["hash_domain"], info["seed"], words, ranks)
                    ms.append(m)
                    ixs.append(capi.Index23.upload(ctxs[r], m, chk_h, tf_h))
                handles = (C.c_void_p * world)(*[ix._h for ix in ixs])
                r_host = ctxs[0].pinned((nb,), np.uint8)
                torch.from_numpy(r_host).copy_(reads[:nb])
                torch.cuda.synchronize(d0)
                st = capi.MultiBuildStats()
                gi = np.zeros(n_keys + 1, dtype=np.uint64)
                if tag == "full":
                    def build(pos_ptr=None):
                        rc = lib.aix_positions_build23_multi(mg, handles, r_host.ctypes.data, nb, gi.ctypes.data, pos_ptr, C.byref(st))
                        if rc != 0:
                            raise RuntimeError((lib.aix_multi_last_error(mg) or b"").decode())
                    build()  # warm-up (pools)
                    build()
                    dev_ms = st.emit_partition_ms + st.alloc_ms + st.exchange_ms + st.sort_finalize_ms
                    res[tag] = {"reads": nr, "index_keys": n_keys, "occurrences": int(st.positions), "stats": st.as_dict(),
                                "value": st.positions / (dev_ms / 1e3), "unit": "occurrences/s (emit + exchange-buffer cudaMalloc + exchange + sort "
                                "phases, max over GPUs; reads upload and positions download excluded)", "device_phases_ms": dev_ms,
                                "nvlink_bytes": int(st.peer_bytes), "single_gpu_ms": (extra.get("c5_positions") or {}).get("ms_per_step")}
                else:
                    gp = np.zeros(nr * 128, dtype=np.uint64)
                    rc = lib.aix_positions_build23_multi(mg, handles, r_host.ctypes.data, nb, gi.ctypes.data, gp.ctypes.data, C.byref(st))
                    if rc != 0:
                        raise RuntimeError((lib.aix_multi_last_error(mg) or b"").decode())
                    with torch.cuda.stream(stream0):
                        pos = capi.Positions.build_dev(index, reads.data_ptr(), nb, 23)
                        si, sp = pos.download()
                        pos.close()
                    res[tag] = {"reads": nr, "index_keys": n_keys, "indices_equal_single_gpu": bool(np.array_equal(gi, si)),
                                "positions_equal_single_gpu": bool(np.array_equal(gp, sp)), "stats": st.as_dict()}
                    del gp, si, sp
                for ix in ixs:
                    ix.close()
                for m in ms:
                    m.close()
                del r_host, mphf, index, checker_t, tf_t
                for c in ctxs:
                    c.trim()
                torch.cuda.empty_cache()
            out["positions_build"] = res
    finally:
        for c in ctxs:
            c._h = C.c_void_p()  # borrowed from the aix_multi
        lib.aix_multi_destroy(mg)
    return out



# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from aindex_b200 import capi

    rank, world, local = dist_setup(args)
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ctx = capi.Context(local)
    lib = capi.lib()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    t_setup = time.perf_counter()

    # ---- C2 setup: reads -> index ------------------------------------------------------------
    reads = make_reads(torch, dev, args.genome, args.reads, 150, 1, 2)
    torch.cuda.synchronize()
    t_idx = time.perf_counter()
    mphf, index, checker_t, tf_t, n_keys = build_index(torch, capi, ctx, reads)
    ctx.sync()
    index_build_s = time.perf_counter() - t_idx
    # Q2 (reported under extra): 50 % substrings of the reads (hits, either strand) + 50 % random
    q2_n = min(args.queries, 20_000_000)
    q2_dev = torch.cat([make_hit_queries(torch, dev, reads, q2_n // 2, 4 + rank), make_queries(torch, dev, q2_n - q2_n // 2, 5 + rank)])
    q2_dev = q2_dev[torch.randperm(q2_n, device=dev)].contiguous()
    del reads
    torch.cuda.empty_cache()
    q_dev = make_queries(torch, dev, args.queries, 3 + rank)
    out_dev = torch.empty(args.queries, device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_dev():
        index.query_dev(q_dev.data_ptr(), 23, None, args.queries, capi.Q_TF, out_dev.data_ptr())

    # ---- device-resident timing ----------------------------------------------------------------
    # clocks are sampled from before the warm-up until the last timed region (tf23 device pass,
    # e2e, 13-mer counting) has ended
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("AIX_BENCH_NO_SAMPLER"):  # (diagnosis only: a run without the sampler has no clocks record)
        sampler.start()
        time.sleep(0.3)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    ctx.sync()
    launches0 = ctx.launches
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record(stream)
    for a, b in evs:
        a.record(stream)
        step_dev()
        b.record(stream)
    e_all1.record(stream)
    ctx.sync()
    barrier()
    launches = ctx.launches - launches0
    total_ms = max_over_ranks(e_all0.elapsed_time(e_all1))
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = total_ms / args.steps
    value = world * args.queries / (ms_per_step / 1e3)
    hits = int((out_dev > 0).sum().item())
    filter_after_q1 = index.filter_stats  # which kernel the launcher chose for the timed batches, and the pass rate it saw
    # the same batches with the front filter switched off: the direct lookup kernel that hit-dominated batches run
    index.set_filter("off")
    for _ in range(3):
        step_dev()
    ctx.sync()
    da, db = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    da.record(stream)
    for _ in range(args.steps):
        step_dev()
    db.record(stream)
    ctx.sync()
    direct_ms = da.elapsed_time(db) / args.steps
    direct_equal = bool(int((out_dev > 0).sum().item()) == hits)
    index.set_filter("auto")

    # ---- Q2: 50 % hits (device resident) ---------------------------------------------------------
    q2_out = torch.empty(q2_n, device=dev, dtype=torch.int32)
    for _ in range(3):
        index.query_dev(q2_dev.data_ptr(), 23, None, q2_n, capi.Q_TF, q2_out.data_ptr())
    ctx.sync()
    qa, qb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qa.record(stream)
    for _ in range(args.steps):
        index.query_dev(q2_dev.data_ptr(), 23, None, q2_n, capi.Q_TF, q2_out.data_ptr())
    qb.record(stream)
    ctx.sync()
    q2_ms = qa.elapsed_time(qb) / args.steps
    q2 = {"queries": q2_n, "hit_fraction": float((q2_out > 0).float().mean().item()), "ms_per_step": q2_ms,
          "value": q2_n / (q2_ms / 1e3), "unit": "queries/s (per GPU, device resident)"}
    del q2_dev, q2_out

    # ---- e2e: host buffers through the C-ABI (pinned) -----------------------------------------
    e2e = None
    if not args.no_e2e:
        q_host = ctx.pinned((args.queries, 23), np.uint8)
        o_host = ctx.pinned((args.queries,), np.uint32)
        torch.from_numpy(q_host).copy_(q_dev)
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, 3))
        index.query(q_host, capi.Q_TF, out=o_host)  # warm-up: allocates the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            index.query(q_host, capi.Q_TF, out=o_host)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        same = bool(np.array_equal(o_host, out_dev.cpu().numpy().view(np.uint32)))
        e2e = {"value": world * args.queries / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(args.queries * 23 * world), "d2h_bytes_per_step": int(args.queries * 4 * world),
               "steps": e2e_steps, "host_memory": "pinned", "matches_device_path": same}
        # the packed form of the same call (PHASH_MAP::get_freq(uint64_t), hash.hpp:123-140): 8 B per query over PCIe
        pk = ctx.pinned((args.queries,), np.uint64)
        po = ctx.pinned((args.queries,), np.uint32)
        code = torch.zeros(256, device=dev, dtype=torch.int64)
        for i_c, ch in enumerate(b"ACGT"):
            code[ch] = i_c
        sh = (2 * (22 - torch.arange(23, device=dev, dtype=torch.int64)))
        for s0 in range(0, args.queries, 20_000_000):
            s1 = min(args.queries, s0 + 20_000_000)
            torch.from_numpy(pk[s0:s1].view(np.int64)).copy_((code[q_dev[s0:s1].long()] << sh[None, :]).sum(1))
        torch.cuda.synchronize()
        ctx.check(lib.aix_get_freq23(ctx.handle, index._h, pk.ctypes.data, args.queries, po.ctypes.data))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.check(lib.aix_get_freq23(ctx.handle, index._h, pk.ctypes.data, args.queries, po.ctypes.data))
        barrier()
        p_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        e2e_packed = {"value": world * args.queries / p_s, "unit": "queries/s", "ms_per_step": p_s * 1e3,
                      "h2d_bytes_per_step": int(args.queries * 8 * world), "d2h_bytes_per_step": int(args.queries * 4 * world),
                      "api": "aix_get_freq23 (packed uint64 k-mers)", "matches_string_path": bool(np.array_equal(po, o_host))}
        del pk
        # the 6-byte dna_bitset record form (aix_get_freq23_packed): 6 B per query over PCIe
        p6 = ctx.pinned((args.queries, 6), np.uint8)
        for s0 in range(0, args.queries, 20_000_000):
            s1 = min(args.queries, s0 + 20_000_000)
            c = torch.zeros((s1 - s0, 24), device=dev, dtype=torch.uint8)
            c[:, :23] = code[q_dev[s0:s1].long()].to(torch.uint8)
            c = c.reshape(-1, 6, 4)
            torch.from_numpy(p6[s0:s1]).copy_((c[:, :, 0] << 6) | (c[:, :, 1] << 4) | (c[:, :, 2] << 2) | c[:, :, 3])
            del c
        torch.cuda.synchronize()
        ctx.check(lib.aix_get_freq23_packed(ctx.handle, index._h, p6.ctypes.data, args.queries, po.ctypes.data))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.check(lib.aix_get_freq23_packed(ctx.handle, index._h, p6.ctypes.data, args.queries, po.ctypes.data))
        barrier()
        p6_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        e2e_packed6 = {"value": world * args.queries / p6_s, "unit": "queries/s", "ms_per_step": p6_s * 1e3,
                       "h2d_bytes_per_step": int(args.queries * 6 * world), "d2h_bytes_per_step": int(args.queries * 4 * world),
                       "api": "aix_get_freq23_packed (6-byte dna_bitset records)", "matches_string_path": bool(np.array_equal(po, o_host))}
        # what the bus can do: one pinned host -> device copy of the query buffer, timed alone (per rank, then summed)
        stage = torch.empty(args.queries * 23, device=dev, dtype=torch.uint8)
        hq = torch.from_numpy(q_host.reshape(-1))
        stage.copy_(hq, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        stage.copy_(hq, non_blocking=True)
        torch.cuda.synchronize()
        h2d_s = max_over_ranks(time.perf_counter() - t0)
        h2d_peak = world * args.queries * 23 / h2d_s / 1e9
        for blk in (e2e, e2e_packed, e2e_packed6):
            gbs = (blk["h2d_bytes_per_step"] + blk["d2h_bytes_per_step"]) / (blk["ms_per_step"] / 1e3) / 1e9
            blk["pcie"] = {"h2d_plus_d2h_gbs": gbs, "pinned_h2d_copy_peak_gbs": h2d_peak, "utilisation_vs_h2d_peak": gbs / h2d_peak,
                           "note": "both directions are counted against a one-direction copy peak (full duplex): > 1 is possible; "
                                   "the e2e rate is a bus number, the kernel takes %.2f ms of the step" % float(np.mean(kernel_ms))}
        del p6, po, stage, hq
    else:
        q_host = None
        e2e_packed = None
        e2e_packed6 = None

    # ---- parity legs that need the queries still in HBM (untimed except where a value is reported) ---------
    ref_built = reference_built_index_check(torch, capi, ctx, stream, args, q_dev, out_dev, n_keys) if rank == 0 else None
    sharded = sharded_index_check(torch, dist, capi, ctx, stream, dev, args, rank, world, mphf, checker_t, tf_t, n_keys,
                                  q_dev, out_dev, barrier, max_over_ranks) if world > 1 else None

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = float(np.mean(kernel_ms))
    achieved = args.queries * Q1_BYTES_PER_QUERY / (k_ms / 1e3) / 1e9
    filtered = filter_after_q1["batches_filter"] > 0
    k_name = ("tf23_filter3_kernel (front Bloom filter, %d B, its lines held in persisting L2; the %.1f %% of the queries that pass it go through the lookup of tf23_stream_kernel)"
              % (filter_after_q1["filter_bytes"], 100.0 * (filter_after_q1["pass_rate"] or 0.0))) if filtered else \
        "tf23_stream_kernel<AIX_Q_TF, canonical> (" + index.layout["records"] + " MPHF records)"
    roofline = {"bound": "hbm", "kernel": k_name, "achieved": achieved,
                "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "bytes_per_unit": Q1_BYTES_PER_QUERY, "units_per_launch": args.queries, "kernel_ms": k_ms,
                "achieved_one_probe_bytes": args.queries * Q1_BYTES_ONE_PROBE / (k_ms / 1e3) / 1e9,
                "traffic": ncu_traffic("tf23_filter_kernel_q1_100M" if filtered else "tf23_stream_kernel_q1_100M", args.queries),
                "traffic_frac": None,  # filled below: the DRAM bytes really moved per second against the same peak
                "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture)",
                # the other denominators of this kernel (profiles/r01_atomic_roofline.txt):
                # random 16-byte gathers from a 1 GiB table run at 50.1 G/s on this GPU (the index is 0.8 GB)
                # what actually bounds the kernel: three scattered L2 requests per query (ncu: L1TEX/L2 request path, not DRAM)
                # the direct kernel (front filter off): three scattered L2 requests per query bound it
                "direct_kernel": {"kernel": "tf23_stream_kernel<AIX_Q_TF, canonical> (" + index.layout["records"] + " MPHF records), front filter off",
                                  "ms_per_step": direct_ms, "achieved_gq_s": args.queries / (direct_ms / 1e3) / 1e9,
                                  "same_hit_count": direct_equal,
                                  "request_only_kernel_gq_s": {"16MiB_table": LOOKUP_SHAPE_16MIB_G, "64MiB_table": LOOKUP_SHAPE_64MIB_G},
                                  "frac_of_16MiB_shape": args.queries / (direct_ms / 1e3) / 1e9 / LOOKUP_SHAPE_16MIB_G,
                                  "frac_of_64MiB_shape": args.queries / (direct_ms / 1e3) / 1e9 / LOOKUP_SHAPE_64MIB_G,
                                  "source": "profiles/r02_atomic_roofline.txt (lookup_shape_stream recs16=3 bytes1=0)",
                                  "note": "61.5 MB structure, 345 instructions per query on top of the requests: between the two request-only "
                                          "shapes, i.e. at the L2-request ceiling of a 3-vertex MPHF lookup; the front filter replaces the "
                                          "three requests by one for absent k-mers"},
                "random_access": {"achieved_gq_s": args.queries / (k_ms / 1e3) / 1e9, "peak_ggathers_s": GATHER_PEAK_G,
                                  "frac": args.queries / (k_ms / 1e3) / 1e9 / GATHER_PEAK_G,
                                  "note": "measured random-gather rate (informative: the front filter answers absent k-mers "
                                          "from one 8-byte word, most of them without touching a record)"}}
    if roofline["traffic"]:
        roofline["traffic_gbs"] = roofline["traffic"] / (k_ms / 1e3) / 1e9
        roofline["traffic_frac"] = roofline["traffic_gbs"] / peak_gbs

    # ---- 13-mer counting (second half of the metric), per-GPU shard of C3 -------------------------
    extra = {"index": {"keys": n_keys, "build_s": index_build_s, "hit_fraction": hits / args.queries,
                       "canonical_only": index.info["canonical_only"]},
             "tf23_q2_half_hits": q2, "tf23_packed_e2e": e2e_packed, "tf23_packed6_e2e": e2e_packed6, "setup_s": setup_s,
             "reference_built_index": ref_built, "sharded_index23": sharded, "front_filter": filter_after_q1}
    creads = None
    if args.count_reads > 0:
        del q_dev
        torch.cuda.empty_cache()
        creads = make_reads(torch, dev, args.count_genome, args.count_reads, 150, 11, 12 + rank)
        n_bytes = creads.numel()
        n_kmers = args.count_reads * 138

        def count_step():
            ctx.check(lib.aix_count13_begin(ctx.handle))
            ctx.check(lib.aix_count13_add_dev(ctx.handle, creads.data_ptr(), n_bytes, capi.FMT_PLAIN))
            if peer is not None:
                # widen + reduce-scatter fused into one kernel over NVLink peer memory (aindex_b200/dist.py)
                peer.reduce(stream)
                return
            ctx.check(lib.aix_count13_flush(ctx.handle))
            if world > 1:
                # per-GPU 4^13 histograms -> NCCL reduce-scatter over k-mer ranges (rank r owns
                # [r, r+1) * 4^13 / world); torch wraps the library's device buffer (no copy) and
                # the collective is ordered on the library's stream
                with torch.cuda.stream(stream):
                    dist.reduce_scatter_tensor(rs_out, hist_tensor, op=dist.ReduceOp.SUM)

        peer = None
        if world > 1 and os.environ.get("AIX_COUNT13_COLLECTIVE", "peer") == "peer":
            from aindex_b200 import dist as D
            try:
                peer = D.PeerHistogram(ctx)
            except Exception as e:  # IPC not available between these processes: NCCL reduce-scatter instead
                sys.stderr.write(f"[bench] peer-memory combine unavailable ({e}); using NCCL reduce-scatter\n")
                peer = None
        if world > 1:
            import ctypes as C
            ctx.check(lib.aix_count13_begin(ctx.handle))
            ptr = lib.aix_count13_hist_dev(ctx.handle)
            hist_tensor = _wrap_device_i64(torch, ptr, 1 << 26, dev)
            rs_out = torch.empty((1 << 26) // world, device=dev, dtype=torch.int64)
        for _ in range(max(1, args.warmup)):
            count_step()
        ctx.sync()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        t0 = time.perf_counter()
        l0 = ctx.launches
        for _ in range(args.steps):
            count_step()
        count_launches = (ctx.launches - l0) / args.steps
        c1.record(stream)
        ctx.sync()
        barrier()
        wall = time.perf_counter() - t0
        c_ms = max_over_ranks(c0.elapsed_time(c1)) / args.steps
        extra_wall_ms = wall * 1e3 / args.steps
        st = capi.CountStats()
        ctx.check(lib.aix_count13_stats(ctx.handle, st))
        peer_equal = None
        if peer is not None:
            # untimed: the fused combine must give exactly the slice the NCCL reduce-scatter gives
            with torch.cuda.stream(stream):
                mine = peer.out.clone()
            ctx.check(lib.aix_count13_flush(ctx.handle))
            with torch.cuda.stream(stream):
                dist.reduce_scatter_tensor(rs_out, hist_tensor, op=dist.ReduceOp.SUM)
                same = torch.tensor([1 if torch.equal(mine, rs_out) else 0], device=dev, dtype=torch.int32)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
            ctx.sync()
            peer_equal = bool(int(same.item()))
        ok = st.valid == n_kmers and st.sequences == args.count_reads
        kps = world * n_kmers / (c_ms / 1e3)
        extra["count13"] = {"metric": "13-mer k-mers counted/s", "value": kps, "unit": "k-mers/s", "ms_per_step": c_ms,
                            "reads_per_gpu": args.count_reads, "kmers_per_gpu": n_kmers, "stats_ok": bool(ok),
                            "collective": (None if world == 1 else
                                           "fused widen + gather-reduce over NVLink peer memory (CUDA IPC), NCCL barriers" if peer is not None
                                           else "nccl reduce_scatter(sum) of the 4^13 u64 histogram"),
                            "hbm_frac_9.09B": kps / world * C3_BYTES_PER_KMER / 1e9 / peak_gbs,
                            "atomic_roofline": {"achieved_gred_s": kps / world / 1e9,
                                                "frac_of_256MiB_table_red_rate": kps / world / 1e9 / RED_PEAK_256M_G,
                                                "frac_of_L2_resident_red_rate": kps / world / 1e9 / RED_PEAK_L2_G,
                                                "peaks_gred_s": {"256MiB_table": RED_PEAK_256M_G, "64MiB_slice": RED_PEAK_L2_G},
                                                "source": "profiles/r01_atomic_roofline.txt"},
                            "gpu_launches_per_step": int(count_launches), "fused_combine_equals_nccl_reduce_scatter": peer_equal}
        extra["count13"]["wall_ms_per_step"] = extra_wall_ms
        if not args.no_e2e:
            # e2e: the shard starts in pinned host memory; H2D chunks overlap the count kernel
            r_host = ctx.pinned((n_bytes,), np.uint8)
            torch.from_numpy(r_host).copy_(creads.reshape(-1))
            torch.cuda.synchronize()

            def count_step_host():
                ctx.check(lib.aix_count13_begin(ctx.handle))
                ctx.check(lib.aix_count13_add(ctx.handle, r_host.ctypes.data, n_bytes, capi.FMT_PLAIN))
                ctx.check(lib.aix_count13_stats(ctx.handle, st))  # D2H read of the step's result

            count_step_host()
            barrier()
            t0 = time.perf_counter()
            n_e = max(1, min(args.steps, 3))
            for _ in range(n_e):
                count_step_host()
            barrier()
            e_s = max_over_ranks(time.perf_counter() - t0) / n_e
            extra["count13"]["e2e"] = {"value": world * n_kmers / e_s, "unit": "k-mers/s", "ms_per_step": e_s * 1e3,
                                       "h2d_bytes_per_step": int(n_bytes * world), "d2h_bytes_per_step": 32 * world,
                                       "stats_ok": bool(st.valid == n_kmers)}
            del r_host
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            # the reference binary on the first 1 M reads of the shard, all host threads; results compared
            n_s = min(args.count_reads, 1_000_000)
            tmpdir = tempfile.mkdtemp(prefix="aix_bench_", dir=_tmp_root())
            try:
                sample_np = creads[:n_s].cpu().numpy()
                cr = cpu_count_run(sample_np, os.cpu_count() or 1, tmpdir)
                if cr:
                    kind_c, secs_c, nk, tf_ref, pf13 = cr
                    m13 = capi.Mphf.load(ctx, pf13)
                    tf_gpu, _ = ctx.count13(m13, sample_np.reshape(-1), capi.FMT_PLAIN)
                    extra["count13"]["cpu_baseline"] = {"value": nk / secs_c, "unit": "k-mers/s", "cores": os.cpu_count() or 1,
                                                        "kind": kind_c, "sample": f"first {n_s} reads of the shard (count_kmers13, all host threads)",
                                                        "seconds": secs_c, "results_equal_gpu": bool(np.array_equal(tf_ref, tf_gpu))}
                    del m13
            finally:
                shutil.rmtree(tmpdir, ignore_errors=True)
        if world > 1:
            extra["count13"]["multi_gpu_equals_reference"] = count13_multi_gpu_vs_reference(
                torch, dist, capi, ctx, stream, dev, rank, world, creads, peer, hist_tensor, rs_out)
        ctx.check(lib.aix_count13_end(ctx.handle))
    del creads
    torch.cuda.empty_cache()

    # ---- CPU baseline on the same box (rank 0, N = 1 only) -------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        tmpdir = tempfile.mkdtemp(prefix="aix_bench_", dir=_tmp_root())
        try:
            prefix = write_index_files(tmpdir, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
            sample = args.cpu_sample or min(args.queries, 2_000_000 * threads)  # ~1.5 s per pass
            if q_host is None:
                raise RuntimeError("cpu baseline needs the host copy of the queries")
            qs = np.ascontiguousarray(q_host[:sample])
            kind, secs, res = cpu_query_runs(prefix, qs, threads, 3)
            gpu_res = o_host[:sample]
            cpu_baseline = {"value": sample / min(secs), "unit": "queries/s", "cores": threads, "kind": kind,
                            "sample": f"first {sample} of the {args.queries} Q1 queries, best of {len(secs)} passes, "
                                      f"{threads} std::threads over PHASH_MAP::get_freq",
                            "seconds": min(secs), "results_equal_gpu": bool(np.array_equal(res, gpu_res))}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
    q_host = o_host = None

    # ---- the other BASELINE configs (C4 coverage on this index, C1 all-4^13 tf query, C5 positions index), N = 1 ---
    if rank == 0 and world == 1 and args.configs:
        import types
        import bench_configs
        torch.cuda.set_stream(stream)  # torch work of the config runs and the library's kernels share one stream
        cargs = types.SimpleNamespace(scale=args.config_scale, checks=False, e2e=not args.no_e2e, cpu=not args.no_cpu_baseline)
        names = [c.strip().lower() for c in args.configs.split(",") if c.strip()]
        if "c4" in names:
            extra["c4_coverage"] = _guard(lambda: bench_configs.run_c4(ctx, stream, dev, cargs,
                                                                      None if args.config_scale != 1.0 else (mphf, index, checker_t, tf_t, n_keys)))
        del index, mphf, checker_t, tf_t
        torch.cuda.empty_cache()
        if "c1" in names:
            extra["c1_tf13_all"] = _guard(lambda: bench_configs.run_c1(ctx, stream, dev, cargs))
            torch.cuda.empty_cache()
        if "c5" in names:
            extra["c5_positions"] = _guard(lambda: bench_configs.run_c5(ctx, stream, dev, cargs))
            torch.cuda.empty_cache()
        if "k1" in names:  # the encoding kernels on their own (no BASELINE config)
            cargs.checks = True
            extra["k1_codec"] = _guard(lambda: bench_configs.run_k1(ctx, stream, dev, cargs))
            torch.cuda.empty_cache()
    # ---- N > 1: the same multi-GPU work from ONE process through the C-ABI (aix_multi: one context + one host thread per
    #      GPU, no torch.distributed on the data path); the other ranks release their memory and wait
    if world > 1 and args.multi_legs:
        del index, mphf, checker_t, tf_t
        torch.cuda.empty_cache()
        ctx.trim()
        barrier()
        # the idle ranks wait on the rendezvous store, not in an NCCL kernel that would spin on the GPUs rank 0 is using
        import datetime
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            extra["single_process_multi_gpu"] = _guard(lambda: single_process_multi_legs(torch, capi, args, world, extra))
            store.set("aix_multi_legs_done", "1")
        else:
            store.wait(["aix_multi_legs_done"], datetime.timedelta(minutes=30))
        barrier()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        line = {
            "metric": "23-mer batch tf queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "vs_baseline_note": "BASELINE.json.published is empty; the reference README quotes 2.3 M q/s for 1 M random queries through "
                                "the Python list API on unstated hardware (README.md:14) -- another config, so no ratio is formed; "
                                "the reference is timed on this box instead (cpu_baseline, --impl reference, profiles/r01_api_path.json)",
            "dtype": "u64", "data": "synthetic",
            "config": c2_config(args, world, n_keys),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "extra": extra,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path, all host threads, same config: the same synthetic reads,
    the index built from them by the reference's own tools, the same 100 M Q1 queries per step through
    PHASH_MAP::get_freq.  Nothing of the product is imported or executed in this arm: torch (data generation only),
    the oracle's canonical 23-mer counter (the counting stage in front of the index build) and oracle/_ref binaries."""
    rank, world, local = dist_setup(args)
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import torch
    harness = ref_harness_path()
    if harness is None:
        _emit({"impl": "reference", "unavailable": "oracle/_ref/bin/ref_harness missing (the reference was not compiled into oracle/_ref)"})
        return
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    cache = ref_index_cache_dir(args)
    t0 = time.perf_counter()
    if not os.path.exists(os.path.join(cache, "meta.json")):
        reads = make_reads(torch, dev, args.genome, args.reads, 150, 1, 2)  # the generator and seeds of our arm
        reads_np = reads.cpu().numpy().reshape(-1)
        del reads
        meta = build_reference_index_c2(args, reads_np, cache)
        del reads_np
    else:
        meta = json.load(open(os.path.join(cache, "meta.json")))
    if meta is None:
        _emit({"impl": "reference", "unavailable": "reference tools missing under oracle/_ref/bin"})
        return
    prefix = os.path.join(cache, "c2.23")
    n_q = args.cpu_sample or args.queries  # the whole batch of our arm per step (same_config)
    q = make_queries(torch, dev, args.queries, 3)[:n_q].cpu().numpy()
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    setup_s = time.perf_counter() - t0
    qdir = tempfile.mkdtemp(prefix="aix_ref_", dir=_tmp_root(4 << 30))
    try:
        qf, of = os.path.join(qdir, "q1.bin"), os.path.join(qdir, "out.bin")
        q.tofile(qf)
        del q
        r = subprocess.run([harness, "tf23", prefix + ".pf", prefix + ".tf.bin", prefix + ".kmers.bin", qf, str(n_q),
                            str(threads), of, str(args.warmup + args.steps)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        if r.returncode != 0:
            _emit({"impl": "reference", "unavailable": f"ref_harness failed with exit code {r.returncode}"})
            return
        secs = [float(l.split()[0].split("=")[1]) for l in r.stdout.splitlines() if l.startswith("seconds=")]
        hits = int(np.count_nonzero(np.fromfile(of, dtype=np.uint32)))
    finally:
        shutil.rmtree(qdir, ignore_errors=True)
    timed = secs[args.warmup:] if len(secs) > args.warmup else secs
    s_per_step = float(np.mean(timed))
    v = n_q / s_per_step
    line = {"impl": "reference", "metric": "23-mer batch tf queries/s", "value": v, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": c2_config(args, max(1, args.gpus), meta["index_keys"]),
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "reference",
                             "sample": f"{n_q} Q1 queries per step (the whole batch of one GPU of our arm), {threads} std::threads over "
                                       f"PHASH_MAP::get_freq (src/hash.hpp:123-140); the rate of this one host does not depend on --gpus"},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": {"index": {"keys": meta["index_keys"], "hit_fraction": hits / n_q, **{k: meta[k] for k in meta if k.endswith("_s")},
                                "built_by": meta.get("built_by")},
                      "setup_s": setup_s, "data_generated_on": dev.type}}
    _emit(line)


def _emit(line: dict):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


_REAL_STDOUT = None

if __name__ == "__main__":
    # libraries (NCCL's version banner, torchrun) write to stdout: keep fd 1 for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
