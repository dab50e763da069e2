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


def _ncu_traffic(key, units, units_profiled):
    """DRAM bytes per launch from the committed ncu capture, scaled if the run uses another size."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        return d["bytes_per_launch"] * (units / units_profiled)
    except Exception:
        return None


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
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic data (torch on the GPU: setup only, never timed)
# ------------------------------------------------------------------------------------------
def make_reads(torch, dev, genome_len, n_reads, read_len, seed_genome, seed_reads):
    """uint8[n_reads, read_len+1] plain-text reads ('\\n' terminated), uniform start, strand
    flipped with p = 0.5, no errors."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed_genome)
    genome = torch.randint(0, 4, (genome_len,), generator=g, device=dev, dtype=torch.uint8)
    g.manual_seed(seed_reads)
    lut = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    out = torch.empty((n_reads, read_len + 1), device=dev, dtype=torch.uint8)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    chunk = 2_000_000
    for s in range(0, n_reads, chunk):
        e = min(n_reads, s + chunk)
        start = torch.randint(0, genome_len - read_len, (e - s,), generator=g, device=dev, dtype=torch.int64)
        flip = torch.rand((e - s,), generator=g, device=dev) < 0.5
        codes = genome[start[:, None] + ar[None, :]]
        rc = (3 - codes).flip(1)
        codes = torch.where(flip[:, None], rc, codes)
        out[s:e, :read_len] = lut[codes.long()]
    out[:, read_len] = 10
    return out


def make_queries(torch, dev, n, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    out = torch.empty((n, 23), device=dev, dtype=torch.uint8)
    chunk = 20_000_000
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        out[s:e] = lut[torch.randint(0, 4, (e - s, 23), generator=g, device=dev, dtype=torch.uint8).long()]
    return out


def make_hit_queries(torch, dev, reads, n, seed):
    """Q2 half: 23-byte substrings of the reads at random offsets, random strand."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_reads, width = reads.shape
    out = torch.empty((n, 23), device=dev, dtype=torch.uint8)
    ar = torch.arange(23, device=dev, dtype=torch.int64)
    comp = torch.zeros(256, device=dev, dtype=torch.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    chunk = 10_000_000
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        r = torch.randint(0, n_reads, (e - s,), generator=g, device=dev, dtype=torch.int64)
        o = torch.randint(0, width - 1 - 23 + 1, (e - s,), generator=g, device=dev, dtype=torch.int64)
        sub = reads.reshape(-1)[(r * width + o)[:, None] + ar[None, :]]
        flip = torch.rand((e - s,), generator=g, device=dev) < 0.5
        out[s:e] = torch.where(flip[:, None], comp[sub.long()].flip(1), sub)
    return out


def build_index(torch, capi, ctx, reads):
    """reads (device tensor) -> canonical 23-mer table -> GPU MPHF -> {checker, tf} fill.
    Returns (mphf, index, checker_dev, tf_dev, n)."""
    import ctypes as C
    lib = capi.lib()
    n = C.c_uint64()
    ctx.check(lib.aix_canonical23_count_dev(ctx.handle, reads.data_ptr(), reads.numel(), C.byref(n)))
    kp, cp = C.c_void_p(), C.c_void_p()
    ctx.check(lib.aix_canonical23_result_dev(ctx.handle, C.byref(kp), C.byref(cp), None))
    n = int(n.value)
    mphf = capi.Mphf.build_dev(ctx, kp.value, n, 23)
    checker = torch.empty(n, device=reads.device, dtype=torch.int64)
    tf = torch.empty(n, device=reads.device, dtype=torch.int32)
    ctx.check(lib.aix_index23_fill_dev(ctx.handle, mphf._h, kp.value, cp.value, n, checker.data_ptr(), tf.data_ptr()))
    index = capi.Index23.upload_dev(ctx, mphf, checker.data_ptr(), tf.data_ptr(), n)
    return mphf, index, checker, tf, n


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="aix_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------
def _tmp_root(min_free: int = 8 << 30):
    """/dev/shm when it has room for the baseline's files (queries 2.3 GB + index 0.6 GB), else the default tmp dir"""
    try:
        if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free >= min_free:
            return "/dev/shm"
    except OSError:
        pass
    return None


def ref_harness_path():
    p = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")
    return p if os.path.exists(p) else None


def write_index_files(tmpdir, mphf, checker_np, tf_np):
    prefix = os.path.join(tmpdir, "c2.23")
    mphf.save(prefix + ".pf")
    checker_np.tofile(prefix + ".kmers.bin")
    tf_np.tofile(prefix + ".tf.bin")
    return prefix


def cpu_query_runs(prefix, queries_np, threads, reps, mphf_info=None, checker_np=None, tf_np=None):
    """Run the reference CPU path over `queries_np` (uint8[q,23]) `reps` times.
    -> (kind, [seconds per rep], results uint32[q])."""
    q = queries_np.shape[0]
    h = ref_harness_path()
    if h:
        qf, of = prefix + ".queries.bin", prefix + ".out.bin"
        queries_np.tofile(qf)
        r = subprocess.run([h, "tf23", prefix + ".pf", prefix + ".tf.bin", prefix + ".kmers.bin", qf, str(q),
                            str(threads), of, str(reps)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        if r.returncode == 0:
            secs = [float(l.split()[0].split("=")[1]) for l in r.stdout.splitlines() if l.startswith("seconds=")]
            return "reference", secs, np.fromfile(of, dtype=np.uint32)
    # the reference was not compiled: time the C oracle port (oracle/aindex_oracle.c, OpenMP)
    from oracle import oracle as O
    oix = O.Index23.load_prefix(prefix)
    secs, res = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = oix.batch(queries_np, None, O.MODE_TF, threads=threads)
        secs.append(time.perf_counter() - t0)
    return "port", secs, res


def cpu_count_run(reads_np, threads, tmpdir):
    """count_kmers13 of the reference on a plain reads sample -> (kind, seconds, k-mers, tf array)."""
    binp = os.path.join(ROOT, "oracle", "_ref", "bin", "count_kmers13")
    pf = os.path.join(ROOT, "oracle", "_ref", "data", "all_13mers.pf")
    n_kmers = (reads_np.shape[1] - 1 - 12) * reads_np.shape[0]
    if os.path.exists(binp) and os.path.exists(pf):
        rp, op = os.path.join(tmpdir, "c3.reads"), os.path.join(tmpdir, "c3.tf.bin")
        reads_np.tofile(rp)
        t0 = time.perf_counter()
        r = subprocess.run([binp, rp, pf, op, str(threads)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        wall = time.perf_counter() - t0
        if r.returncode == 0:
            ms = [l for l in r.stdout.splitlines() if l.startswith("Processing completed in")]
            secs = float(ms[0].split()[3]) / 1e3 if ms else wall
            tf = np.fromfile(op, dtype=np.uint64)
            os.unlink(op)
            os.unlink(rp)
            return "reference", secs, n_kmers, tf, pf
    return None


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
    if rank == 0:
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
        del pk, po
    else:
        q_host = None
        e2e_packed = None

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = float(np.mean(kernel_ms))
    achieved = args.queries * Q1_BYTES_PER_QUERY / (k_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "tf23_stream_kernel<AIX_Q_TF, canonical>", "achieved": achieved,
                "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "bytes_per_unit": Q1_BYTES_PER_QUERY, "units_per_launch": args.queries, "kernel_ms": k_ms,
                "achieved_one_probe_bytes": args.queries * Q1_BYTES_ONE_PROBE / (k_ms / 1e3) / 1e9,
                "traffic": _ncu_traffic("tf23_fixed_kernel_q1_100M", args.queries, 100_000_000),
                "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture)",
                # the other denominators of this kernel (profiles/r01_atomic_roofline.txt):
                # random 16-byte gathers from a 1 GiB table run at 50.1 G/s on this GPU (the index is 0.8 GB)
                "random_access": {"achieved_gq_s": args.queries / (k_ms / 1e3) / 1e9, "peak_ggathers_s": GATHER_PEAK_G,
                                  "frac": args.queries / (k_ms / 1e3) / 1e9 / GATHER_PEAK_G,
                                  "note": "measured random-gather rate; above 1.0 is possible because the L2-resident "
                                          "fingerprint tier answers misses without touching the HBM record"}}

    # ---- 13-mer counting (second half of the metric), per-GPU shard of C3 -------------------------
    extra = {"index": {"keys": n_keys, "build_s": index_build_s, "hit_fraction": hits / args.queries,
                       "canonical_only": index.info["canonical_only"]},
             "tf23_q2_half_hits": q2, "tf23_packed_e2e": e2e_packed, "setup_s": setup_s}
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
        ctx.check(lib.aix_count13_end(ctx.handle))

    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline on the same box (rank 0, N = 1 only) -------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        tmpdir = tempfile.mkdtemp(prefix="aix_bench_", dir=_tmp_root())
        try:
            prefix = write_index_files(tmpdir, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
            sample = args.cpu_sample or min(args.queries, 6_250_000 * threads)  # 16 threads: the whole 100 M batch (~4 s per pass)
            if q_host is None:
                raise RuntimeError("cpu baseline needs the host copy of the queries")
            qs = np.ascontiguousarray(q_host[:sample])
            kind, secs, res = cpu_query_runs(prefix, qs, threads, 2)
            gpu_res = o_host[:sample]
            cpu_baseline = {"value": sample / min(secs), "unit": "queries/s", "cores": threads, "kind": kind,
                            "sample": f"first {sample} of the {args.queries} Q1 queries, best of {len(secs)} passes, "
                                      f"{threads} std::threads over PHASH_MAP::get_freq",
                            "seconds": min(secs), "results_equal_gpu": bool(np.array_equal(res, gpu_res))}
            if creads is not None:
                n_s = min(args.count_reads, 1_000_000)
                cr = cpu_count_run(creads[:n_s].cpu().numpy(), threads, tmpdir)
                if cr:
                    kind_c, secs_c, nk, tf_ref, pf13 = cr
                    m13 = capi.Mphf.load(ctx, pf13)
                    tf_gpu, _ = ctx.count13(m13, creads[:n_s].cpu().numpy().reshape(-1), capi.FMT_PLAIN)
                    extra["count13"]["cpu_baseline"] = {"value": nk / secs_c, "unit": "k-mers/s", "cores": threads,
                                                        "kind": kind_c, "sample": f"first {n_s} reads of the shard (count_kmers13, {threads} threads)",
                                                        "seconds": secs_c, "results_equal_gpu": bool(np.array_equal(tf_ref, tf_gpu))}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)

    if rank == 0:
        line = {
            "metric": "23-mer batch tf queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "vs_baseline_note": "BASELINE.json.published is empty; the reference README quotes 2.3 M q/s for 1 M random queries through "
                                "the Python list API on unstated hardware (README.md:14) -- another config, so no ratio is formed; "
                                "the reference is timed on this box instead (cpu_baseline, --impl reference, profiles/r01_api_path.json)",
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C2: 23-mer index over 10M synthetic 150bp reads; 100M random batch tf queries (Q1, ~100% miss) per GPU",
                       "reads": args.reads, "genome_bp": args.genome, "queries_per_gpu": args.queries, "index_keys": n_keys,
                       "parallelism": f"replicated index, queries sharded x{world}",
                       "l2": "inputs larger than L2 (2.3 GB of queries, 0.8 GB index per pass); no flush needed"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "extra": extra,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _wrap_device_i64(torch, ptr, n, dev):
    """torch view of a device buffer owned by libaindex_cuda (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=dev)


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path, all host threads, same config."""
    rank, world, local = dist_setup(args)
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import torch
    from aindex_b200 import capi
    if not torch.cuda.is_available():
        _emit({"impl": "reference", "unavailable": "index setup for config C2 needs the GPU builder (no GPU visible)"})
        return
    dev = torch.device("cuda", 0)
    ctx = capi.Context(0)
    # setup only (untimed): the same index and the same Q1 queries as our arm
    reads = make_reads(torch, dev, args.genome, args.reads, 150, 1, 2)
    mphf, index, checker_t, tf_t, n_keys = build_index(torch, capi, ctx, reads)
    del reads
    sample = args.cpu_sample or min(args.queries, 1_000_000 * threads)
    q = make_queries(torch, dev, args.queries, 3)[:sample].cpu().numpy()
    tmpdir = tempfile.mkdtemp(prefix="aix_ref_", dir=_tmp_root())
    try:
        prefix = write_index_files(tmpdir, mphf, checker_t.cpu().numpy().view(np.uint64), tf_t.cpu().numpy().view(np.uint32))
        del index, mphf, checker_t, tf_t
        ctx.close()
        torch.cuda.empty_cache()
        kind, secs, _ = cpu_query_runs(prefix, q, threads, args.warmup + args.steps)
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    timed = secs[args.warmup:] if len(secs) > args.warmup else secs
    s_per_step = float(np.mean(timed))
    v = sample / s_per_step
    line = {"impl": "reference", "metric": "23-mer batch tf queries/s", "value": v, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C2: 23-mer index over 10M synthetic 150bp reads; random batch tf queries (Q1, ~100% miss)",
                       "reads": args.reads, "genome_bp": args.genome, "index_keys": n_keys,
                       "queries_per_step": sample, "note": "bounded sample of the 100M-query batch per step; CPU only"},
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": kind,
                             "sample": f"{sample} Q1 queries per step, {threads} std::threads over PHASH_MAP::get_freq (src/hash.hpp:123-140)"},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
