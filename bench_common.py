"""Shared pieces of bench.py / bench_configs.py / the full-size GPU tests: the seeded synthetic-data
generators of SURVEY 8(d) (torch on the GPU: setup only, never timed), the index builders, the clock
sampler and the CPU legs (the compiled reference under oracle/_ref, or the C oracle port)."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
PF13 = os.path.join(ROOT, "oracle", "_ref", "data", "all_13mers.pf")


def peak_hbm_gbs():
    """(GB/s, source): the driver-measured copy bandwidth of this pool's B200s, else the profiling guide's fallback"""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def ncu_traffic(key, units):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), scaled to `units`."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        return d["bytes_per_launch"] * (units / d.get("units_profiled", units))
    except Exception:
        return None


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
    Returns (mphf, index, checker_dev, tf_dev, n).  AIX_TRACE=1 prints the phase times on stderr."""
    import ctypes as C
    lib = capi.lib()
    trace = bool(os.environ.get("AIX_TRACE"))
    t = [time.perf_counter()]

    def mark(what):
        if trace:
            ctx.sync()
            t.append(time.perf_counter())
            sys.stderr.write(f"[aix trace] build_index: {what:<60s} {(t[-1] - t[-2]) * 1e3:9.3f} ms\n")

    n = C.c_uint64()
    ctx.check(lib.aix_canonical23_count_dev(ctx.handle, reads.data_ptr(), reads.numel(), C.byref(n)))
    kp, cp = C.c_void_p(), C.c_void_p()
    ctx.check(lib.aix_canonical23_result_dev(ctx.handle, C.byref(kp), C.byref(cp), None))
    n = int(n.value)
    mark("canonical 23-mer table (emit + sort + run-length)")
    mphf = capi.Mphf.build_dev(ctx, kp.value, n, 23)
    mark("MPHF construction (parallel peeling)")
    checker = torch.empty(n, device=reads.device, dtype=torch.int64)
    tf = torch.empty(n, device=reads.device, dtype=torch.int32)
    ctx.check(lib.aix_index23_fill_dev(ctx.handle, mphf._h, kp.value, cp.value, n, checker.data_ptr(), tf.data_ptr()))
    mark("checker / tf fill")
    index = capi.Index23.upload_dev(ctx, mphf, checker.data_ptr(), tf.data_ptr(), n)
    mark("index upload (records + fused MPHF records / fingerprint tier)")
    return mphf, index, checker, tf, n


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md: -lms 200).  A faster poll
    contends with the driver for its lock: at -lms 50 the cudaMalloc / cudaFree of the 100 GB the positions build
    uses took up to 10x longer (profiles/r02_c5_trace.txt)."""

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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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




def _wrap_device_i64(torch, ptr, n, dev):
    """torch view of a device buffer owned by libaindex_cuda (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=dev)
