"""Sweep of the 13-mer counting kernel variants (run on the GPU box; prints one line per config).
AIX_COUNT13_VARIANT: 0 one RED per window, 1 + thread run-length merge, 2 warp match_any merge.
AIX_COUNT13_PASSES_LOG2: the histogram is updated in 2^p passes over k-mer slices (L2 residency)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aindex_b200 import capi  # noqa: E402
from bench import make_reads  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ctx = capi.Context(0)
lib = capi.lib()
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
n_kmers = n_reads * 138


def make_repetitive_reads(n_reads, seed=7):
    """reads of a genome in which 30 % of the positions are homopolymer / dinucleotide tracts and copies of one
    5 kb element: the skewed histogram real genomes have (hot counters, runs of identical k-mers)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = 20_000_000
    genome = torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.uint8)
    elem = torch.randint(0, 4, (5000,), generator=g, device=dev, dtype=torch.uint8)
    for st in torch.randint(0, L - 5000, (600,), generator=g, device=dev).tolist():
        genome[st:st + 5000] = elem
    for st in torch.randint(0, L - 300, (15000,), generator=g, device=dev).tolist():
        ln = 50 + (st % 200)
        genome[st:st + ln] = st % 4 if st % 3 else torch.arange(ln, device=dev, dtype=torch.uint8) % 2 * 2
    lut = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    out = torch.empty((n_reads, 151), device=dev, dtype=torch.uint8)
    ar = torch.arange(150, device=dev)
    for s in range(0, n_reads, 2_000_000):
        e = min(n_reads, s + 2_000_000)
        start = torch.randint(0, L - 150, (e - s,), generator=g, device=dev)
        out[s:e, :150] = lut[genome[start[:, None] + ar[None, :]].long()]
    out[:, 150] = 10
    return out


data = sys.argv[2] if len(sys.argv) > 2 else "random"
reads = make_reads(torch, dev, 100_000_000, n_reads, 150, 11, 12) if data == "random" else make_repetitive_reads(n_reads)
print(f"# data={data} reads={n_reads}")
for variant in (0, 1, 2, 3):
    for p in ((2,) if data != "random" else (0, 1, 2, 3)):
        os.environ["AIX_COUNT13_VARIANT"] = str(variant)
        os.environ["AIX_COUNT13_PASSES_LOG2"] = str(p)
        ts = []
        for it in range(3):
            ctx.check(lib.aix_count13_begin(ctx.handle))
            ctx.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.check(lib.aix_count13_add_dev(ctx.handle, reads.data_ptr(), reads.numel(), capi.FMT_PLAIN))
            b.record(stream)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        st = capi.CountStats()
        ctx.check(lib.aix_count13_stats(ctx.handle, st))
        print(f"variant={variant} passes=2^{p} ms={min(ts):.2f} kmers/s={n_kmers / (min(ts) / 1e3):.3e} valid_ok={st.valid == n_kmers}", flush=True)
