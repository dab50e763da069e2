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
reads = make_reads(torch, dev, 100_000_000, n_reads, 150, 11, 12)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
n_kmers = n_reads * 138
for variant in (0, 1, 2):
    for p in (0, 1, 2, 3, 4):
        if variant == 2 and p:
            continue
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
