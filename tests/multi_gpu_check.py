"""Run under torchrun (one rank per GPU): distributed 13-mer counting (shard -> count ->
NCCL reduce-scatter -> permute -> sum to rank 0) must equal the single-GPU result bit for bit,
and sharded 23-mer queries must concatenate to the single-GPU answers."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from aindex_b200 import capi, dist as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    g = os.path.join(ROOT, "tests", "golden")
    ix = capi.Index23.load_prefix(ctx, os.path.join(g, "idx23"))
    rng = np.random.default_rng(5)
    reads = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=(40000, 101), p=[.2475, .2475, .2475, .2475, .01])
    reads[:, 100] = 10
    data = reads.reshape(-1)
    # the 13-mer MPHF: the reference-built one when present, else built on the GPU (identical on
    # every rank: same keys, same seed sequence)
    pf13 = os.path.join(ROOT, "oracle", "_ref", "data", "all_13mers.pf")
    m13 = capi.Mphf.load(ctx, pf13) if os.path.exists(pf13) else capi.Mphf.build(ctx, np.arange(1 << 26, dtype=np.uint64), 13)
    b, e = D.shard_reads(data, world)[rank]
    tf, st = D.count13_distributed(ctx, m13, data[b:e], capi.FMT_PLAIN)
    ok = True
    if rank == 0:
        want, wst = ctx.count13(m13, data, capi.FMT_PLAIN)
        ok = bool(np.array_equal(tf, want)) and st == wst
        print(f"count13 x{world}: equal={ok} valid={st['valid']} md5={hashlib.md5(tf.tobytes()).hexdigest()}")
    # the fused combine over peer memory must give the slices the NCCL reduce-scatter gives
    lib = capi.lib()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    try:
        peer = D.PeerHistogram(ctx)
    except RuntimeError as ex:
        peer = None
        if rank == 0:
            print(f"peer combine unavailable: {ex}")
    if peer is not None:
        for force_flush in (False, True):
            ctx.check(lib.aix_count13_begin(ctx.handle))
            half = ((e - b) // 2 // 101) * 101
            ctx.check(lib.aix_count13_add(ctx.handle, data[b:b + half].ctypes.data, half, capi.FMT_PLAIN))
            if force_flush and rank % 2 == 0:
                ctx.check(lib.aix_count13_flush(ctx.handle))  # some ranks hold part of their counts in the u64 histogram
            ctx.check(lib.aix_count13_add(ctx.handle, data[b + half:e].ctypes.data, e - b - half, capi.FMT_PLAIN))
            out = peer.reduce(stream)
            with torch.cuda.stream(stream):  # everything that touches the library's buffers stays on its stream
                mine = out.clone()
            stream.synchronize()
            ctx.check(lib.aix_count13_flush(ctx.handle))
            hist = D.wrap_device_i64(lib.aix_count13_hist_dev(ctx.handle), 1 << 26, torch.device("cuda", local))
            with torch.cuda.stream(stream):
                ref = D.reduce_scatter_hist(hist)
            stream.synchronize()
            same = bool(torch.equal(mine, ref))
            ok = ok and same
            ctx.check(lib.aix_count13_end(ctx.handle))
            if rank == 0:
                print(f"peer combine (flush on even ranks={force_flush}): equal={same}")
        peer.close()
    # sharded queries
    q = ctx.decode(np.fromfile(os.path.join(g, "idx23.kmers.bin"), dtype=np.uint64)[:5000], 23)
    q[::3] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(len(q[::3]), 23))
    qb, qe = D.shard_range(len(q), rank, world)
    got = D.gather_concat(ix.query(q[qb:qe]))
    ok = ok and bool(np.array_equal(got, ix.query(q)))
    # index split by hash-id range: every rank holds 1/world of the records and queries its OWN batch; the answers
    # must equal those of the replicated index (hits on both strands, misses, non-ACGT and lower-case bytes)
    kb = np.fromfile(os.path.join(g, "idx23.kmers.bin"), dtype=np.uint64)
    tfb = np.fromfile(os.path.join(g, "idx23.tf.bin"), dtype=np.uint32)
    sh = D.ShardedIndex23(kb.size).attach(ctx, ix.mphf, kb, tfb, stream)
    rng2 = np.random.default_rng(900 + rank)
    recs = rng2.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(20011, 23))
    hit = rng2.random(recs.shape[0]) < 0.6
    recs[hit] = ctx.decode(kb[rng2.integers(0, kb.size, size=int(hit.sum()))], 23)
    flip = hit & (rng2.random(recs.shape[0]) < 0.5)
    recs[flip] = ctx.decode(ctx.revcomp(ctx.encode(recs[flip], 23), 23), 23)
    odd = rng2.random(recs.shape[0]) < 0.05
    recs[odd, rng2.integers(0, 23, size=int(odd.sum()))] = rng2.choice(np.frombuffer(b"Nnacgt~", dtype=np.uint8), size=int(odd.sum()))
    with torch.cuda.stream(stream):
        recs_t = torch.from_numpy(recs).cuda()
    got_sh = sh.query(recs_t)
    stream.synchronize()
    same = bool(np.array_equal(got_sh.cpu().numpy().astype(np.uint32), ix.query(recs)))
    ok = ok and same
    if os.environ.get("AIX_CHECK_TIMING"):
        import time
        big = torch.from_numpy(rng2.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(int(os.environ.get("AIX_CHECK_TIMING_Q", "5000000")), 23))).cuda()
        sh.query(big)
        stream.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            sh.query(big)
        stream.synchronize()
        dist.barrier()
        if rank == 0:
            print(f"sharded index x{world}: {world * 3 * big.shape[0] / (time.perf_counter() - t0) / 1e9:.2f} G queries/s aggregate ({big.shape[0]} per rank and call)")
    if rank == 0:
        print(f"sharded index x{world}: equal={same} hits={int((got_sh > 0).sum())} canonical_only={sh.canonical_only} range=[{sh.lo},{sh.hi})")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag.item()) else "MULTI_GPU_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
