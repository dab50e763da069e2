"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: read sharding, the k-mer range
reduce-scatter of the 13-mer histogram, query sharding + gather.  The per-shard counting is done
by the CPU oracle here (the GPU kernel is covered by the -m gpu tests); what is under test is
that shard -> count -> reduce-scatter -> concatenate reproduces the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from aindex_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _reads(seed=3, n=400, fastq=False):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        ln = int(rng.integers(5, 90))
        s = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=ln, p=[.245, .245, .245, .245, .02]).tobytes()
        out.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * ln) if fastq else s + b"\n")
    return np.frombuffer(b"".join(out), dtype=np.uint8)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_reads_plain_and_fastq(world):
    for fastq in (False, True):
        data = _reads(fastq=fastq)
        shards = D.shard_reads(data, world, 4 if fastq else 1)
        assert len(shards) == world and shards[0][0] == 0 and shards[-1][1] == data.size
        for (b0, e0), (b1, e1) in zip(shards, shards[1:]):
            assert e0 == b1 and b0 <= e0
        for b, e in shards:
            assert b == 0 or data[b - 1] == 10
            if fastq and e > b:
                assert data[b] == ord("@") and np.count_nonzero(data[b:e] == 10) % 4 == 0
    assert D.shard_reads(np.zeros(0, np.uint8), 4) == [(0, 0)] * 4
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert D.kmer_range(1, 4) == (1 << 24, 1 << 25)
    with pytest.raises(ValueError):
        D.kmer_range(0, 3)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_fasta_keeps_records_whole(world):
    """every shard starts at a header; the per-shard oracle counts add up to the whole file's"""
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(5)
    recs = []
    for i in range(60):
        recs.append(b">r%d with > inside" % i)
        for _ in range(int(rng.integers(0, 6))):
            recs.append(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(1, 70))).tobytes())
    data = np.frombuffer(b"\n".join(recs) + b"\n", dtype=np.uint8)
    shards = D.shard_fasta(data, world)
    assert len(shards) == world and shards[0][0] == 0 and shards[-1][1] == data.size
    whole, wst = O.count13_direct(data, O.FMT_FASTA)
    acc = np.zeros_like(whole)
    tot = {k: 0 for k in wst}
    for b, e in shards:
        assert b == e or (data[b] == ord(">") and (b == 0 or data[b - 1] == 10))
        if e > b:
            h, st = O.count13_direct(data[b:e], O.FMT_FASTA)
            acc += h
            for k in tot:
                tot[k] += st[k]
    assert np.array_equal(acc, whole) and tot == wst


def _worker(rank, world, port, fastq, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        data = _reads(fastq=fastq)
        fmt = O.FMT_FASTQ if fastq else O.FMT_PLAIN
        b, e = D.shard_reads(data, world, 4 if fastq else 1)[rank]
        hist, st = O.count13_direct(data[b:e], fmt)
        mine = D.reduce_scatter_hist(torch.from_numpy(hist.view(np.int64)))
        lo, hi = D.kmer_range(rank, world)
        whole, wst = O.count13_direct(data, fmt)
        ok = bool(np.array_equal(mine.numpy().view(np.uint64), whole[lo:hi]))
        stats = torch.tensor([st[k] for k in ("sequences", "windows", "valid", "invalid")])
        dist.all_reduce(stats)
        ok = ok and stats.tolist() == [wst[k] for k in ("sequences", "windows", "valid", "invalid")]
        # query sharding: every rank answers its slice, gather_concat restores the batch order
        full = np.arange(1001, dtype=np.uint32) * 7
        qb, qe = D.shard_range(full.size, rank, world)
        ok = ok and bool(np.array_equal(D.gather_concat(full[qb:qe]), full))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fastq", [False, True])
def test_count13_reduce_scatter_world2(fastq):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fastq, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
