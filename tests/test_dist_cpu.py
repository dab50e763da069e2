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


@pytest.mark.timeout(60)
def test_shard_reads_very_long_lines():
    """chromosome-length lines (plain reads built from genomes): the newline search must advance, not spin"""
    data = np.full(40 << 20, ord("A"), dtype=np.uint8)
    assert D.shard_reads(data, 2) == [(0, data.size), (data.size, data.size)]      # no newline at all
    data[-1] = 10
    assert D.shard_reads(data, 2) == [(0, data.size), (data.size, data.size)]      # one 40 MiB line
    data[30 << 20] = 10                                                           # newline 10 MiB past the split target
    assert D.shard_reads(data, 2) == [(0, (30 << 20) + 1), ((30 << 20) + 1, data.size)]
    s4 = D.shard_reads(data, 4)
    assert s4[0] == (0, (30 << 20) + 1) and s4[1][0] == s4[1][1] and s4[-1][1] == data.size


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


# ---------------------------------------------------------------------------- index split by hash-id range
def _sharded_worker(rank, world, port, q):
    """ShardedIndex23 host logic on gloo: the two device steps are replaced by oracle-based stand-ins."""
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        prefix = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "idx23")
        oix = O.Index23.load_prefix(prefix)
        NONE = -1

        class CpuSharded(D.ShardedIndex23):
            def _probes(self, recs):
                out = np.full((2 * recs.shape[0], 2), NONE, dtype=np.int64)
                for i, row in enumerate(recs.numpy()):
                    s = row.tobytes()
                    us = O.dna23_bitset(s)
                    rs = O.reverse_dna23(us)
                    if not s.strip(b"ACGT"):                       # valid: canonical index -> one probe of min(u, r)
                        c = min(us, rs)
                        h = oix.mphf.lookup(O.bitset_dna23(c).encode())
                        out[2 * i] = (h if h < oix.n else NONE, c)
                    else:                                          # raw bytes forward, decoded reverse complement backward
                        h1 = oix.mphf.lookup(s)
                        h2 = oix.mphf.lookup(O.bitset_dna23(rs).encode())
                        out[2 * i] = (h1 if h1 < oix.n else NONE, us)
                        out[2 * i + 1] = (h2 if h2 < oix.n else NONE, rs)
                return torch.from_numpy(out)

            def _verify(self, probes):
                p = probes.numpy()
                res = np.zeros(p.shape[0], dtype=np.int64)
                for j, (hl, km) in enumerate(p):
                    if 0 <= hl < self.hi - self.lo and int(oix.checker[self.lo + hl]) == int(km) & ((1 << 64) - 1):
                        res[j] = (1 << 32) | int(oix.tf[self.lo + hl])
                return torch.from_numpy(res)

        sh = CpuSharded(oix.n, None)
        assert sh.bounds[0] == 0 and sh.bounds[-1] == oix.n and len(sh.bounds) == world + 1
        rng = np.random.default_rng(100 + rank)                       # every rank has its own batch
        recs = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(300, 23))
        pick = rng.integers(0, oix.n, size=200)
        for j, kid in enumerate(pick):
            km = O.bitset_dna23(int(oix.checker[kid])).encode()
            if j % 2:
                km = km.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]
            recs[j] = np.frombuffer(km, dtype=np.uint8)
        recs[250:260, 5] = ord("N")
        recs[260:270, 0] = ord("a")
        got = sh.query(torch.from_numpy(recs)).numpy().astype(np.uint32)
        want = oix.batch(recs, None, O.MODE_TF)
        q.put((rank, bool(np.array_equal(got, want)), int((want > 0).sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_index_host_logic(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(r, True) for r in range(world)]
    assert all(r[2] >= 150 for r in res)  # the batches really contain hits on both strands
