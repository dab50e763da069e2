"""Sharding and the one collective of the hot path, for one-process-per-GPU runs.

The path shards naturally:
  * tf / coverage / position queries: index replicated on every GPU, queries split evenly,
    results concatenated -- no data-path collective;
  * 13-mer counting: reads split into contiguous byte ranges cut at line (record) boundaries,
    every GPU builds a full direct-address 4^13 histogram, then ONE exchange step: an NCCL
    reduce-scatter (sum) over k-mer ranges; rank r owns [r, r+1) * 4^13 / world.

Everything here is host logic on torch.distributed process groups: it runs on NCCL with the
device histogram exposed by libaindex_cuda (aix_count13_hist_dev) and on gloo with CPU tensors
(tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

TOTAL_13MERS = 1 << 26


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def kmer_range(rank: int, world: int) -> Tuple[int, int]:
    """k-mer value range of the 4^13 histogram owned by `rank` after the reduce-scatter."""
    if TOTAL_13MERS % world:
        raise ValueError("world size must divide 4^13 (use 1, 2, 4, 8, ...)")
    step = TOTAL_13MERS // world
    return rank * step, (rank + 1) * step


def shard_reads(data: np.ndarray, world: int, lines_per_record: int = 1) -> List[Tuple[int, int]]:
    """Byte ranges [begin, end) of a reads buffer, one per rank, cut right after a '\\n' that
    ends a record (plain text: every line; FASTQ: every 4th line).  The ranges are disjoint,
    cover the buffer, and every range starts at a record start, so per-shard k-mer counting
    sees exactly the windows of the whole file (no window spans a newline)."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    n = int(data.size)
    cuts = [0]
    if world > 1 and n:
        if lines_per_record == 1:
            for r in range(1, world):
                target = max(cuts[-1], (n * r) // world)
                # next newline at or after target-1 ends a line; cut after it.  Searched in slices whose
                # upper bound advances every pass (chromosome-length lines), cut = n when there is none.
                lo, cut, span = max(target - 1, 0), n, 1 << 20
                while lo < n:
                    hi = min(n, lo + span)
                    nl = np.flatnonzero(data[lo:hi] == 10)
                    if nl.size:
                        cut = lo + int(nl[0]) + 1
                        break
                    lo, span = hi, min(span * 4, 1 << 28)
                cuts.append(min(max(cut, cuts[-1]), n))
        else:
            nl_pos = np.flatnonzero(data == 10)
            rec_ends = nl_pos[lines_per_record - 1::lines_per_record] + 1  # byte after each record
            for r in range(1, world):
                target = (n * r) // world
                i = int(np.searchsorted(rec_ends, target, side="left"))
                cut = int(rec_ends[i]) if i < rec_ends.size else n
                cuts.append(min(max(cut, cuts[-1]), n))
    while len(cuts) < world:
        cuts.append(n)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_fasta(data: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Byte ranges of a FASTA buffer, one per rank, every range starting at a header line ('>' at a line start), so
    no record (whose lines are concatenated before counting, count_kmers13.cpp:211-235) is split between ranks."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    n = int(data.size)
    cuts = [0]
    if world > 1 and n:
        gt = np.flatnonzero(data == ord(">"))
        starts = gt[(gt == 0) | (data[np.maximum(gt - 1, 0)] == 10)]  # '>' at a line start only
        for r in range(1, world):
            target = (n * r) // world
            i = int(np.searchsorted(starts, target, side="left"))
            cut = int(starts[i]) if i < starts.size else n
            cuts.append(min(max(cut, cuts[-1]), n))
    while len(cuts) < world:
        cuts.append(n)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def reduce_scatter_hist(hist, group=None):
    """Sum the per-rank direct-address histograms; return this rank's k-mer slice.

    hist: int64 tensor [4^13] (a view of the library's device buffer on NCCL, a CPU tensor on gloo).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return hist
    lo, hi = kmer_range(rank, world)
    out = torch.empty(hi - lo, dtype=hist.dtype, device=hist.device)
    try:
        dist.reduce_scatter_tensor(out, hist, op=dist.ReduceOp.SUM, group=group)
    except (RuntimeError, NotImplementedError):
        # backends without reduce-scatter (gloo): all-reduce a copy, keep the own slice
        tmp = hist.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out.copy_(tmp[lo:hi])
    return out


class PeerHistogram:
    """The fused form of `flush + reduce_scatter_hist` for one-node runs: every rank maps the u32 histograms of
    all ranks (CUDA IPC over NVLink / NVSwitch) and sums its k-mer range out of them in one kernel
    (aix_count13_reduce_peers_dev).  The two stream-ordered barriers around the kernel are tiny NCCL all-reduces.

        ph = PeerHistogram(ctx, group)            # once; raises RuntimeError if peer mapping is not possible
        ... aix_count13_begin / add ...
        mine = ph.reduce(stream)                  # int64[4^13 / world], this rank's range, summed over ranks
    """

    def __init__(self, ctx, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import capi
        self.ctx, self.group, self.lib = ctx, group, capi.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        # no rank may leave this constructor early: every failure is carried to the agreement all-reduce below
        mine = (C.c_uint8 * 192)()
        rc = self.lib.aix_count13_ipc_export(ctx.handle, mine)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t, group=group)
        blob = bytes(torch.cat(parts).cpu().numpy().tobytes())
        if rc == 0:
            rc = self.lib.aix_count13_peers_open(ctx.handle, blob, self.world, self.rank)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # all ranks agree on the path
        if int(ok.item()) == 0:
            self.lib.aix_count13_peers_close(ctx.handle)
            raise RuntimeError("peer mapping of the 13-mer histograms failed on some rank")
        self.lo, self.hi = kmer_range(self.rank, self.world)
        self.out = torch.empty(self.hi - self.lo, dtype=torch.int64, device=self.dev)
        self._token = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def reduce(self, stream):
        """Enqueue barrier -> gather-reduce kernel -> barrier on `stream` (the library's stream); returns the slice."""
        import torch
        import torch.distributed as dist
        with torch.cuda.stream(stream):
            dist.all_reduce(self._token, group=self.group)   # every rank has finished counting
            self.ctx.check(self.lib.aix_count13_reduce_peers_dev(self.ctx.handle, self.lo, self.hi, self.out.data_ptr()))
            dist.all_reduce(self._token, group=self.group)   # nobody clears its histogram while a peer still reads it
        return self.out

    def close(self):
        self.lib.aix_count13_peers_close(self.ctx.handle)


class ShardedIndex23:
    """A 23-mer index whose {checker, tf} records are split by hash-id range over the ranks (the MPHF, 0.4 B/key, is
    replicated): for indexes that do not fit one GPU.  `query` is a collective: every rank passes its own batch.

        probes (aix_tf23_probes_dev) -> all-to-all by owner of the id -> verify (aix_probe23_dev) -> all-to-all back
        -> first hit of the (at most two) probes of a query.

    The two device steps are methods so that the host logic (bucketing, the two all-to-alls, the inverse permutation,
    the combine rule) can be exercised on CPU tensors with gloo (tests/test_dist_cpu.py)."""

    def __init__(self, n_total: int, group=None):
        import torch.distributed as dist
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_total = int(n_total)
        self.bounds = [shard_range(self.n_total, r, self.world)[0] for r in range(self.world)] + [self.n_total]
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]

    # ---- device steps (GPU implementation; tests substitute CPU stand-ins) ------------------------------------
    def attach(self, ctx, mphf, checker, tf, stream):
        """Upload this rank's slice of the host arrays checker (u64[n]) / tf (u32[n])."""
        import torch
        import torch.distributed as dist
        from . import capi
        self.ctx, self.mphf, self.stream, self.lib = ctx, mphf, stream, capi.lib()
        self.shard = capi.Index23.upload(ctx, mphf, np.ascontiguousarray(checker[self.lo:self.hi]),
                                         np.ascontiguousarray(tf[self.lo:self.hi]))
        dev = torch.device("cuda", torch.cuda.current_device())
        flag = torch.tensor([1 if self.shard.info["canonical_only"] else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.canonical_only = int(flag.item())
        return self

    def _probes(self, recs):
        """recs: uint8[q, 23] device tensor -> int64[2q, 2] {global id or -1, packed k-mer}"""
        import torch
        q = recs.shape[0]
        out = torch.empty((2 * q, 2), dtype=torch.int64, device=recs.device)
        self.ctx.check(self.lib.aix_tf23_probes_dev(self.ctx.handle, self.mphf._h, self.n_total, self.canonical_only,
                                                    recs.data_ptr(), recs.shape[1], None, q, out.data_ptr()))
        return out

    def _verify(self, probes):
        """probes: int64[c, 2] {local id, k-mer} -> int64[c] (hit << 32) | tf"""
        import torch
        out = torch.empty(probes.shape[0], dtype=torch.int64, device=probes.device)
        self.ctx.check(self.lib.aix_probe23_dev(self.ctx.handle, self.shard._h, probes.data_ptr(), probes.shape[0], out.data_ptr()))
        return out

    def _bucket(self, probes):
        """probes int64[2q, 2] -> (send int64[c, 2] {local id, k-mer} grouped by owner in rank order,
        tag int64[c] probe index of every row, counts int64[world]).  GPU: one counting sort on the device
        (aix_probes_bucket_dev); CPU tensors (gloo tests): the same thing with torch ops."""
        import ctypes as C
        import torch
        n = probes.shape[0]
        if probes.is_cuda:
            counts = torch.empty(self.world, dtype=torch.int64, device=probes.device)
            send = torch.empty((n, 2), dtype=torch.int64, device=probes.device)
            tag = torch.empty(n, dtype=torch.int32, device=probes.device)
            bounds = (C.c_uint64 * (self.world + 1))(*self.bounds)
            self.ctx.check(self.lib.aix_probes_bucket_dev(self.ctx.handle, probes.data_ptr(), n, bounds, self.world,
                                                          counts.data_ptr(), send.data_ptr(), tag.data_ptr()))
            total = int(counts.sum().item())  # the split sizes are needed on the host anyway
            return send[:total], tag[:total].long(), counts
        ids = probes[:, 0]
        live = torch.nonzero(ids >= 0).reshape(-1)                         # -1 = no probe
        inner = torch.tensor(self.bounds[1:-1], dtype=torch.int64)
        owner = torch.bucketize(ids[live], inner, right=True)              # rank that holds the id
        order = torch.argsort(owner, stable=True)
        live, owner = live[order], owner[order]
        lo = torch.tensor(self.bounds[:-1], dtype=torch.int64)
        send = torch.stack([ids[live] - lo[owner], probes[live, 1]], dim=1).contiguous()
        return send, live, torch.bincount(owner, minlength=self.world)

    # ---- host logic --------------------------------------------------------------------------------------------
    def query(self, recs):
        """tf of every 23-byte record of this rank's batch (uint8[q, 23] tensor) -> int64[q] tensor (values < 2^32)."""
        import contextlib
        import torch
        import torch.distributed as dist
        stream = getattr(self, "stream", None)
        cm = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
        with cm:
            q = recs.shape[0]
            probes = self._probes(recs)
            send, tag, n_send = self._bucket(probes)
            n_recv = torch.empty_like(n_send)
            dist.all_to_all_single(n_recv, n_send, group=self.group)
            s_split, r_split = n_send.tolist(), n_recv.tolist()
            recv = torch.empty((sum(r_split), 2), dtype=torch.int64, device=probes.device)
            dist.all_to_all_single(recv, send, r_split, s_split, group=self.group)
            answers = self._verify(recv)
            back = torch.empty(send.shape[0], dtype=torch.int64, device=probes.device)
            dist.all_to_all_single(back, answers, s_split, r_split, group=self.group)
            res = torch.zeros(2 * q, dtype=torch.int64, device=probes.device)
            res[tag] = back
            res = res.reshape(q, 2)
            first_hit = (res[:, 0] >> 32) != 0
            return torch.where(first_hit, res[:, 0], res[:, 1]) & 0xFFFFFFFF


def sum_to_rank0(t, group=None):
    """Element-wise sum of a tensor onto rank 0 (the scattered MPHF-order partial results)."""
    import torch.distributed as dist
    if dist.get_world_size(group) > 1:
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_concat(local: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank result arrays (queries were split with shard_range) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local), group=group)
    return np.concatenate(parts)


def count13_distributed(ctx, mphf, shard: np.ndarray, fmt: int, group=None):
    """13-mer counting of one rank's shard + reduce-scatter + MPHF permutation.

    Returns (tf uint64[4^13] in .tf.bin order on rank 0 else None, merged stats dict)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import capi
    lib = capi.lib()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    shard = np.ascontiguousarray(shard, dtype=np.uint8).reshape(-1)
    ctx.check(lib.aix_count13_begin(ctx.handle))
    ctx.check(lib.aix_count13_add(ctx.handle, shard.ctypes.data, shard.size, fmt))
    ctx.check(lib.aix_count13_flush(ctx.handle))
    st = capi.CountStats()
    ctx.check(lib.aix_count13_stats(ctx.handle, st))
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    hist = wrap_device_i64(lib.aix_count13_hist_dev(ctx.handle), TOTAL_13MERS, dev)
    tf_dev = torch.zeros(TOTAL_13MERS, dtype=torch.int64, device=dev)
    with torch.cuda.stream(stream):
        mine = reduce_scatter_hist(hist, group)
        lo, hi = kmer_range(rank, world)
        if world > 1:
            hist[lo:hi].copy_(mine)  # the library permutes out of its own buffer
        ctx.check(lib.aix_count13_finish_dev(ctx.handle, mphf._h, lo, hi, tf_dev.data_ptr()))
        sum_to_rank0(tf_dev, group)
        stats = torch.tensor([st.sequences, st.windows, st.valid, st.invalid], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(stats, group=group)
    stream.synchronize()
    ctx.check(lib.aix_count13_end(ctx.handle))
    keys = ("sequences", "windows", "valid", "invalid")
    merged = dict(zip(keys, (int(x) for x in stats.cpu())))
    return (tf_dev.cpu().numpy().view(np.uint64) if rank == 0 else None), merged


def wrap_device_i64(ptr: int, n: int, dev):
    """torch int64 view of a device buffer owned by libaindex_cuda (no copy)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=dev)
