"""ctypes binding of the C-ABI (include/aindex_cuda.h) of libaindex_cuda.so.

This is the thinnest possible host layer over the CUDA library: numpy arrays in, numpy
arrays out, every call going through the `extern "C"` entry points a cgo / JNI / ctypes
binding of the reference would use.  There is no CPU fallback: if the shared library is
missing or no CUDA device is usable, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libaindex_cuda.so")
TOTAL_13MERS = 1 << 26

Q_TF, Q_TOTAL, Q_BOTH, Q_PFID, Q_STRAND, Q_KID = range(6)
FMT_DETECT, FMT_PLAIN, FMT_FASTA, FMT_FASTQ = -1, 0, 1, 2


class AixError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libaindex_cuda error {code}: {msg}")
        self.code = code


class CountStats(C.Structure):
    _fields_ = [("sequences", C.c_uint64), ("windows", C.c_uint64), ("valid", C.c_uint64),
                ("invalid", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class MultiBuildStats(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("upload_scan_ms", C.c_double), ("emit_partition_ms", C.c_double),
                ("exchange_ms", C.c_double), ("sort_finalize_ms", C.c_double), ("download_ms", C.c_double),
                ("keys", C.c_uint64), ("peer_bytes", C.c_uint64), ("positions", C.c_uint64), ("alloc_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

# name -> (restype, argtypes); every symbol include/aindex_cuda.h declares
_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "aix_ctx_create": (_i, [_i, _pp]),
    "aix_ctx_destroy": (None, [_vp]),
    "aix_last_error": (C.c_char_p, [_vp]),
    "aix_ctx_device": (_i, [_vp]),
    "aix_ctx_stream": (_vp, [_vp]),
    "aix_ctx_sync": (_i, [_vp]),
    "aix_ctx_launch_count": (_u64, [_vp]),
    "aix_ctx_trim": (_i, [_vp]),
    "aix_host_alloc": (_i, [_vp, C.c_size_t, _pp]),
    "aix_host_free": (_i, [_vp, _vp]),
    "aix_version": (C.c_char_p, []),
    "aix_mphf_upload": (_i, [_vp, _u64, _u64, _u64, _vp, _u64, _vp, _u64, _pp]),
    "aix_mphf_load_pf": (_i, [_vp, C.c_char_p, _pp]),
    "aix_mphf_save_pf": (_i, [_vp, _vp, C.c_char_p]),
    "aix_mphf_destroy": (None, [_vp, _vp]),
    "aix_mphf_info": (_i, [_vp, _vp]),
    "aix_mphf_arrays": (_i, [_vp, _vp, _vp]),
    "aix_mphf_lookup": (_i, [_vp, _vp, _vp, _u32, _vp, _u64, _vp]),
    "aix_jenkins64": (_i, [_vp, _u64, _vp, _u32, _vp, _u64, _vp]),
    "aix_perm13": (_i, [_vp, _vp, _vp]),
    "aix_mphf_build": (_i, [_vp, _vp, _u64, _i, _pp]),
    "aix_mphf_build_dev": (_i, [_vp, _vp, _u64, _i, _pp]),
    "aix_encode_kmers": (_i, [_vp, _vp, _u32, _vp, _u64, _i, _vp]),
    "aix_decode_kmers": (_i, [_vp, _vp, _u64, _i, _vp]),
    "aix_revcomp_kmers": (_i, [_vp, _vp, _u64, _i, _vp]),
    "aix_pack_2bit": (_i, [_vp, _vp, _u64, _vp]),
    "aix_rolling_kmers": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _vp]),
    "aix_rolling_kmers_dev": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _vp]),
    "aix_pack_2bit_dev": (_i, [_vp, _vp, _u64, _vp]),
    "aix_ukmers": (_i, [_vp, _vp, _u64, _vp, _u64, _i, _vp]),
    "aix_ukmers_dev": (_i, [_vp, _vp, _u64, _vp, _u64, _i, _vp]),
    "aix_index23_upload": (_i, [_vp, _vp, _vp, _vp, _u64, _pp]),
    "aix_index23_upload_dev": (_i, [_vp, _vp, _vp, _vp, _u64, _pp]),
    "aix_index23_load_prefix": (_i, [_vp, C.c_char_p, _pp, _pp]),
    "aix_index23_destroy": (None, [_vp, _vp]),
    "aix_index23_info": (_i, [_vp, _vp]),
    "aix_index23_layout": (_i, [_vp, _vp]),
    "aix_index23_fill": (_i, [_vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "aix_index23_fill_dev": (_i, [_vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "aix_tf23_batch": (_i, [_vp, _vp, _vp, _u32, _vp, _u64, _i, _vp]),
    "aix_tf23_batch_dev": (_i, [_vp, _vp, _vp, _u32, _vp, _u64, _i, _vp]),
    "aix_index23_filter_stats": (_i, [_vp, _vp]),
    "aix_index23_set_filter": (_i, [_vp, _i]),
    "aix_tf23_single_call_latency": (_i, [_vp, _vp, _vp, _u32, _u64, _vp, _vp, _vp]),
    "aix_tf23_probes_dev": (_i, [_vp, _vp, _u64, _i, _vp, _u32, _vp, _u64, _vp]),
    "aix_probe23_dev": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "aix_probes_bucket_dev": (_i, [_vp, _vp, _u64, _vp, _i, _vp, _vp, _vp]),
    "aix_get_freq23": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "aix_get_freq23_packed": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "aix_get_freq23_packed_dev": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "aix_index13_upload": (_i, [_vp, _vp, _vp, _pp]),
    "aix_index13_destroy": (None, [_vp, _vp]),
    "aix_index13_tf_direct": (_i, [_vp, _vp, _vp]),
    "aix_tf13_batch": (_i, [_vp, _vp, _vp, _u32, _vp, _u64, _i, _vp]),
    "aix_tf13_batch_dev": (_i, [_vp, _vp, _vp, _u32, _vp, _u64, _i, _vp]),
    "aix_count13": (_i, [_vp, _vp, _vp, _u64, _i, _vp, C.POINTER(CountStats)]),
    "aix_count13_begin": (_i, [_vp]),
    "aix_count13_add": (_i, [_vp, _vp, _u64, _i]),
    "aix_count13_add_dev": (_i, [_vp, _vp, _u64, _i]),
    "aix_count13_flush": (_i, [_vp]),
    "aix_count13_hist_dev": (_vp, [_vp]),
    "aix_count13_stats": (_i, [_vp, C.POINTER(CountStats)]),
    "aix_count13_finish": (_i, [_vp, _vp, _u64, _u64, _vp, C.POINTER(CountStats)]),
    "aix_count13_finish_dev": (_i, [_vp, _vp, _u64, _u64, _vp]),
    "aix_count13_end": (_i, [_vp]),
    "aix_count13_ipc_export": (_i, [_vp, _vp]),
    "aix_count13_peers_open": (_i, [_vp, _vp, _i, _i]),
    "aix_count13_reduce_peers_dev": (_i, [_vp, _u64, _u64, _vp]),
    "aix_count13_peers_close": (_i, [_vp]),
    "aix_multi_create": (_i, [_i, _vp, _pp]),
    "aix_multi_destroy": (None, [_vp]),
    "aix_multi_size": (_i, [_vp]),
    "aix_multi_ctx": (_vp, [_vp, _i]),
    "aix_multi_peer_access": (_i, [_vp]),
    "aix_multi_last_error": (C.c_char_p, [_vp]),
    "aix_count13_multi": (_i, [_vp, _vp, _vp, _u64, _i, _vp, C.POINTER(CountStats)]),
    "aix_count13_multi_dev": (_i, [_vp, _vp, _vp, _vp, _i, _vp, C.POINTER(CountStats)]),
    "aix_positions_build23_multi": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, C.POINTER(MultiBuildStats)]),
    "aix_coverage": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _i, _u32, _vp]),
    "aix_coverage_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _u64, _i, _u32, _vp]),
    "aix_positions_total23": (_i, [_vp, _vp, _vp]),
    "aix_positions_build23": (_i, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "aix_positions_total13": (_i, [_vp, _vp, _vp]),
    "aix_positions_build13": (_i, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "aix_positions_build23_dev": (_i, [_vp, _vp, _vp, _u64, _pp]),
    "aix_positions_build13_dev": (_i, [_vp, _vp, _vp, _u64, _pp]),
    "aix_positions_info": (_i, [_vp, _vp]),
    "aix_positions_arrays_dev": (_i, [_vp, _pp, _pp]),
    "aix_positions_download": (_i, [_vp, _vp, _vp, _vp]),
    "aix_positions_upload": (_i, [_vp, _vp, _u64, _vp, _u64, _pp]),
    "aix_positions_destroy": (None, [_vp, _vp]),
    "aix_positions_query": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _u64, _i, _vp, _vp, _vp]),
    "aix_positions_query_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _u64, _i, _vp, _vp, _vp]),
    "aix_canonical23_count": (_i, [_vp, _vp, _u64, _vp, _vp, _vp]),
    "aix_canonical23_count_dev": (_i, [_vp, _vp, _u64, _vp]),
    "aix_canonical23_result_dev": (_i, [_vp, _pp, _pp, _vp]),
    "aix_write_dat": (_i, [_vp, _vp, _vp, _u64, C.c_char_p, C.c_char_p]),
    "aix_sort_u64_dev": (_i, [_vp, _vp, _vp, _u64, _i, _i, C.POINTER(_i)]),
    "aix_partition_u64_dev": (_i, [_vp, _vp, _vp, _u64, _vp, _i, _vp]),
    "aix_rle_u64_dev": (_i, [_vp, _vp, _u64, _vp, _vp, C.POINTER(_u64)]),
}


def lib():
    """Load libaindex_cuda.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m aindex_b200.build` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a


def as_records(kmers, stride: Optional[int] = None):
    """list[str|bytes] or uint8[q, stride] -> (uint8[q, stride], lens or None)."""
    if isinstance(kmers, np.ndarray):
        recs = np.ascontiguousarray(kmers, dtype=np.uint8)
        if recs.ndim != 2:
            raise ValueError("record array must be 2-D uint8[q, stride]")
        return recs, None
    bs = [k.encode("latin-1") if isinstance(k, str) else bytes(k) for k in kmers]
    if stride is None:
        stride = max([len(b) for b in bs] + [1])
    uniform = all(len(b) == stride for b in bs)
    if uniform:
        return np.frombuffer(b"".join(bs), dtype=np.uint8).reshape(len(bs), stride), None
    recs = np.zeros((len(bs), stride), dtype=np.uint8)
    lens = np.zeros(len(bs), dtype=np.uint8)
    for i, b in enumerate(bs):
        if len(b) > stride or len(b) > 255:
            raise ValueError("query longer than the record stride / 255 bytes")
        recs[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        lens[i] = len(b)
    return recs, lens


class Context:
    """aix_ctx: one per (thread, GPU)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().aix_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise AixError(rc, (lib().aix_last_error(None) or b"").decode())

    def check(self, rc: int):
        if rc != 0:
            raise AixError(rc, (lib().aix_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            lib().aix_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return int(lib().aix_ctx_stream(self._h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().aix_ctx_launch_count(self._h))

    def sync(self):
        self.check(lib().aix_ctx_sync(self._h))

    def trim(self):
        """return the memory the builders' pool has cached to the driver"""
        self.check(lib().aix_ctx_trim(self._h))

    def pinned(self, shape, dtype) -> np.ndarray:
        """numpy array backed by cudaHostAlloc memory (freed when the array is collected)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        self.check(lib().aix_host_alloc(self._h, n, C.byref(ptr)))
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)
        import weakref
        weakref.finalize(buf, _free_pinned, ptr.value)
        return arr

    # ---- codec --------------------------------------------------------------------
    def jenkins64(self, seed: int, kmers) -> np.ndarray:
        recs, lens = as_records(kmers)
        out = np.zeros((recs.shape[0], 3), dtype=np.uint64)
        self.check(lib().aix_jenkins64(self._h, seed, _p(recs), recs.shape[1], _p(lens), recs.shape[0], _p(out)))
        return out

    def encode(self, kmers, k: int) -> np.ndarray:
        recs, lens = as_records(kmers)
        out = np.zeros(recs.shape[0], dtype=np.uint64)
        self.check(lib().aix_encode_kmers(self._h, _p(recs), recs.shape[1], _p(lens), recs.shape[0], k, _p(out)))
        return out

    def decode(self, values, k: int) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.zeros((v.size, k), dtype=np.uint8)
        self.check(lib().aix_decode_kmers(self._h, _p(v), v.size, k, _p(out)))
        return out

    def revcomp(self, values, k: int) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.zeros(v.size, dtype=np.uint64)
        self.check(lib().aix_revcomp_kmers(self._h, _p(v), v.size, k, _p(out)))
        return out

    def pack_2bit(self, seq) -> np.ndarray:
        a = _bytes(seq)
        out = np.zeros((a.size + 3) // 4, dtype=np.uint8)
        self.check(lib().aix_pack_2bit(self._h, _p(a), a.size, _p(out)))
        return out

    def ukmers(self, packed, n_bases: int, pos, k: int) -> np.ndarray:
        """dna_bitset::ukmer(pos, k) for every position of `pos` (packed = pack_2bit of the sequence)"""
        pk = np.ascontiguousarray(packed, dtype=np.uint8)
        ps = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros(ps.size, dtype=np.uint64)
        self.check(lib().aix_ukmers(self._h, _p(pk), n_bases, _p(ps), ps.size, k, _p(out)))
        return out

    def rolling_kmers(self, data, k: int):
        a = _bytes(data)
        n = max(0, a.size - k + 1)
        fwd = np.zeros(n, dtype=np.uint64)
        rc = np.zeros(n, dtype=np.uint64)
        valid = np.zeros(n, dtype=np.uint8)
        self.check(lib().aix_rolling_kmers(self._h, _p(a), a.size, k, _p(fwd), _p(rc), _p(valid)))
        return fwd, rc, valid

    # ---- counting -------------------------------------------------------------------
    def count13(self, mphf: "Mphf", data, fmt: int = FMT_DETECT):
        a = _bytes(data)
        out = np.zeros(TOTAL_13MERS, dtype=np.uint64)
        st = CountStats()
        self.check(lib().aix_count13(self._h, mphf._h, _p(a), a.size, fmt, _p(out), C.byref(st)))
        return out, st.as_dict()

    def canonical23_count(self, reads):
        a = _bytes(reads)
        n = C.c_uint64()
        self.check(lib().aix_canonical23_count(self._h, _p(a), a.size, C.byref(n), None, None))
        kmers = np.zeros(n.value, dtype=np.uint64)
        counts = np.zeros(n.value, dtype=np.uint32)
        self.check(lib().aix_canonical23_count(self._h, _p(a), a.size, C.byref(n), _p(kmers), _p(counts)))
        return kmers, counts


    def write_dat(self, kmers, counts, dat_path=None, keys_path=None):
        """`.dat` ("KMER\\tCOUNT") and / or key-file ("KMER") text of a canonical 23-mer table"""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        counts = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
        self.check(lib().aix_write_dat(self._h, _p(kmers), _p(counts), kmers.size,
                                       os.fsencode(dat_path) if dat_path else None, os.fsencode(keys_path) if keys_path else None))


def pack23(kmers) -> np.ndarray:
    """uint8[q, 23] upper-case ACGT records -> uint8[q, 6] dna_bitset records (dna_bitseq.hpp:22-61: 4 bases per byte,
    first base in bits 7:6, non-ACGT -> A; the last two bits are zero).  Host-side helper for the 6-byte query form."""
    recs, _ = as_records(kmers, 23)
    code = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    c = np.zeros((recs.shape[0], 24), dtype=np.uint8)
    c[:, :23] = code[recs[:, :23]]
    c = c.reshape(-1, 6, 4)
    return ((c[:, :, 0] << 6) | (c[:, :, 1] << 4) | (c[:, :, 2] << 2) | c[:, :, 3]).astype(np.uint8)


def _free_pinned(ptr):
    try:
        lib().aix_host_free(None, ptr)
    except Exception:
        pass


def _bytes(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    if isinstance(b, str):
        b = b.encode("latin-1")
    return np.frombuffer(bytes(b), dtype=np.uint8)


class Mphf:
    """aix_mphf: emphf::mphf<jenkins64_hasher> resident in HBM."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    @classmethod
    def load(cls, ctx: Context, pf_path: str) -> "Mphf":
        h = C.c_void_p()
        ctx.check(lib().aix_mphf_load_pf(ctx.handle, os.fsencode(pf_path), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_arrays(cls, ctx: Context, n, hash_domain, seed, words, block_ranks) -> "Mphf":
        words = np.ascontiguousarray(words, dtype=np.uint64)
        block_ranks = np.ascontiguousarray(block_ranks, dtype=np.uint64)
        h = C.c_void_p()
        ctx.check(lib().aix_mphf_upload(ctx.handle, n, hash_domain, seed, _p(words), words.size,
                                        _p(block_ranks), block_ranks.size, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def build(cls, ctx: Context, kmers, k: int = 23) -> "Mphf":
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        h = C.c_void_p()
        ctx.check(lib().aix_mphf_build(ctx.handle, _p(kmers), kmers.size, k, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def build_dev(cls, ctx: Context, kmers_dev_ptr: int, n: int, k: int = 23) -> "Mphf":
        h = C.c_void_p()
        ctx.check(lib().aix_mphf_build_dev(ctx.handle, kmers_dev_ptr, n, k, C.byref(h)))
        return cls(ctx, h)

    def close(self):
        if self._h:
            lib().aix_mphf_destroy(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self):
        a = np.zeros(6, dtype=np.uint64)
        self.ctx.check(lib().aix_mphf_info(self._h, _p(a)))
        return dict(zip(("n", "hash_domain", "seed", "bv_size", "n_words", "n_blocks"), map(int, a)))

    n = property(lambda s: s.info["n"])

    def arrays(self):
        i = self.info
        words = np.zeros(i["n_words"], dtype=np.uint64)
        ranks = np.zeros(i["n_blocks"], dtype=np.uint64)
        self.ctx.check(lib().aix_mphf_arrays(self._h, _p(words), _p(ranks)))
        return words, ranks

    def save(self, path: str):
        self.ctx.check(lib().aix_mphf_save_pf(self.ctx.handle, self._h, os.fsencode(path)))

    def lookup(self, kmers) -> np.ndarray:
        recs, lens = as_records(kmers)
        out = np.zeros(recs.shape[0], dtype=np.uint64)
        self.ctx.check(lib().aix_mphf_lookup(self.ctx.handle, self._h, _p(recs), recs.shape[1], _p(lens),
                                             recs.shape[0], _p(out)))
        return out

    def perm13(self) -> np.ndarray:
        out = np.zeros(TOTAL_13MERS, dtype=np.uint32)
        self.ctx.check(lib().aix_perm13(self.ctx.handle, self._h, _p(out)))
        return out


def _out_for(mode: int, q: int, k: int):
    if mode == Q_TF:
        return np.zeros(q, dtype=np.uint32)
    if mode == Q_BOTH:
        return np.zeros((q, 2), dtype=np.uint32 if k == 23 else np.uint64)
    return np.zeros(q, dtype=np.uint64)


class Index23:
    """aix_index23: PHASH_MAP ({checker, tf} records) resident in HBM."""

    def __init__(self, ctx: Context, mphf: Mphf, handle):
        self.ctx, self.mphf, self._h = ctx, mphf, handle

    @classmethod
    def upload(cls, ctx: Context, mphf: Mphf, checker, tf) -> "Index23":
        checker = np.ascontiguousarray(checker, dtype=np.uint64)
        tf = np.ascontiguousarray(tf, dtype=np.uint32)
        if checker.size != tf.size:
            raise ValueError("checker and tf must have the same length")
        h = C.c_void_p()
        ctx.check(lib().aix_index23_upload(ctx.handle, mphf._h, _p(checker), _p(tf), checker.size, C.byref(h)))
        return cls(ctx, mphf, h)

    @classmethod
    def upload_dev(cls, ctx: Context, mphf: Mphf, checker_ptr: int, tf_ptr: int, n: int) -> "Index23":
        h = C.c_void_p()
        ctx.check(lib().aix_index23_upload_dev(ctx.handle, mphf._h, checker_ptr, tf_ptr, n, C.byref(h)))
        return cls(ctx, mphf, h)

    @classmethod
    def load_prefix(cls, ctx: Context, prefix: str) -> "Index23":
        mh, h = C.c_void_p(), C.c_void_p()
        ctx.check(lib().aix_index23_load_prefix(ctx.handle, os.fsencode(prefix), C.byref(mh), C.byref(h)))
        return cls(ctx, Mphf(ctx, mh), h)

    def close(self):
        if self._h:
            lib().aix_index23_destroy(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self):
        a = np.zeros(2, dtype=np.uint64)
        self.ctx.check(lib().aix_index23_info(self._h, _p(a)))
        return {"n": int(a[0]), "canonical_only": bool(a[1])}

    @property
    def layout(self):
        a = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().aix_index23_layout(self._h, _p(a)))
        return {"fp_bits": int(a[0]), "fp_bytes": int(a[1]), "mphf_bytes": int(a[2]), "mphf_compact": int(a[3]) >= 1,
                "records": ("wide", "compact", "fused")[int(a[3])]}

    def query(self, kmers, mode: int = Q_TF, out: Optional[np.ndarray] = None) -> np.ndarray:
        recs, lens = as_records(kmers)
        q = recs.shape[0]
        if out is None:
            out = _out_for(mode, q, 23)
        self.ctx.check(lib().aix_tf23_batch(self.ctx.handle, self._h, _p(recs), recs.shape[1], _p(lens), q,
                                            mode, _p(out)))
        return out

    @property
    def filter_stats(self) -> dict:
        a = np.zeros(6, dtype=np.uint64)
        self.ctx.check(lib().aix_index23_filter_stats(self._h, _p(a)))
        rate = None if int(a[5]) == 0xFFFFFFFFFFFFFFFF else int(a[5]) / 1e6
        return {"filter_bytes": int(a[0]), "queries_counted": int(a[1]), "passed": int(a[2]), "batches_filter": int(a[3]),
                "batches_direct": int(a[4]), "pass_rate": rate}

    def set_filter(self, mode: str = "auto"):
        self.ctx.check(lib().aix_index23_set_filter(self._h, {"auto": 0, "on": 1, "off": 2}[mode]))

    def single_call_latency(self, kmers) -> dict:
        """ns per single get_tf_value call measured from C (aix_tf23_single_call_latency): echo = transport only."""
        recs, _ = as_records(kmers)
        n = recs.shape[0]
        tf = np.zeros(n, dtype=np.uint32)
        echo, query = C.c_double(0), C.c_double(0)
        self.ctx.check(lib().aix_tf23_single_call_latency(self.ctx.handle, self._h, _p(recs), recs.shape[1], n, _p(tf),
                                                          C.byref(echo), C.byref(query)))
        return {"echo_ns": echo.value, "query_ns": query.value, "tf": tf}

    def query_dev(self, recs_ptr: int, stride: int, lens_ptr, q: int, mode: int, out_ptr: int):
        self.ctx.check(lib().aix_tf23_batch_dev(self.ctx.handle, self._h, recs_ptr, stride, lens_ptr, q, mode,
                                                out_ptr))

    def get_freq(self, ukmers) -> np.ndarray:
        u = np.ascontiguousarray(ukmers, dtype=np.uint64)
        out = np.zeros(u.size, dtype=np.uint32)
        self.ctx.check(lib().aix_get_freq23(self.ctx.handle, self._h, _p(u), u.size, _p(out)))
        return out

    def get_freq_packed(self, packed6) -> np.ndarray:
        """tf of 23-mers given as uint8[q, 6] dna_bitset records (pack23)"""
        p6 = np.ascontiguousarray(packed6, dtype=np.uint8).reshape(-1, 6)
        out = np.zeros(p6.shape[0], dtype=np.uint32)
        self.ctx.check(lib().aix_get_freq23_packed(self.ctx.handle, self._h, _p(p6), p6.shape[0], _p(out)))
        return out

    def coverage(self, seqs, offs=None, cutoff: int = 0) -> np.ndarray:
        return _coverage(self.ctx, self, None, seqs, offs, 23, cutoff)

    def positions_total(self) -> int:
        t = C.c_uint64()
        self.ctx.check(lib().aix_positions_total23(self.ctx.handle, self._h, C.byref(t)))
        return int(t.value)

    def positions_build(self, reads):
        a = _bytes(reads)
        n = self.info["n"]
        indices = np.zeros(n + 1, dtype=np.uint64)
        positions = np.zeros(self.positions_total(), dtype=np.uint64)
        self.ctx.check(lib().aix_positions_build23(self.ctx.handle, self._h, _p(a), a.size, _p(indices),
                                                   _p(positions)))
        return indices, positions


class Index13:
    """aix_index13: the 4^13 x u64 tf array (MPHF order + direct-address copy) in HBM."""

    def __init__(self, ctx: Context, mphf: Mphf, handle):
        self.ctx, self.mphf, self._h = ctx, mphf, handle

    @classmethod
    def upload(cls, ctx: Context, mphf: Mphf, tf64) -> "Index13":
        tf64 = np.ascontiguousarray(tf64, dtype=np.uint64)
        if tf64.size != TOTAL_13MERS:
            raise ValueError("13-mer tf array must have 4^13 entries")
        h = C.c_void_p()
        ctx.check(lib().aix_index13_upload(ctx.handle, mphf._h, _p(tf64), C.byref(h)))
        return cls(ctx, mphf, h)

    def close(self):
        if self._h:
            lib().aix_index13_destroy(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, kmers, mode: int = Q_TF) -> np.ndarray:
        recs, lens = as_records(kmers)
        q = recs.shape[0]
        out = _out_for(mode, q, 13)
        self.ctx.check(lib().aix_tf13_batch(self.ctx.handle, self._h, _p(recs), recs.shape[1], _p(lens), q,
                                            mode, _p(out)))
        return out

    def coverage(self, seqs, offs=None, cutoff: int = 0) -> np.ndarray:
        return _coverage(self.ctx, None, self, seqs, offs, 13, cutoff)

    def positions_total(self) -> int:
        t = C.c_uint64()
        self.ctx.check(lib().aix_positions_total13(self.ctx.handle, self._h, C.byref(t)))
        return int(t.value)

    def positions_build(self, reads):
        a = _bytes(reads)
        indices = np.zeros(TOTAL_13MERS + 1, dtype=np.uint64)
        positions = np.zeros(self.positions_total(), dtype=np.uint64)
        self.ctx.check(lib().aix_positions_build13(self.ctx.handle, self._h, _p(a), a.size, _p(indices),
                                                   _p(positions)))
        return indices, positions


def _coverage(ctx, ix23, ix13, seqs, offs, k, cutoff):
    a = _bytes(seqs)
    if offs is None:
        offs = np.array([0, a.size], dtype=np.int64)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    lens = np.diff(offs)
    total = int(np.maximum(lens - (k - 1), 0).sum())
    out = np.zeros(total, dtype=np.uint32)
    ctx.check(lib().aix_coverage(ctx.handle, ix23._h if ix23 else None, ix13._h if ix13 else None, _p(a),
                                 _p(offs), offs.size - 1, k, cutoff, _p(out)))
    return out


class Positions:
    """aix_positions: .indices.bin / .index.bin resident in HBM."""

    def __init__(self, ctx: Context, indices=None, positions=None, handle=None):
        self.ctx = ctx
        if handle is not None:
            self._h = handle
            return
        indices = np.ascontiguousarray(indices, dtype=np.uint64)
        positions = np.ascontiguousarray(positions, dtype=np.uint64)
        self._h = C.c_void_p()
        ctx.check(lib().aix_positions_upload(ctx.handle, _p(indices), indices.size, _p(positions),
                                             positions.size, C.byref(self._h)))

    @classmethod
    def build_dev(cls, index, reads_ptr: int, n_bytes: int, k: int) -> "Positions":
        """Build from a .reads image already in HBM (readable 8 bytes past n_bytes); stays in HBM."""
        h = C.c_void_p()
        fn = lib().aix_positions_build23_dev if k == 23 else lib().aix_positions_build13_dev
        index.ctx.check(fn(index.ctx.handle, index._h, reads_ptr, n_bytes, C.byref(h)))
        return cls(index.ctx, handle=h)

    @property
    def info(self):
        a = (C.c_uint64 * 2)()
        self.ctx.check(lib().aix_positions_info(self._h, a))
        return {"n_indices": int(a[0]), "n_positions": int(a[1])}

    def device_arrays(self):
        """-> (indices_ptr, positions_ptr) device addresses of the u64 arrays."""
        a, b = C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().aix_positions_arrays_dev(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def download(self):
        """-> (indices uint64[n+1], positions uint64[total]) = .indices.bin / .index.bin contents."""
        inf = self.info
        indices = np.zeros(inf["n_indices"], dtype=np.uint64)
        positions = np.zeros(inf["n_positions"], dtype=np.uint64)
        self.ctx.check(lib().aix_positions_download(self.ctx.handle, self._h, _p(indices), _p(positions)))
        return indices, positions

    def close(self):
        if self._h:
            lib().aix_positions_destroy(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, index, kmers, k: int):
        """-> (offsets uint64[q+1], positions uint64[total]), 0-based positions per query."""
        recs, lens = as_records(kmers)
        q = recs.shape[0]
        ix23 = index._h if k == 23 else None
        ix13 = index._h if k == 13 else None
        counts = np.zeros(q, dtype=np.uint64)
        self.ctx.check(lib().aix_positions_query(self.ctx.handle, ix23, ix13, self._h, _p(recs), recs.shape[1],
                                                 _p(lens), q, k, _p(counts), None, None))
        offs = np.zeros(q + 1, dtype=np.uint64)
        np.cumsum(counts, out=offs[1:])
        out = np.zeros(int(offs[-1]), dtype=np.uint64)
        if out.size:
            self.ctx.check(lib().aix_positions_query(self.ctx.handle, ix23, ix13, self._h, _p(recs), recs.shape[1],
                                                     _p(lens), q, k, None, _p(offs), _p(out)))
        return offs, out
