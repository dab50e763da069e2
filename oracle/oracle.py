"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(aindex_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "bin")
PF13_PATH = os.path.join(REF_DIR, "data", "all_13mers.pf")
TOTAL_13MERS = 1 << 26


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "aindex_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class _Mphf(C.Structure):
    _fields_ = [("n", C.c_uint64), ("hash_domain", C.c_uint64), ("seed", C.c_uint64),
                ("bv_size", C.c_uint64), ("n_words", C.c_uint64), ("n_blocks", C.c_uint64),
                ("words", C.POINTER(C.c_uint64)), ("block_ranks", C.POINTER(C.c_uint64))]


class _Index23(C.Structure):
    _fields_ = [("mphf", C.POINTER(_Mphf)), ("checker", C.c_void_p), ("tf", C.c_void_p),
                ("n", C.c_uint64)]


class CountStats(C.Structure):
    _fields_ = [("sequences", C.c_uint64), ("windows", C.c_uint64), ("valid", C.c_uint64),
                ("invalid", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        u64, u8p, vp = C.c_uint64, C.c_void_p, C.c_void_p
        L.orc_jenkins64.argtypes = [u64, u8p, u64, vp]
        L.orc_mphf_load.restype = C.POINTER(_Mphf)
        L.orc_mphf_load.argtypes = [C.c_char_p]
        L.orc_mphf_from_arrays.restype = C.POINTER(_Mphf)
        L.orc_mphf_from_arrays.argtypes = [u64, u64, u64, vp, u64, vp, u64]
        L.orc_mphf_save.argtypes = [C.POINTER(_Mphf), C.c_char_p]
        L.orc_mphf_free.argtypes = [C.POINTER(_Mphf)]
        L.orc_mphf_lookup.restype = u64
        L.orc_mphf_lookup.argtypes = [C.POINTER(_Mphf), u8p, u64]
        L.orc_mphf_lookup_batch.argtypes = [C.POINTER(_Mphf), vp, u64, vp, u64, vp, C.c_int]
        L.orc_dna23_bitset.restype = u64
        L.orc_dna23_bitset.argtypes = [u8p, u64]
        L.orc_dna13_bitset.restype = C.c_uint32
        L.orc_dna13_bitset.argtypes = [u8p, u64]
        L.orc_bitset_dna23.argtypes = [u64, vp, C.c_int]
        L.orc_bitset_dna13.argtypes = [C.c_uint32, vp, C.c_int]
        L.orc_reverse_dna23.restype = u64
        L.orc_reverse_dna23.argtypes = [u64]
        L.orc_reverse_dna13.restype = C.c_uint32
        L.orc_reverse_dna13.argtypes = [C.c_uint32]
        L.orc_dna_bitset_pack.argtypes = [vp, u64, vp]
        L.orc_dna_bitset_ukmer.restype = u64
        L.orc_dna_bitset_ukmer.argtypes = [vp, u64, C.c_int]
        L.orc_tf23_batch.argtypes = [C.POINTER(_Index23), vp, u64, vp, u64, C.c_int, vp, C.c_int]
        L.orc_get_freq23.restype = C.c_uint32
        L.orc_get_freq23.argtypes = [C.POINTER(_Index23), u64]
        L.orc_tf13_batch.argtypes = [C.POINTER(_Mphf), vp, vp, u64, vp, u64, C.c_int, vp, C.c_int]
        L.orc_detect_format.restype = C.c_int
        L.orc_detect_format.argtypes = [vp, u64]
        L.orc_count13.argtypes = [C.POINTER(_Mphf), vp, u64, C.c_int, vp, C.POINTER(CountStats)]
        L.orc_count13_direct.argtypes = [vp, u64, C.c_int, vp, C.POINTER(CountStats)]
        L.orc_coverage23.argtypes = [C.POINTER(_Index23), vp, u64, C.c_uint32, vp]
        L.orc_coverage13.argtypes = [C.POINTER(_Mphf), vp, vp, u64, C.c_uint32, vp]
        L.orc_positions_build23.argtypes = [C.POINTER(_Index23), vp, u64, vp, vp]
        L.orc_positions_build13.argtypes = [C.POINTER(_Mphf), vp, vp, u64, vp, vp]
        L.orc_positions_query23.restype = u64
        L.orc_positions_query23.argtypes = [C.POINTER(_Index23), vp, vp, vp, u64, vp, u64]
        L.orc_positions_query13.restype = u64
        L.orc_positions_query13.argtypes = [C.POINTER(_Mphf), vp, vp, u64, vp, u64, vp, u64]
        L.orc_canonical23_count.restype = u64
        L.orc_canonical23_count.argtypes = [vp, u64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orc_free.argtypes = [vp]
        L.orc_write_dat.restype = C.c_int
        L.orc_write_dat.argtypes = [vp, vp, u64, C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bytes_arr(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    if isinstance(b, str):
        b = b.encode("latin-1")
    return np.frombuffer(bytes(b), dtype=np.uint8)


def pack_queries(kmers, stride=None):
    """list[str|bytes] -> (uint8[q, stride] zero padded, uint8[q] lens)."""
    bs = [k.encode("latin-1") if isinstance(k, str) else bytes(k) for k in kmers]
    if stride is None:
        stride = max([len(b) for b in bs] + [1])
    recs = np.zeros((len(bs), stride), dtype=np.uint8)
    lens = np.zeros(len(bs), dtype=np.uint8)
    for i, b in enumerate(bs):
        if len(b) > stride or len(b) > 255:
            raise ValueError("query longer than stride")
        recs[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        lens[i] = len(b)
    return recs, lens


def jenkins64(seed: int, s) -> tuple:
    a = _bytes_arr(s)
    out = np.zeros(3, dtype=np.uint64)
    lib().orc_jenkins64(seed, _ptr(a), a.size, _ptr(out))
    return tuple(int(x) for x in out)


class Mphf:
    """emphf::mphf<jenkins64_hasher> (src/emphf/mphf.hpp) -- lookup side."""

    def __init__(self, handle):
        if not handle:
            raise FileNotFoundError("could not load .pf")
        self._h = handle

    @classmethod
    def load(cls, path: str) -> "Mphf":
        return cls(lib().orc_mphf_load(os.fsencode(path)))

    @classmethod
    def from_arrays(cls, n, hash_domain, seed, words, block_ranks) -> "Mphf":
        words = np.ascontiguousarray(words, dtype=np.uint64)
        block_ranks = np.ascontiguousarray(block_ranks, dtype=np.uint64)
        return cls(lib().orc_mphf_from_arrays(n, hash_domain, seed, _ptr(words), words.size,
                                              _ptr(block_ranks), block_ranks.size))

    def save(self, path: str):
        if lib().orc_mphf_save(self._h, os.fsencode(path)) != 0:
            raise OSError("cannot write " + path)

    def __del__(self):
        try:
            lib().orc_mphf_free(self._h)
        except Exception:
            pass

    n = property(lambda s: int(s._h.contents.n))
    hash_domain = property(lambda s: int(s._h.contents.hash_domain))
    seed = property(lambda s: int(s._h.contents.seed))
    bv_size = property(lambda s: int(s._h.contents.bv_size))
    n_words = property(lambda s: int(s._h.contents.n_words))
    n_blocks = property(lambda s: int(s._h.contents.n_blocks))

    @property
    def words(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._h.contents.words, shape=(self.n_words,)).copy()

    @property
    def block_ranks(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._h.contents.block_ranks, shape=(self.n_blocks,)).copy()

    def lookup(self, s) -> int:
        a = _bytes_arr(s)
        return int(lib().orc_mphf_lookup(self._h, _ptr(a), a.size))

    def lookup_batch(self, recs: np.ndarray, lens=None, threads: int = 1) -> np.ndarray:
        recs = np.ascontiguousarray(recs, dtype=np.uint8)
        q, stride = recs.shape
        out = np.zeros(q, dtype=np.uint64)
        lib().orc_mphf_lookup_batch(self._h, _ptr(recs), stride, _ptr(lens), q, _ptr(out), threads)
        return out


def dna23_bitset(s) -> int:
    a = _bytes_arr(s)
    return int(lib().orc_dna23_bitset(_ptr(a), a.size))


def dna13_bitset(s) -> int:
    a = _bytes_arr(s)
    return int(lib().orc_dna13_bitset(_ptr(a), a.size))


def bitset_dna23(x: int, k: int = 23) -> str:
    out = np.zeros(k, dtype=np.uint8)
    lib().orc_bitset_dna23(x, _ptr(out), k)
    return out.tobytes().decode()


def bitset_dna13(x: int, k: int = 13) -> str:
    out = np.zeros(k, dtype=np.uint8)
    lib().orc_bitset_dna13(x, _ptr(out), k)
    return out.tobytes().decode()


def reverse_dna23(x: int) -> int:
    return int(lib().orc_reverse_dna23(x))


def reverse_dna13(x: int) -> int:
    return int(lib().orc_reverse_dna13(x))


def dna_bitset_pack(s) -> np.ndarray:
    a = _bytes_arr(s)
    out = np.zeros((a.size + 3) // 4, dtype=np.uint8)
    lib().orc_dna_bitset_pack(_ptr(a), a.size, _ptr(out))
    return out


def dna_bitset_ukmer(packed: np.ndarray, pos: int, k: int) -> int:
    return int(lib().orc_dna_bitset_ukmer(_ptr(packed), pos, k))


MODE_TF, MODE_TOTAL, MODE_BOTH, MODE_PFID, MODE_STRAND, MODE_KID = range(6)


class Index23:
    """PHASH_MAP (src/hash.hpp:82-353) + AindexWrapper 23-mer queries."""

    def __init__(self, mphf: Mphf, checker: np.ndarray, tf: np.ndarray):
        self.mphf = mphf
        self.checker = np.ascontiguousarray(checker, dtype=np.uint64)
        self.tf = np.ascontiguousarray(tf, dtype=np.uint32)
        assert self.checker.size == self.tf.size
        self.n = int(self.checker.size)
        self._s = _Index23(mphf._h, self.checker.ctypes.data, self.tf.ctypes.data, self.n)

    @classmethod
    def load_prefix(cls, prefix: str) -> "Index23":
        return cls(Mphf.load(prefix + ".pf"), np.fromfile(prefix + ".kmers.bin", dtype=np.uint64),
                   np.fromfile(prefix + ".tf.bin", dtype=np.uint32))

    def batch(self, recs: np.ndarray, lens=None, mode: int = MODE_TF, threads: int = 1):
        recs = np.ascontiguousarray(recs, dtype=np.uint8)
        q, stride = recs.shape
        if mode == MODE_TF:
            out = np.zeros(q, dtype=np.uint32)
        elif mode == MODE_BOTH:
            out = np.zeros((q, 2), dtype=np.uint32)
        else:
            out = np.zeros(q, dtype=np.uint64)
        lib().orc_tf23_batch(C.byref(self._s), _ptr(recs), stride, _ptr(lens), q, mode, _ptr(out),
                             threads)
        return out

    def query(self, kmers, mode: int = MODE_TF):
        recs, lens = pack_queries(kmers)
        return self.batch(recs, lens, mode)

    def get_freq(self, ukmer: int) -> int:
        return int(lib().orc_get_freq23(C.byref(self._s), ukmer))

    def coverage(self, seq, cutoff: int = 0) -> np.ndarray:
        a = _bytes_arr(seq)
        out = np.zeros(max(0, a.size - 22), dtype=np.uint32)
        lib().orc_coverage23(C.byref(self._s), _ptr(a), a.size, cutoff, _ptr(out))
        return out

    def positions_build(self, reads) -> tuple:
        a = _bytes_arr(reads)
        indices = np.zeros(self.n + 1, dtype=np.uint64)
        total = int(self.tf.astype(np.uint64).sum())
        positions = np.zeros(total, dtype=np.uint64)
        lib().orc_positions_build23(C.byref(self._s), _ptr(a), a.size, _ptr(indices), _ptr(positions))
        return indices, positions

    def positions_query(self, indices, positions, kmer) -> np.ndarray:
        a = _bytes_arr(kmer)
        cap = int(self.tf.max()) if self.n else 0
        out = np.zeros(max(cap, 1), dtype=np.uint64)
        c = lib().orc_positions_query23(C.byref(self._s), _ptr(indices), _ptr(positions), _ptr(a),
                                        a.size, _ptr(out), cap)
        return out[:int(c)].copy()


class Index13:
    """13-mer mode of AindexWrapper (python_wrapper.cpp:404-437, 482-608)."""

    def __init__(self, mphf: Mphf, tf64: np.ndarray):
        self.mphf = mphf
        self.tf64 = np.ascontiguousarray(tf64, dtype=np.uint64)
        assert self.tf64.size == TOTAL_13MERS

    def batch(self, recs: np.ndarray, lens=None, mode: int = MODE_TF, threads: int = 1):
        recs = np.ascontiguousarray(recs, dtype=np.uint8)
        q, stride = recs.shape
        if mode == MODE_TF:
            out = np.zeros(q, dtype=np.uint32)
        elif mode == MODE_BOTH:
            out = np.zeros((q, 2), dtype=np.uint64)
        else:
            out = np.zeros(q, dtype=np.uint64)
        lib().orc_tf13_batch(self.mphf._h, _ptr(self.tf64), _ptr(recs), stride, _ptr(lens), q, mode,
                             _ptr(out), threads)
        return out

    def query(self, kmers, mode: int = MODE_TF):
        recs, lens = pack_queries(kmers)
        return self.batch(recs, lens, mode)

    def coverage(self, seq, cutoff: int = 0) -> np.ndarray:
        a = _bytes_arr(seq)
        out = np.zeros(max(0, a.size - 12), dtype=np.uint32)
        lib().orc_coverage13(self.mphf._h, _ptr(self.tf64), _ptr(a), a.size, cutoff, _ptr(out))
        return out

    def positions_build(self, reads) -> tuple:
        a = _bytes_arr(reads)
        indices = np.zeros(TOTAL_13MERS + 1, dtype=np.uint64)
        positions = np.zeros(int(self.tf64.sum()), dtype=np.uint64)
        lib().orc_positions_build13(self.mphf._h, _ptr(self.tf64), _ptr(a), a.size, _ptr(indices),
                                    _ptr(positions))
        return indices, positions

    def positions_query(self, indices, positions, kmer) -> np.ndarray:
        a = _bytes_arr(kmer)
        cap = 1 << 20
        out = np.zeros(cap, dtype=np.uint64)
        c = lib().orc_positions_query13(self.mphf._h, _ptr(indices), _ptr(positions), positions.size,
                                        _ptr(a), a.size, _ptr(out), cap)
        return out[:int(c)].copy()


FMT_DETECT, FMT_PLAIN, FMT_FASTA, FMT_FASTQ = -1, 0, 1, 2


def detect_format(data) -> int:
    a = _bytes_arr(data)
    return int(lib().orc_detect_format(_ptr(a), a.size))


def count13(mphf: Mphf, data, fmt: int = FMT_DETECT):
    """count_kmers13 (src/count_kmers13.cpp): uint64[4^13] in MPHF order + stats."""
    a = _bytes_arr(data)
    counts = np.zeros(TOTAL_13MERS, dtype=np.uint64)
    st = CountStats()
    lib().orc_count13(mphf._h, _ptr(a), a.size, fmt, _ptr(counts), C.byref(st))
    return counts, st.as_dict()


def count13_direct(data, fmt: int = FMT_DETECT):
    """Same counting semantics, direct-address order hist[2-bit value]."""
    a = _bytes_arr(data)
    hist = np.zeros(TOTAL_13MERS, dtype=np.uint64)
    st = CountStats()
    lib().orc_count13_direct(_ptr(a), a.size, fmt, _ptr(hist), C.byref(st))
    return hist, st.as_dict()


def canonical23_count(reads, threads: int = 0):
    """Sorted distinct canonical 23-mers (uint64) + counts (uint32) of a reads image
    (tests/analyze_kmers.py:25-33: ACGT-only windows, min(kmer, revcomp))."""
    a = _bytes_arr(reads)
    kp, cp = C.c_void_p(), C.c_void_p()
    n = int(lib().orc_canonical23_count(_ptr(a), a.size, threads or (os.cpu_count() or 1), C.byref(kp), C.byref(cp)))
    if not kp.value:  # image shorter than one window
        return np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint32)
    try:
        kmers = np.ctypeslib.as_array(C.cast(kp, C.POINTER(C.c_uint64)), shape=(max(n, 1),))[:n].copy()
        counts = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_uint32)), shape=(max(n, 1),))[:n].copy()
    finally:
        lib().orc_free(kp)
        lib().orc_free(cp)
    return kmers, counts


def write_dat(kmers, counts, dat_path=None, keys_path=None):
    """`KMER\\tCOUNT` lines (compute_index input) and / or `KMER` lines (compute_mphf_seq input)."""
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    counts = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
    rc = lib().orc_write_dat(_ptr(kmers), _ptr(counts), kmers.size, os.fsencode(dat_path) if dat_path else None,
                             os.fsencode(keys_path) if keys_path else None)
    if rc != 0:
        raise OSError("orc_write_dat failed")


def build_reference_index23(kmers, counts, prefix, threads: int = 0):
    """{prefix}.pf/.kmers.bin/.tf.bin with the UNMODIFIED reference tools (oracle/_ref/bin):
    compute_mphf_seq (emphf/compute_mphf_generic.hpp:19-61) + compute_index (compute_index.cpp:53-67).
    Returns the seconds each tool took, or None when the reference was not compiled."""
    import time
    mp, ci = os.path.join(REF_BIN, "compute_mphf_seq"), os.path.join(REF_BIN, "compute_index")
    if not (os.path.exists(mp) and os.path.exists(ci)):
        return None
    write_dat(kmers, counts, prefix + ".dat", prefix + ".kmers")
    t0 = time.perf_counter()
    subprocess.check_call([mp, prefix + ".kmers", prefix + ".pf"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t1 = time.perf_counter()
    subprocess.check_call([ci, prefix + ".dat", prefix + ".pf", prefix, str(threads or (os.cpu_count() or 1)), "0"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t2 = time.perf_counter()
    os.unlink(prefix + ".dat")
    os.unlink(prefix + ".kmers")
    return {"compute_mphf_seq_s": t1 - t0, "compute_index_s": t2 - t1}


def all_13mers_block(start: int, count: int) -> np.ndarray:
    """ASCII 13-mers for v in [start, start+count) in numeric order (uint8[count,13])."""
    v = np.arange(start, start + count, dtype=np.uint64)
    out = np.empty((count, 13), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for j in range(13):
        out[:, j] = lut[((v >> np.uint64(2 * (12 - j))) & np.uint64(3)).astype(np.int64)]
    return out


def ref_module():
    """Import the UNMODIFIED reference pybind11 module built into oracle/_ref (or None)."""
    import importlib.util
    import glob
    cands = glob.glob(os.path.join(REF_DIR, "aindex_cpp*.so"))
    if not cands:
        return None
    spec = importlib.util.spec_from_file_location("aindex_cpp", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
