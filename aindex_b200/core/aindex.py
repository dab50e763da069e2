"""AIndex -- the Python API of ad3002/aindex (aindex/core/aindex.py) on the CUDA backend.

Same public surface as the reference class (load_from_prefix, item access, single / batch tf
queries, sequence coverage, positions, read access, iterators); the work is done by
aindex_cpp.AindexWrapper, i.e. by libaindex_cuda on the GPU.  Differences from the reference
are limited to its documented defects:
  * load_from_prefix auto-detection looks at {prefix}.kmers.bin to tell 23-mer from 13-mer
    indexes (the reference tests the same two files for both and always answers 13);
  * get_tf_values_13mer is actually bound;
  * get_sequence_coverage is one batched device call instead of one Python->C++ call per
    position (aindex.py:314-322) -- results are identical;
  * intervaltree / editdistance / Bio are imported lazily (only load_reads_index and
    get_kmer_info needed them).
"""
from __future__ import annotations

import logging
import os
from collections import defaultdict
from enum import IntEnum
from typing import Dict, Iterator, List, Optional, Tuple

try:
    from . import aindex_cpp
except ImportError as exc:  # pragma: no cover - the extension is mandatory
    raise ImportError("aindex_b200.core.aindex_cpp is not built (python -m aindex_b200.build); "
                      "there is no pure-Python or CPU fallback") from exc

log = logging.getLogger(__name__)

_COMP = str.maketrans("ATCGNatcgn~[]", "TAGCNtagcn~][")


class Strand(IntEnum):
    NOT_FOUND = 0
    FORWARD = 1
    REVERSE = 2


def get_revcomp(sequence: str) -> str:
    """Reverse complement of a read ('~' and brackets kept, as aindex.py:35-43).

    >>> get_revcomp('ATCGN')
    'NCGAT'
    """
    return sequence.translate(_COMP)[::-1]


def hamming_distance(s1: str, s2: str) -> int:
    """Mismatches between two strings, positions holding 'N' ignored (aindex.py:45-47)."""
    return sum(1 for a, b in zip(s1, s2) if a != b and a != "N" and b != "N")


class AIndex:
    """K-mer index over reads: term frequencies, positions, coverage."""

    def __init__(self):
        self._wrapper = aindex_cpp.AindexWrapper()
        self._loaded = False
        self.reads_size = 0
        self.loaded_header = False
        self.loaded_intervals = False
        self.loaded_reads = False
        self.max_tf = 0
        self.k = 23

    # ------------------------------------------------------------------ loading
    def load_hash(self, hash_file: str, tf_file: str, kmers_bin_file: str, kmers_text_file: str = ""):
        for what, path in (("hash", hash_file), ("tf", tf_file), ("kmers_bin", kmers_bin_file)):
            if not os.path.exists(path):
                raise FileNotFoundError(f"{what} file not found: {path}")
        if kmers_text_file and not os.path.exists(kmers_text_file):
            raise FileNotFoundError(f"kmers_text file not found: {kmers_text_file}")
        self._wrapper.load(hash_file, tf_file, kmers_bin_file, kmers_text_file)
        self._loaded, self.k = True, 23

    load_hash_file = load_hash

    def load_reads(self, reads_file: str):
        if not os.path.exists(reads_file):
            raise FileNotFoundError(f"Reads file not found: {reads_file}")
        self._wrapper.load_reads(reads_file)
        self.reads_size = self._wrapper.reads_size
        self.loaded_reads = True

    def load_aindex(self, index_file: str, indices_file: str, max_tf: int):
        for what, path in (("index", index_file), ("indices", indices_file)):
            if not os.path.exists(path):
                raise FileNotFoundError(f"{what} file not found: {path}")
        self._wrapper.load_aindex(index_file, indices_file, max_tf)
        self.max_tf = max_tf

    def load_reads_index(self, index_file: str, header_file: Optional[str] = None):
        """rid -> (start, end) map and an interval tree over read (and header) spans."""
        from intervaltree import IntervalTree  # optional dependency, as in the reference
        self.rid2start, self.IT, self.chrm2start, self.headers = {}, IntervalTree(), {}, {}
        with open(index_file) as fh:
            for line in fh:
                rid, start, end = (int(x) for x in line.split("\t")[:3])
                self.rid2start[rid] = (start, end)
                self.IT.addi(start, end, rid)
        self.loaded_intervals = True
        if header_file:
            with open(header_file) as fh:
                for rid, line in enumerate(fh):
                    head, start, length = line.rstrip("\n").split("\t")
                    self.headers[rid] = head
                    self.chrm2start[head.split()[0].split(".")[0]] = int(start)
                    self.IT.addi(int(start), int(start) + int(length), head)
            self.loaded_header = True

    def load_13mer_index(self, hash_file: str, tf_file: str):
        for what, path in (("hash", hash_file), ("tf", tf_file)):
            if not os.path.exists(path):
                raise FileNotFoundError(f"13-mer {what} file not found: {path}")
        self._wrapper.load_13mer_index(hash_file, tf_file)
        self._loaded, self.k = True, 13

    def load_13mer_aindex(self, index_file: str, indices_file: str):
        for what, path in (("index", index_file), ("indices", indices_file)):
            if not os.path.exists(path):
                raise FileNotFoundError(f"13-mer {what} file not found: {path}")
        self._wrapper.load_13mer_aindex(index_file, indices_file)

    @staticmethod
    def load_13mer_index_static(hash_file: str, tf_file: str) -> "AIndex":
        ix = AIndex()
        ix.load_13mer_index(hash_file, tf_file)
        return ix

    @staticmethod
    def load_23mer_index(hash_file: str, tf_file: str, kmers_bin_file: str, kmers_text_file: str = "") -> "AIndex":
        ix = AIndex()
        ix.load_hash(hash_file, tf_file, kmers_bin_file, kmers_text_file)
        return ix

    @staticmethod
    def load_from_prefix(prefix: str, kmer_size: Optional[int] = None, max_tf: int = 100000,
                         load_aindex: bool = True, load_reads: bool = False) -> "AIndex":
        """Load {prefix}.pf/.tf.bin[/.kmers.bin] (+ .index.bin/.indices.bin, + reads)."""
        if kmer_size is None:
            have = {ext: os.path.exists(prefix + ext) for ext in (".pf", ".tf.bin", ".kmers.bin")}
            if have[".pf"] and have[".tf.bin"] and have[".kmers.bin"]:
                kmer_size = 23
            elif have[".pf"] and have[".tf.bin"]:
                kmer_size = 13
            else:
                raise FileNotFoundError(f"Could not auto-detect k-mer size for prefix '{prefix}': expected "
                                        f"{prefix}.pf + {prefix}.tf.bin (13-mers) [+ {prefix}.kmers.bin (23-mers)]")
        if kmer_size not in (13, 23):
            raise ValueError(f"Unsupported kmer size: {kmer_size}. Only 13 and 23 are supported.")
        reads_file = ""
        if load_reads:
            for cand in (prefix + ".reads", (prefix + ".reads").replace(".23.", ".").replace(".13.", ".")):
                if os.path.exists(cand):
                    reads_file = cand
                    break
            else:
                log.warning("Reads file not found for prefix %s", prefix)
        ix = AIndex()
        if kmer_size == 13:
            ix.load_from_prefix_13mer(prefix, load_aindex=load_aindex, reads_file=reads_file)
        else:
            ix.load_from_prefix_23mer(prefix, max_tf=max_tf if max_tf is not None else 100000,
                                      load_aindex=load_aindex, reads_file=reads_file)
        return ix

    def load_from_prefix_23mer(self, prefix: str, max_tf: int = 100, load_aindex: bool = True, reads_file: str = ""):
        self._wrapper.load_from_prefix_23mer(prefix, reads_file)
        self._loaded, self.k = True, 23
        if reads_file:
            self.reads_size, self.loaded_reads = self._wrapper.reads_size, True
        if load_aindex:
            try:
                self._wrapper.load_aindex_from_prefix_23mer(prefix, max_tf, reads_file)
                self.max_tf = max_tf
            except (FileNotFoundError, RuntimeError) as e:
                log.warning("Could not load 23-mer AIndex from prefix %s: %s", prefix, e)

    def load_from_prefix_13mer(self, prefix: str, load_aindex: bool = True, reads_file: str = ""):
        self._wrapper.load_from_prefix_13mer(prefix, reads_file)
        self._loaded, self.k = True, 13
        if reads_file:
            self.reads_size, self.loaded_reads = self._wrapper.reads_size, True
        if load_aindex:
            try:
                self._wrapper.load_aindex_from_prefix_13mer(prefix, reads_file)
            except (FileNotFoundError, RuntimeError) as e:
                log.warning("Could not load 13-mer AIndex from prefix %s: %s", prefix, e)

    # ------------------------------------------------------------------ tf queries
    def get_tf_value(self, kmer: str) -> int:
        return self._wrapper.get_tf_value(kmer) if self._loaded else 0

    def get_tf_values(self, kmers):
        """list[str] -> list[int]; uint8[q, k] ndarray -> uint32 ndarray (zero-copy path)."""
        if not self._loaded:
            return [0] * len(kmers)
        return self._wrapper.get_tf_values(kmers)

    def get_tf_values_13mer(self, kmers: List[str]) -> List[int]:
        return self._wrapper.get_tf_values_13mer(kmers) if self._loaded else [0] * len(kmers)

    def get_total_tf_value_13mer(self, kmer: str) -> int:
        return self._wrapper.get_total_tf_value_13mer(kmer)

    def get_total_tf_values_13mer(self, kmers: List[str]) -> List[int]:
        return self._wrapper.get_total_tf_values_13mer(kmers)

    def get_tf_both_directions_13mer(self, kmer: str) -> Tuple[int, int]:
        return self._wrapper.get_tf_both_directions_13mer(kmer)

    def get_tf_both_directions_13mer_batch(self, kmers: List[str]) -> List[Tuple[int, int]]:
        return self._wrapper.get_tf_both_directions_13mer_batch(kmers)

    def get_total_tf_value_23mer(self, kmer: str) -> int:
        return self._wrapper.get_total_tf_value_23mer(kmer)

    def get_total_tf_values_23mer(self, kmers: List[str]) -> List[int]:
        return self._wrapper.get_total_tf_values_23mer(kmers)

    def get_tf_both_directions_23mer(self, kmer: str) -> Tuple[int, int]:
        return self._wrapper.get_tf_both_directions_23mer(kmer)

    def get_tf_both_directions_23mer_batch(self, kmers: List[str]) -> List[Tuple[int, int]]:
        return self._wrapper.get_tf_both_directions_23mer_batch(kmers)

    def _need_index(self):
        if not self._loaded:
            raise RuntimeError("Index not loaded")

    def get_hash_value(self, kmer: str) -> int:
        self._need_index()
        return self._wrapper.get_hash_value(kmer)

    def get_hash_values(self, kmers: List[str]) -> List[int]:
        self._need_index()
        return self._wrapper.get_hash_values(kmers)

    def get_kid_by_kmer(self, kmer: str) -> int:
        self._need_index()
        return self._wrapper.get_kid_by_kmer(kmer)

    def get_kmer_by_kid(self, kid: int) -> str:
        self._need_index()
        return self._wrapper.get_kmer_by_kid(kid)

    def get_strand(self, kmer: str) -> Strand:
        self._need_index()
        return Strand(self._wrapper.get_strand(kmer))

    def get_kmer_info(self, kid: int) -> Tuple[str, str, int]:
        """(kmer, reverse complement, tf) of a k-mer id."""
        self._need_index()
        tf, kmer, rkmer = self._wrapper.get_kmer_info(kid)
        return kmer, rkmer, tf

    def get_kmer_info_by_kid(self, kid: int, k: int = 23):
        return self.get_kmer_info(kid)

    # ------------------------------------------------------------------ reads / positions
    def _need_aindex(self):
        if not self._wrapper.aindex_loaded:
            raise RuntimeError("Aindex not loaded")

    def get_reads_by_kmer(self, kmer: str, max_reads: int = 100) -> List[str]:
        self._need_aindex()
        return self._wrapper.get_reads_se_by_kmer(kmer, max_reads)

    def get_read_by_rid(self, rid: int) -> str:
        return self._wrapper.get_read_by_rid(rid)

    def get_read(self, start: int, end: int, revcomp: bool = False) -> str:
        return self._wrapper.get_read(start, end, revcomp)

    def get_rid(self, pos: int) -> int:
        self._need_aindex()
        return self._wrapper.get_rid(pos)

    def get_start(self, pos: int) -> int:
        self._need_aindex()
        return self._wrapper.get_start(pos)

    def get_positions(self, kmer: str) -> List[int]:
        if len(kmer) == 13:
            return self._wrapper.get_positions(kmer)
        if len(kmer) == 23:
            if not self._wrapper.aindex_loaded:
                raise RuntimeError("23-mer Aindex not loaded")
            return self._wrapper.get_positions(kmer)
        raise ValueError(f"Unsupported k-mer length: {len(kmer)}. Only 13-mers and 23-mers are supported.")

    def get_positions_13mer(self, kmer: str) -> List[int]:
        return self._wrapper.get_positions_13mer(kmer)

    def get_positions_batch(self, kmers: List[str], k: int = 23):
        """-> (offsets uint64[q+1], positions uint64[total]) in one device call."""
        return self._wrapper.get_positions_batch(kmers, k)

    pos = get_positions

    def get_rid2poses(self, kmer: str) -> Dict[int, List[int]]:
        hits = defaultdict(list)
        for p in self.pos(kmer):
            hits[self.get_rid(p)].append(p - self.get_start(p))
        return hits

    def get_header(self, pos: int):
        if not self.loaded_header:
            return None
        found = self.IT[pos]
        return self.headers.get(next(iter(found)).data, "") if found else ""

    def get_hash_size(self) -> int:
        self._need_index()
        return self._wrapper.get_hash_size()

    def get_reads_size(self) -> int:
        return self._wrapper.get_reads_size()

    def __len__(self) -> int:
        return self.get_hash_size()

    def __getitem__(self, kmer: str) -> int:
        return self.get_tf_value(kmer)

    def __contains__(self, kmer: str) -> bool:
        return self[kmer] > 0

    def get(self, kmer: str, default: int = 0) -> int:
        tf = self[kmer]
        return tf if tf > 0 else default

    @property
    def n_reads(self) -> int:
        return self._wrapper.n_reads

    @property
    def n_kmers(self) -> int:
        return self._wrapper.n_kmers

    @property
    def aindex_loaded(self) -> bool:
        return self._wrapper.aindex_loaded

    def iter_reads(self):
        if self.reads_size == 0:
            raise RuntimeError("Reads were not loaded.")
        for rid in range(self.n_reads):
            yield rid, self.get_read_by_rid(rid)

    def iter_reads_se(self):
        for rid, read in self.iter_reads():
            for idx, sub in enumerate(read.split("~")):
                yield rid, idx, sub

    # ------------------------------------------------------------------ coverage
    def iter_sequence_kmers(self, sequence: str, k: int = 23):
        """(kmer, tf) for every window without a line / pair separator (aindex.py:306-312)."""
        n = len(sequence) - k + 1
        if n <= 0:
            return
        tfs = self._wrapper.get_sequence_coverage(sequence, 0, k)
        for i in range(n):
            kmer = sequence[i:i + k]
            if "\n" in kmer or "~" in kmer:
                continue
            yield kmer, int(tfs[i])

    def get_sequence_coverage(self, seq: str, cutoff: int = 0, k: int = 23) -> list:
        """coverage[i] = tf(seq[i:i+k]) if tf >= cutoff else 0 (aindex.py:314-322), one device call."""
        if not self._loaded or len(seq) < k:
            return [0] * max(0, len(seq) - k + 1)
        return self._wrapper.get_sequence_coverage(seq, cutoff, k).tolist()

    def get_sequence_coverage_batch(self, seqs: List[str], cutoff: int = 0, k: int = 23):
        """Coverage of many sequences at once -> (offsets int64[n+1], uint32 ndarray)."""
        import numpy as np
        raw = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
        offs = np.zeros(len(raw) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in raw], out=offs[1:])
        cov = self._wrapper.get_sequence_coverage_batch(b"".join(raw), offs, cutoff, k)
        out_offs = np.zeros(len(raw) + 1, dtype=np.int64)
        np.cumsum(np.maximum(np.diff(offs) - (k - 1), 0), out=out_offs[1:])
        return out_offs, cov

    def print_sequence_coverage(self, seq: str, cutoff: int = 0):
        cov = self.get_sequence_coverage(seq, cutoff)
        for i, tf in enumerate(cov):
            print(f"{i}\t{seq[i:i + 23]}\t{tf}")
        return cov

    # ------------------------------------------------------------------ 13-mer array access / stats
    def get_13mer_tf_array(self) -> List[int]:
        return self._wrapper.get_13mer_tf_array()

    def get_tf_by_index_13mer(self, index: int) -> int:
        return self._wrapper.get_tf_by_index_13mer(index)

    def get_index_info(self) -> str:
        return self._wrapper.get_index_info()

    def get_13mer_statistics(self) -> Dict[str, int]:
        return self._wrapper.get_13mer_statistics()

    def get_top_kmers(self, n: int = 100, min_tf: int = 1, kmer_type: str = "auto") -> List[Tuple[str, int]]:
        """The n most frequent k-mers as (kmer, tf), most frequent first (aindex.py:683-701)."""
        return list(self.iter_kmers_by_frequency(min_tf=min_tf, max_kmers=n, kmer_type=kmer_type))

    def get_kmer_frequency_stats(self, kmer_type: str = "auto") -> dict:
        """Summary of the tf array (aindex.py:703-795): same keys, computed with numpy instead of Python loops."""
        import numpy as np
        if not self._loaded:
            raise RuntimeError("Index not loaded")
        if kmer_type == "auto":
            kmer_type = "13mer" if self.k == 13 else "23mer"
        if kmer_type == "13mer":
            tf = np.asarray(self._wrapper.get_13mer_tf_array(), dtype=np.uint64)
        elif kmer_type == "23mer":
            if self.get_hash_size() == 0:
                raise RuntimeError("23-mer index not properly loaded")
            tf = np.asarray(self._wrapper.get_tf_array_23mer(), dtype=np.uint64)
        else:
            raise ValueError(f"Unsupported kmer_type: {kmer_type}. Use '13mer', '23mer', or 'auto'")
        nz = tf[tf > 0]
        total = int(tf.size)
        return {
            "kmer_type": kmer_type, "total_kmers": total, "non_zero_kmers": int(nz.size), "zero_kmers": total - int(nz.size),
            "max_tf": int(nz.max()) if nz.size else 0, "min_tf": int(nz.min()) if nz.size else 0,
            "avg_tf": float(nz.sum() / nz.size) if nz.size else 0, "total_tf": int(tf.sum()) if nz.size else 0,
            "coverage": nz.size / total if total else 0,
        }

    def get_23mer_statistics(self) -> str:
        return self._wrapper.get_23mer_statistics()

    @staticmethod
    def _index_to_13mer(index: int) -> str:
        return "".join("ACGT"[(index >> (2 * (12 - j))) & 3] for j in range(13))

    def iter_kmers_by_frequency(self, min_tf: int = 1, max_kmers: Optional[int] = None,
                                kmer_type: str = "auto") -> Iterator[Tuple[str, int]]:
        """k-mers by decreasing tf (13-mer mode: forward-strand counts of the tf array)."""
        import numpy as np
        if kmer_type == "auto":
            kmer_type = "13mer" if self.k == 13 else "23mer"
        if kmer_type == "13mer":
            # the tf file is in MPHF order; the direct-address copy (tf by 2-bit value) is what makes
            # _index_to_13mer(index) name the right k-mer (the reference labels MPHF ids as if they were values)
            tf = np.asarray(self._wrapper.get_13mer_tf_array_direct())
            keep = np.flatnonzero(tf >= max(int(min_tf), 0))
            order = keep[np.argsort(-tf[keep].astype(np.int64), kind="stable")]
            if max_kmers is not None:
                order = order[:max_kmers]
            for v in order:
                yield self._index_to_13mer(int(v)), int(tf[v])
            return
        tfs = np.asarray(self._wrapper.get_tf_array_23mer(), dtype=np.uint64)
        order = np.argsort(-tfs.astype(np.int64), kind="stable")
        emitted = 0
        for kid in order:
            if tfs[kid] < min_tf or (max_kmers is not None and emitted >= max_kmers):
                break
            yield self._wrapper.get_kmer_by_kid(int(kid)), int(tfs[kid])
            emitted += 1
