#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs oracle/_ref, i.e. /root/reference compiled by
oracle/build_ref.sh).  Everything written here is produced by reference binaries /
the reference pybind11 module -- never by the oracle restatement or the CUDA path:

  idx23.{pf,kmers.bin,tf.bin,index.bin,indices.bin}, idx23.reads
      compute_mphf_seq -> compute_index -> compute_aindex (1 thread) on a seeded
      synthetic read set; the canonical 23-mer table (.dat) follows the brute-force
      definition of tests/analyze_kmers.py:25-33 (min(kmer, revcomp), ACGT-only windows).
  golden23.npz   aindex_cpp.AindexWrapper answers for a query set (tf, total, both,
                 kid, strand, hash values, positions, coverage, kmer info).
  golden13.npz   count_kmers13 output (non-zero id/count pairs + md5 of the 512 MiB
                 file + stats), 13-mer query answers, compute_aindex13 positions.

  golden_reads.npz + idx23.ridx   read access / rid / start answers (python tests/golden/make_golden.py reads)

Usage: python tests/golden/make_golden.py
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (only for paths / ref_module loader)

BIN = O.REF_BIN
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def rc(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode(errors="replace"))
        raise SystemExit(f"failed: {cmd}")
    return r.stdout.decode(errors="replace")


def make_reads23(rng):
    genome = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=6000).tobytes()
    # a low-complexity stretch so some k-mers are highly repeated
    genome = genome[:3000] + b"ACACACACACACACACACACACACACACACACACACACAC" * 3 + genome[3000:]
    lines = []
    for i in range(420):
        ln = int(rng.integers(30, 121))
        st = int(rng.integers(0, len(genome) - ln))
        r = genome[st:st + ln]
        if rng.random() < 0.5:
            r = rc(r)
        if i % 37 == 5:
            p = int(rng.integers(0, ln))
            r = r[:p] + b"N" + r[p + 1:]
        if i % 11 == 3:  # paired record: read1~revcomp(read2)  (compute_reads.cpp:89-96)
            ln2 = int(rng.integers(30, 101))
            st2 = int(rng.integers(0, len(genome) - ln2))
            r = r + b"~" + rc(genome[st2:st2 + ln2])
        lines.append(r)
    lines.append(b"ACGTACGTAC")  # shorter than k
    return genome, b"\n".join(lines) + b"\n"


def canonical_counts(reads: bytes, k=23):
    """tests/analyze_kmers.py: ACGT-only windows of each line, canonical = min(kmer, revcomp).

    The reference pipeline splits PE records at '~' implicitly (a window containing '~'
    is not ACGT-only)."""
    counts = {}
    for line in reads.split(b"\n"):
        for i in range(len(line) - k + 1):
            km = line[i:i + k]
            if km.strip(b"ACGT"):
                continue
            c = min(km, rc(km))
            counts[c] = counts.get(c, 0) + 1
    return counts


def build_idx23(tmp, rng):
    genome, reads = make_reads23(rng)
    counts = canonical_counts(reads)
    kmers = sorted(counts)  # jellyfish dump order is arbitrary; sorted is deterministic
    prefix = os.path.join(tmp, "idx23")
    with open(prefix + ".reads", "wb") as f:
        f.write(reads)
    with open(prefix + ".dat", "wb") as f:
        for km in kmers:
            f.write(km + b"\t" + str(counts[km]).encode() + b"\n")
    with open(prefix + ".kmers", "wb") as f:
        for km in kmers:
            f.write(km + b"\n")
    run([f"{BIN}/compute_mphf_seq", prefix + ".kmers", prefix + ".pf"])
    run([f"{BIN}/compute_index", prefix + ".dat", prefix + ".pf", prefix, "1", "0"])
    run([f"{BIN}/compute_aindex", prefix + ".reads", prefix + ".pf", prefix, "1", "23",
         prefix + ".tf.bin", prefix + ".kmers.bin", prefix + ".kmers"])
    for ext in (".pf", ".kmers.bin", ".tf.bin", ".index.bin", ".indices.bin", ".reads"):
        with open(prefix + ext, "rb") as src, open(os.path.join(HERE, "idx23" + ext), "wb") as dst:
            dst.write(src.read())
    return genome, reads, kmers, counts, prefix


def queries23(rng, genome, kmers):
    q = []
    pick = [kmers[i] for i in rng.choice(len(kmers), size=300, replace=False)]
    q += pick                                   # canonical, present
    q += [rc(k) for k in pick[:200]]            # non-canonical, present via reverse probe
    q += [rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=23).tobytes() for _ in range(300)]
    for k in pick[:60]:                         # one N / lowercase / junk char
        p = int(rng.integers(0, 23))
        q.append(k[:p] + b"N" + k[p + 1:])
        q.append(k.lower())
        q.append(k[:p] + k[p:p + 1].lower() + k[p + 1:])
        q.append(k[:p] + b"~" + k[p + 1:])
        q.append(rc(k)[:p] + b"N" + rc(k)[p + 1:])
    for k in pick[60:100]:                      # k-mers whose A-substituted form exists
        q.append(k.replace(b"A", b"N", 1))
        q.append(rc(k).replace(b"A", b"X", 1))
        q.append(rc(k).replace(b"A", b"a"))
    for k in pick[100:140]:                     # 22 chars (well defined: reads the NUL) and > 23
        q.append(k[:22])
        q.append(rc(k)[:22])
        q.append(k + b"A")
        q.append(rc(k) + b"ACGTACG")
        q.append(k + k + b"AC")                 # 48 chars: enters the 24-byte Jenkins loop twice
        q.append(rc(k) + genome[:24])           # 47 chars
        q.append(rc(k) + genome[:49])           # 72 chars
    q.append(b"A" * 23)
    q.append(b"T" * 23)
    q.append(b"N" * 23)
    q.append(b"ACACACACACACACACACACACA")
    q.append(b"GTGTGTGTGTGTGTGTGTGTGTG")
    return q


def golden23(rng, ref, genome, reads, kmers, counts, prefix):
    w = ref.AindexWrapper()
    w.load_from_prefix_23mer(prefix)
    w.load_aindex_from_prefix_23mer(prefix, 100000)
    q = queries23(rng, genome, kmers)
    qs = [x.decode("latin-1") for x in q]
    recs, lens = O.pack_queries(q, stride=80)
    out = {"recs": recs, "lens": lens}
    out["tf"] = np.array(w.get_tf_values(qs), dtype=np.uint32)
    out["tf_single"] = np.array([w.get_tf_value(s) for s in qs], dtype=np.uint32)
    out["total"] = np.array(w.get_total_tf_values_23mer(qs), dtype=np.uint64)
    out["both"] = np.array(w.get_tf_both_directions_23mer_batch(qs), dtype=np.uint32).reshape(-1, 2)
    out["kid"] = np.array([w.get_kid_by_kmer(s) for s in qs], dtype=np.uint64)
    out["strand"] = np.array([w.get_strand(s) for s in qs], dtype=np.uint64)
    out["hash"] = np.array(w.get_hash_values(qs), dtype=np.uint64)
    # positions: only for present 23-mers (absent ones abort the reference)
    present = [i for i, s in enumerate(q) if len(s) == 23 and out["tf"][i] > 0
               and not s.strip(b"ACGT")]
    pos_off = [0]
    pos_val = []
    for i in present:
        p = w.get_positions(qs[i])
        pos_val += list(p)
        pos_off.append(len(pos_val))
    out["pos_qidx"] = np.array(present, dtype=np.int64)
    out["pos_off"] = np.array(pos_off, dtype=np.uint64)
    out["pos_val"] = np.array(pos_val, dtype=np.uint64)
    # kid -> kmer info
    kids = np.arange(0, len(kmers), 7, dtype=np.uint64)
    info = [w.get_kmer_info(int(k)) for k in kids]
    out["info_kid"] = kids
    out["info_tf"] = np.array([t[0] for t in info], dtype=np.uint64)
    out["info_kmer"] = np.array([t[1].encode() for t in info])
    out["info_rkmer"] = np.array([t[2].encode() for t in info])
    # coverage = the aindex.py:314-322 loop over get_tf_value
    seqs = []
    for j in range(6):
        st = int(rng.integers(0, len(genome) - 400))
        s = bytearray(genome[st:st + 300 + 17 * j])
        for _ in range(4):
            p = int(rng.integers(0, len(s)))
            s[p] = b"ACGT"[int(rng.integers(0, 4))]
        if j == 2:
            s[50] = ord("N")
            s[120] = ord("a")
        if j == 3:
            s = bytearray(rc(bytes(s)))
        if j == 5:
            s = s[:30]
        seqs.append(bytes(s))
    seqs.append(b"ACGT")
    seqs.append(b"")
    cov_off = [0]
    cov_val = []
    for s in seqs:
        st = s.decode("latin-1")
        for i in range(len(st) - 23 + 1):
            cov_val.append(w.get_tf_value(st[i:i + 23]))
        cov_off.append(len(cov_val))
    out["cov_seq"] = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    out["cov_seq_off"] = np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64)
    out["cov_off"] = np.array(cov_off, dtype=np.int64)
    out["cov_val"] = np.array(cov_val, dtype=np.uint32)
    out["n_kmers"] = np.array([w.n_kmers], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "golden23.npz"), **out)
    print(f"golden23: n={len(kmers)} queries={len(q)} hits={(out['tf'] > 0).sum()} "
          f"positions={len(pos_val)} cov={len(cov_val)}")


def make_reads13(rng):
    genome = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=3000).tobytes()
    lines = []
    for i in range(260):
        ln = int(rng.integers(10, 90))
        st = int(rng.integers(0, len(genome) - ln))
        r = bytearray(genome[st:st + ln])
        if i % 9 == 1:
            r[int(rng.integers(0, ln))] = ord("N")
        if i % 13 == 2:
            r = bytearray(bytes(r).lower())
        if i % 17 == 3:
            r[int(rng.integers(0, ln))] = ord("x")
        lines.append(bytes(r))
    lines.insert(40, b"")
    lines.insert(90, b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA")
    return genome, lines


def golden13(rng, ref, tmp):
    genome, lines = make_reads13(rng)
    pf = O.PF13_PATH
    out = {}
    texts = {
        "plain": b"\n".join(lines) + b"\n",
        "plain_nonl": b"\n".join(lines),
        "fastq": b"".join(b"@r%d\n" % i + l + b"\n+\n" + b"I" * len(l) + b"\n"
                          for i, l in enumerate(lines) if l),
        "fasta": b"".join(b">s%d desc\n" % i + b"\n".join(l[j:j + 25] for j in range(0, len(l), 25))
                          + b"\n" for i, l in enumerate(lines) if l),
        "crlf": b"\r\n".join(lines[:50]) + b"\r\n",
    }
    for name, data in texts.items():
        path = os.path.join(tmp, f"r13_{name}.txt")
        with open(path, "wb") as f:
            f.write(data)
        tfp = os.path.join(tmp, f"r13_{name}.tf.bin")
        log = run([f"{BIN}/count_kmers13", path, pf, tfp, "2"])
        tf = np.fromfile(tfp, dtype=np.uint64)
        assert tf.size == 1 << 26
        nz = np.nonzero(tf)[0]
        out[f"{name}_data"] = np.frombuffer(data, dtype=np.uint8)
        out[f"{name}_ids"] = nz.astype(np.uint32)
        out[f"{name}_counts"] = tf[nz]
        out[f"{name}_md5"] = np.array(hashlib.md5(tf.tobytes()).hexdigest())
        st = {}
        for key, pat in (("sequences", r"Sequences processed: (\d+)"),
                         ("windows", r"Total k-mers processed: (\d+)"),
                         ("valid", r"Valid k-mers: (\d+)"), ("invalid", r"Invalid k-mers: (\d+)")):
            st[key] = int(re.search(pat, log).group(1))
        out[f"{name}_stats"] = np.array([st["sequences"], st["windows"], st["valid"], st["invalid"]],
                                        dtype=np.uint64)
        print(f"golden13 {name}: {st} distinct={nz.size}")
        if name != "plain":
            os.unlink(tfp)
    # queries against the plain index through the reference module
    tfp = os.path.join(tmp, "r13_plain.tf.bin")
    w = ref.AindexWrapper()
    w.load_13mer_index(pf, tfp)
    tf = np.fromfile(tfp, dtype=np.uint64)
    q = []
    for l in lines[:120]:
        if len(l) >= 13:
            q.append(l[:13])
            q.append(rc(l[:13].upper()) if not l[:13].upper().strip(b"ACGT") else l[-13:])
    q += [rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=13).tobytes() for _ in range(200)]
    q += [b"ACGTACGTACGT", b"ACGTACGTACGTAC", b"", b"NNNNNNNNNNNNN", b"acgtacgtacgta",
          b"ACGTNCGTACGTA", b"AAAAAAAAAAAAA", b"TTTTTTTTTTTTT"]
    qs = [x.decode("latin-1") for x in q]
    recs, lens = O.pack_queries(q, stride=16)
    out["q_recs"], out["q_lens"] = recs, lens
    out["q_tf"] = np.array(w.get_tf_values(qs), dtype=np.uint32)
    # total / both: the reference has no validity check and indexes the mmap with an id that
    # can be 4^13 for non-keys -> only ask it about valid ACGT 13-mers (others: oracle-defined 0)
    ok = [i for i, s in enumerate(q) if len(s) == 13 and not s.strip(b"ACGT")]
    out["q_ok"] = np.array(ok, dtype=np.int64)
    out["q_total"] = np.array(w.get_total_tf_values_13mer([qs[i] for i in ok]), dtype=np.uint64)
    out["q_both"] = np.array(w.get_tf_both_directions_13mer_batch([qs[i] for i in ok]),
                             dtype=np.uint64).reshape(-1, 2)
    # positions index: compute_aindex13 needs the uint32-narrowed tf
    tf32 = os.path.join(tmp, "r13_plain.tf32.bin")
    tf.astype(np.uint32).tofile(tf32)
    pre = os.path.join(tmp, "r13")
    run([f"{BIN}/compute_aindex13", os.path.join(tmp, "r13_plain.txt"), pf, tf32, pre, "1"])
    ind = np.fromfile(pre + ".indices.bin", dtype=np.uint64)
    pos = np.fromfile(pre + ".index.bin", dtype=np.uint64)
    assert ind.size == (1 << 26) + 1
    assert np.array_equal(ind[1:], np.cumsum(tf)), "indices != cumsum(tf)"
    out["pos13_positions"] = pos
    out["pos13_indices_md5"] = np.array(hashlib.md5(ind.tobytes()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "golden13.npz"), **out)
    print(f"golden13: queries={len(q)} positions={pos.size}")


def golden_kat():
    """Small known-answer table straight from the reference module / binaries."""
    ref = O.ref_module()
    w = ref.AindexWrapper()
    out = {}
    rcs = ["ACGTACGTACGTACGTACGTACG", "AAAAAAAAAAAAAAAAAAAAAAA", "GATTACAGATTACAGATTACAGA",
           "NNNNNNNNNNNNNNNNNNNNNNN", "ACGTNCGTACGTACGTACGTACG"]
    out["rc23_in"] = np.array([s.encode() for s in rcs])
    out["rc23_out"] = np.array([w.get_reverse_complement_23mer(s).encode() for s in rcs])
    rcs13 = ["GATTACAGATTAC", "AAAAAAAAAAAAA", "ACGTNCGTACGTA", "acgtacgtacgta"]
    out["rc13_in"] = np.array([s.encode() for s in rcs13])
    out["rc13_out"] = np.array([w.get_reverse_complement_13mer(s).encode() for s in rcs13])
    np.savez_compressed(os.path.join(HERE, "golden_kat.npz"), **out)


def golden_reads():
    """Read access / position -> read mapping of the reference module on the committed idx23
    fixture (python_wrapper.cpp:261-322, :666-698, :757-789).  Writes idx23.ridx (the
    compute_reads layout: rid, start, end-exclusive) and golden_reads.npz."""
    ref = O.ref_module()
    reads = open(os.path.join(HERE, "idx23.reads"), "rb").read()
    with open(os.path.join(HERE, "idx23.ridx"), "w") as f:
        pos = 0
        for rid, line in enumerate(reads.split(b"\n")[:-1]):
            f.write(f"{rid}\t{pos}\t{pos + len(line)}\n")
            pos += len(line) + 1
    w = ref.AindexWrapper()
    prefix = os.path.join(HERE, "idx23")
    w.load_from_prefix_23mer(prefix, prefix + ".reads")
    w.load_aindex_from_prefix_23mer(prefix, 100000)
    rng = np.random.default_rng(77)
    n_reads = int(w.n_reads)
    starts = np.cumsum([0] + [len(l) + 1 for l in reads.split(b"\n")[:-1]])
    probe = sorted(set(int(x) for x in rng.integers(0, len(reads), size=300)) |
                   set(int(s + d) for s in starts[:40] for d in (-2, -1, 0, 1) if 0 <= s + d < len(reads)))
    out = {"pos": np.array(probe, dtype=np.uint64),
           "rid": np.array([w.get_rid(p) for p in probe], dtype=np.uint64),
           "start": np.array([w.get_start(p) for p in probe], dtype=np.uint64),
           "n_reads": np.array([n_reads], dtype=np.uint64),
           "reads_size": np.array([w.reads_size], dtype=np.uint64)}
    rids = [0, 1, 2, 7, n_reads - 1, n_reads, n_reads + 5]
    out["read_rid"] = np.array(rids, dtype=np.uint64)
    out["read_str"] = np.array([w.get_read_by_rid(r).encode() for r in rids])
    spans = [(0, 30, False), (5, 40, True), (100, 100, False), (200, 150, False), (len(reads) - 10, len(reads) - 2, True)]
    out["span"] = np.array([(a, b, int(c)) for a, b, c in spans], dtype=np.int64)
    out["span_str"] = np.array([w.get_read(a, b, c).encode() for a, b, c in spans])
    np.savez_compressed(os.path.join(HERE, "golden_reads.npz"), **out)
    print(f"golden_reads: n_reads={n_reads} probes={len(probe)}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "reads":
        return golden_reads()
    if not os.path.isdir(BIN):
        raise SystemExit("oracle/_ref missing: run oracle/build_ref.sh in the build container")
    ref = O.ref_module()
    rng = np.random.default_rng(20261018)
    with tempfile.TemporaryDirectory(prefix="aix_golden_") as tmp:
        genome, reads, kmers, counts, prefix = build_idx23(tmp, rng)
        golden23(rng, ref, genome, reads, kmers, counts, prefix)
        golden13(rng, ref, tmp)
    golden_kat()
    golden_reads()


if __name__ == "__main__":
    main()
