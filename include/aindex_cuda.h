/*
 * aindex_cuda.h -- C-ABI of libaindex_cuda.so, the B200 (sm_100a) implementation of
 * the data-parallel hot path of ad3002/aindex.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only.  The host side
 * (aindex_b200/csrc/python_wrapper.cpp = the pybind11 module `aindex_cpp`, and the
 * count_kmers13 / compute_aindex / compute_aindex13 executables) is written against
 * this header exactly as the reference's host code is written against its own
 * hash.hpp / emphf headers.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference repository).
 *
 * Conventions
 *   - every function returns 0 (AIX_OK) or a negative AIX_ERR_* code; the message is
 *     available from aix_last_error(ctx) (ctx may be NULL for creation failures).
 *   - one aix_ctx per (host thread, GPU).  A ctx owns one CUDA stream family; it is
 *     not thread-safe.  Multi-GPU = one process (or thread) per GPU, each with its own
 *     ctx; the only cross-GPU step on this path (sum of 13-mer histograms) is done by
 *     the caller with NCCL on the device buffer exposed by aix_count13_hist_dev().
 *   - pointers are HOST pointers unless the parameter name ends in `_dev`.  Host
 *     buffers may be pageable; pinned buffers (aix_host_alloc) are copied
 *     asynchronously and overlap with compute.
 *   - there is no CPU fallback: without a usable CUDA device every call fails with
 *     AIX_ERR_CUDA.
 *   - query records: q records of `stride` bytes each; record i holds lens[i] bytes of
 *     the query string (lens == NULL means every record is exactly `stride` bytes long).
 *     Bytes are the raw characters the reference would receive as std::string.
 */
#ifndef AINDEX_CUDA_H
#define AINDEX_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIX_OK 0
#define AIX_ERR_CUDA (-1)    /* CUDA runtime failure / no device */
#define AIX_ERR_ARG (-2)     /* invalid argument */
#define AIX_ERR_IO (-3)      /* file missing / short / unwritable */
#define AIX_ERR_NOMEM (-4)   /* host or device allocation failed */
#define AIX_ERR_STATE (-5)   /* object not in the required state */
#define AIX_ERR_BUILD (-6)   /* MPHF construction did not converge */

#define AIX_TOTAL_13MERS 67108864ULL /* 4^13, src/count_kmers13.cpp:27 */

typedef struct aix_ctx aix_ctx;
typedef struct aix_mphf aix_mphf;       /* emphf::mphf<jenkins64_hasher>, src/hash.hpp:25 */
typedef struct aix_index23 aix_index23; /* PHASH_MAP, src/hash.hpp:82-103 */
typedef struct aix_index13 aix_index13; /* 13-mer mode of AindexWrapper, src/python_wrapper.cpp:137-147 */
typedef struct aix_positions aix_positions; /* AIndexCompressed, src/hash.hpp:357-363 */

/* ---- context ------------------------------------------------------------------ */
int aix_ctx_create(int device, aix_ctx **out);
void aix_ctx_destroy(aix_ctx *ctx);
const char *aix_last_error(const aix_ctx *ctx);
int aix_ctx_device(const aix_ctx *ctx);
/* the CUDA stream (cudaStream_t) the *_dev entry points launch on */
void *aix_ctx_stream(const aix_ctx *ctx);
int aix_ctx_sync(aix_ctx *ctx);
/* The builders (positions index, canonical table, sort) take their large temporaries from a per-ctx CUDA memory pool that
 * keeps freed blocks for the next call; aix_ctx_trim returns the cached memory to the driver (e.g. before another
 * library needs the HBM). */
int aix_ctx_trim(aix_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t aix_ctx_launch_count(const aix_ctx *ctx);
/* pinned host memory for the e2e paths */
int aix_host_alloc(aix_ctx *ctx, size_t bytes, void **out);
int aix_host_free(aix_ctx *ctx, void *p);
const char *aix_version(void);

/* ---- MPHF: src/emphf/mphf.hpp, base_hash.hpp, ranked_bitpair_vector.hpp -------- */
/* mphf::load (mphf.hpp:107-113): u64 n, u64 hash_domain, u64 seed, u64 bv_size,
 * u64 words[(bv_size+31)/32], u64 block_ranks[(bv_size+511)/512] */
int aix_mphf_upload(aix_ctx *ctx, uint64_t n, uint64_t hash_domain, uint64_t seed,
                    const uint64_t *words, uint64_t n_words, const uint64_t *block_ranks,
                    uint64_t n_blocks, aix_mphf **out);
int aix_mphf_load_pf(aix_ctx *ctx, const char *pf_path, aix_mphf **out);
/* mphf::save (mphf.hpp:99-105): byte-identical layout */
int aix_mphf_save_pf(aix_ctx *ctx, const aix_mphf *m, const char *pf_path);
void aix_mphf_destroy(aix_ctx *ctx, aix_mphf *m);
/* header fields: info[0..5] = n, hash_domain, seed, bv_size, n_words, n_blocks */
int aix_mphf_info(const aix_mphf *m, uint64_t info[6]);
/* copies of the host-side arrays in .pf order (sizes from aix_mphf_info) */
int aix_mphf_arrays(const aix_mphf *m, uint64_t *words_out, uint64_t *block_ranks_out);
/* mphf::lookup (mphf.hpp:79-89) on raw byte strings = AindexWrapper::get_hash_values
 * (python_wrapper.cpp:629-642) */
int aix_mphf_lookup(aix_ctx *ctx, const aix_mphf *m, const uint8_t *recs, uint32_t stride,
                    const uint8_t *lens, uint64_t q, uint64_t *ids_out);
/* jenkins64_hasher::operator() (base_hash.hpp:38-91): out[3*i..3*i+2] */
int aix_jenkins64(aix_ctx *ctx, uint64_t seed, const uint8_t *recs, uint32_t stride,
                  const uint8_t *lens, uint64_t q, uint64_t *triples_out);
/* perm13[v] = lookup(ASCII 13-mer of 2-bit value v), v in [0,4^13)  (the table
 * count_kmers13.cpp:148 evaluates once per k-mer occurrence) */
int aix_perm13(aix_ctx *ctx, const aix_mphf *m, uint32_t *perm_out);
/* MPHF construction on the GPU (replaces emphf compute_mphf_seq, mphf.hpp:22-67,
 * hypergraph_sorter_seq.hpp): keys are the ASCII strings of n distinct packed k-mers
 * (k = 23: uint64 values as in .kmers.bin; k = 13: low 26 bits).  gamma = 1.23 and the
 * seed sequence std::mt19937_64(37) follow the reference; the peeling order does not,
 * so the .pf is a valid emphf file for the same keys but not byte-identical. */
int aix_mphf_build(aix_ctx *ctx, const uint64_t *kmers, uint64_t n, int k, aix_mphf **out);
int aix_mphf_build_dev(aix_ctx *ctx, const uint64_t *kmers_dev, uint64_t n, int k, aix_mphf **out);

/* ---- codec: src/kmers.cpp, src/dna_bitseq.hpp ---------------------------------- */
/* get_dna23_bitset / get_dna13_bitset (kmers.cpp:12-85): out[i] = 2-bit value of the
 * first k characters of record i (anything but upper-case ACGT encodes as 0) */
int aix_encode_kmers(aix_ctx *ctx, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                     uint64_t q, int k, uint64_t *out);
/* get_bitset_dna23 / get_bitset_dna13 (kmers.cpp:89-257): q values -> q*k characters */
int aix_decode_kmers(aix_ctx *ctx, const uint64_t *values, uint64_t q, int k, uint8_t *out);
/* reverseDNA (kmers.cpp:376-388) */
int aix_revcomp_kmers(aix_ctx *ctx, const uint64_t *values, uint64_t q, int k, uint64_t *out);
/* dna_bitset ctor (dna_bitseq.hpp:22-61): 4 bases per byte, MSB first, non-ACGT -> A */
int aix_pack_2bit(aix_ctx *ctx, const uint8_t *seq, uint64_t len, uint8_t *packed_out);
int aix_pack_2bit_dev(aix_ctx *ctx, const uint8_t *seq_dev, uint64_t len, uint8_t *packed_dev);
/* dna_bitset::ukmer(pos, k) (dna_bitseq.hpp:124-151), batched: out[i] = the k bases (k <= 32) that start at base pos[i] of
 * the packed sequence, as a 2k-bit number (first base in the top bits).  Bases at or past n_bases read as A. */
int aix_ukmers(aix_ctx *ctx, const uint8_t *packed, uint64_t n_bases, const uint64_t *pos, uint64_t q, int k, uint64_t *out);
int aix_ukmers_dev(aix_ctx *ctx, const uint8_t *packed_dev, uint64_t n_bases, const uint64_t *pos_dev, uint64_t q, int k,
                   uint64_t *out_dev);
/* rolling canonical k-mers of a reads buffer: for every window start i in [0,len-k]
 * fwd_out[i], rc_out[i] (either may be NULL) and valid_out[i] = 1 iff all k characters
 * are upper-case ACGT (the loop of hash.cpp:1006-1032 without the lookup) */
int aix_rolling_kmers(aix_ctx *ctx, const uint8_t *bytes, uint64_t len, int k, uint64_t *fwd_out,
                      uint64_t *rc_out, uint8_t *valid_out);
/* device form: bytes_dev 16-byte aligned and readable up to the next multiple of 16 past len */
int aix_rolling_kmers_dev(aix_ctx *ctx, const uint8_t *bytes_dev, uint64_t len, int k, uint64_t *fwd_dev,
                          uint64_t *rc_dev, uint8_t *valid_dev);

/* ---- 23-mer index: PHASH_MAP + AindexWrapper 23-mer queries --------------------- */
/* load_hash (hash.cpp:367-450): checker = .kmers.bin (u64[n]), tf = .tf.bin (u32[n]) */
int aix_index23_upload(aix_ctx *ctx, const aix_mphf *m, const uint64_t *checker,
                       const uint32_t *tf, uint64_t n, aix_index23 **out);
int aix_index23_upload_dev(aix_ctx *ctx, const aix_mphf *m, const uint64_t *checker_dev,
                           const uint32_t *tf_dev, uint64_t n, aix_index23 **out);
/* AindexWrapper::load_from_prefix_23mer (python_wrapper.cpp:1103-1132):
 * {prefix}.pf + {prefix}.kmers.bin + {prefix}.tf.bin; *mphf_out is owned by the caller */
int aix_index23_load_prefix(aix_ctx *ctx, const char *prefix, aix_mphf **mphf_out,
                            aix_index23 **out);
void aix_index23_destroy(aix_ctx *ctx, aix_index23 *ix);
/* info[0] = n, info[1] = 1 if every stored k-mer is canonical (enables the one-probe
 * path, which is result-identical) */
int aix_index23_info(const aix_index23 *ix, uint64_t info[2]);
/* HBM/L2 layout chosen at upload: info[0] = fingerprint bits per key (0 = none, 4, 8), info[1] = bytes of
 * the separate fingerprint tier (0 when the fingerprints ride inside the MPHF records), info[2] = MPHF record bytes,
 * info[3] = 0 wide MPHF records, 1 compact (48 pairs + u32 rank), 2 fused (16 pairs + 16 x 4-bit fingerprint + u32 rank) */
int aix_index23_layout(const aix_index23 *ix, uint64_t info[4]);
/* Front filter of the batch tf path (a blocked Bloom filter over the stored k-mers of a canonical-only index, built at
 * upload; AIX_BLOOM_BITS per key, default 8, 0 = none).  Batches in which most queries are absent are answered through it
 * (one 8-byte request per absent query instead of the hash + three MPHF records); the launcher decides per batch from the
 * pass rate the filter kernel reports (AIX_INDEX23_FILTER=on|off forces).  Answers do not depend on the choice.
 * info = {filter bytes, queries counted (sampled CTAs), of those passed, batches through the filter, batches direct,
 *         last pass rate * 1e6 (~0 = none observed yet)}. */
int aix_index23_filter_stats(const aix_index23 *ix, uint64_t info[6]);
/* mode 0: the launcher decides per batch (default), 1: every batch through the filter, 2: never */
int aix_index23_set_filter(aix_index23 *ix, int mode);
/* index fill on the GPU (replaces compute_index / index_hash_pp, hash.cpp:671-723,
 * :779-881): checker_out[h] = kmers[i], tf_out[h] = counts[i], h = mphf(kmers[i]) */
int aix_index23_fill(aix_ctx *ctx, const aix_mphf *m, const uint64_t *kmers,
                     const uint32_t *counts, uint64_t n, uint64_t *checker_out, uint32_t *tf_out);
int aix_index23_fill_dev(aix_ctx *ctx, const aix_mphf *m, const uint64_t *kmers_dev,
                         const uint32_t *counts_dev, uint64_t n, uint64_t *checker_dev,
                         uint32_t *tf_dev);

#define AIX_Q_TF 0     /* get_tf_value(s)_23mer      python_wrapper.cpp:610-627  -> u32[q]   */
#define AIX_Q_TOTAL 1  /* get_total_tf_value(s)_23mer :1230-1258                  -> u64[q]   */
#define AIX_Q_BOTH 2   /* get_tf_both_directions_23mer(_batch) :1260-1286         -> u32[2q]  */
#define AIX_Q_PFID 3   /* PHASH_MAP::get_pfid          hash.hpp:150-170 (n = absent) -> u64[q] */
#define AIX_Q_STRAND 4 /* get_strand                   python_wrapper.cpp:726-742 -> u64[q]   */
#define AIX_Q_KID 5    /* get_kid_by_kmer              :700-716                   -> u64[q]   */
int aix_tf23_batch(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *recs, uint32_t stride,
                   const uint8_t *lens, uint64_t q, int mode, void *out);
/* same, all buffers already in HBM; asynchronous on aix_ctx_stream() */
int aix_tf23_batch_dev(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *recs_dev,
                       uint32_t stride, const uint8_t *lens_dev, uint64_t q, int mode,
                       void *out_dev);
/* Latency of single calls (aix_tf23_batch with q == 1, mode AIX_Q_TF = AindexWrapper::get_tf_value, python_wrapper.cpp:610-627)
 * measured from C, without an interpreter in the loop: the n queries (n x stride bytes, host) are sent one by one through the
 * resident mailbox kernel, first as echo requests (answered at once: host store -> device poll over PCIe -> device store ->
 * host load = the transport's share), then as real lookups (tf_out[n], may be NULL).  Nanoseconds per call, averaged. */
int aix_tf23_single_call_latency(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *recs, uint32_t stride, uint64_t n,
                                 uint32_t *tf_out, double *echo_ns, double *query_ns);
/* Index split by hash-id range over several GPUs (one aix_index23 per rank holding records [lo, hi) of the
 * checker / tf arrays, uploaded with aix_index23_upload[_dev] on the slice; the MPHF is replicated):
 *   aix_tf23_probes_dev: query i -> probes_dev[4*i .. 4*i+3] = {id1, kmer1, id2, kmer2}; id = ~0 means "no probe".
 *     These are the (at most two) checker comparisons of get_tf_value_23mer (python_wrapper.cpp:610-622) with GLOBAL ids;
 *     canonical_only = 1 only if every rank's slice is canonical (then one probe per valid query).
 *   [caller: route each probe to the owner of its id, subtract the owner's lo]
 *   aix_probe23_dev: probes {local id, kmer} -> out = (hit << 32) | tf.
 *   answer of query i = first hit of its two probes, else 0  (aindex_b200/dist.py::ShardedIndex23). */
int aix_tf23_probes_dev(aix_ctx *ctx, const aix_mphf *m, uint64_t n_total, int canonical_only,
                        const uint8_t *recs_dev, uint32_t stride, const uint8_t *lens_dev, uint64_t q,
                        uint64_t *probes_dev);
int aix_probe23_dev(aix_ctx *ctx, const aix_index23 *shard, const uint64_t *probes_dev, uint64_t cnt,
                    uint64_t *out_dev);
/* the routing step on the device: probes (as written by aix_tf23_probes_dev, n_probes = 2q) -> buckets by owner.
 * bounds: HOST array of world + 1 ascending ids (rank r owns [bounds[r], bounds[r+1])).  counts_dev[world] (u64) =
 * probes per owner; send_dev[n_probes] x {id - bounds[owner], kmer} grouped by owner in rank order (first
 * sum(counts) entries used); tag_dev[slot] = index of the probe that sits in slot (to scatter the answers back). */
int aix_probes_bucket_dev(aix_ctx *ctx, const uint64_t *probes_dev, uint64_t n_probes, const uint64_t *bounds,
                          int world, uint64_t *counts_dev, uint64_t *send_dev, uint32_t *tag_dev);
/* PHASH_MAP::get_freq(uint64_t) (hash.hpp:123-140) for packed k-mers */
int aix_get_freq23(aix_ctx *ctx, const aix_index23 *ix, const uint64_t *ukmers, uint64_t q,
                   uint32_t *out);

/* the same for k-mers held as 6-byte dna_bitset records (dna_bitseq.hpp:22-61: 4 bases per byte, first base in bits 7:6,
 * 23 bases + 2 zero bits; aix_pack_2bit of the 23 characters): 6 B per query over PCIe instead of 23 */
int aix_get_freq23_packed(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *packed, uint64_t q, uint32_t *out);
int aix_get_freq23_packed_dev(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *packed_dev, uint64_t q,
                              uint32_t *out_dev);

/* ---- 13-mer index: AindexWrapper 13-mer mode ----------------------------------- */
/* load_13mer_index (python_wrapper.cpp:404-437): tf64 = the 4^13 x u64 .tf.bin */
int aix_index13_upload(aix_ctx *ctx, const aix_mphf *m, const uint64_t *tf64, aix_index13 **out);
void aix_index13_destroy(aix_ctx *ctx, aix_index13 *ix);
/* out[v] = tf of the 13-mer with 2-bit value v (the tf file re-indexed by the MPHF), 4^13 x u64: the array the
 * reference's frequency iterator (aindex/core/aindex.py:632-652, `_index_to_13mer(index)`) assumes it is given */
int aix_index13_tf_direct(aix_ctx *ctx, const aix_index13 *ix, uint64_t *out);
/* AIX_Q_TF    get_tf_value(s)_13mer :482-503, :938-980                 -> u32[q]
 * AIX_Q_TOTAL get_total_tf_value(s)_13mer :522-566                     -> u64[q]
 * AIX_Q_BOTH  get_tf_both_directions_13mer(_batch) :567-608            -> u64[2q] */
int aix_tf13_batch(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *recs, uint32_t stride,
                   const uint8_t *lens, uint64_t q, int mode, void *out);
int aix_tf13_batch_dev(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *recs_dev,
                       uint32_t stride, const uint8_t *lens_dev, uint64_t q, int mode,
                       void *out_dev);

/* ---- 13-mer counting: src/count_kmers13.cpp ------------------------------------- */
#define AIX_FMT_DETECT (-1) /* detect_format, count_kmers13.cpp:194-206 */
#define AIX_FMT_PLAIN 0     /* read_plain_file :262-272 */
#define AIX_FMT_FASTA 1     /* read_fasta_file :211-235 */
#define AIX_FMT_FASTQ 2     /* read_fastq_file :240-257 */
typedef struct aix_count_stats {
    uint64_t sequences; /* total_sequences       count_kmers13.cpp:135 */
    uint64_t windows;   /* total_kmers_processed :143 */
    uint64_t valid;     /* valid_kmers           :152 */
    uint64_t invalid;   /* invalid_kmers         :155, :158 */
} aix_count_stats;
/* One-shot: Kmer13Counter::count_kmers_from_file + save_counts (:277-388) on a file
 * image in host memory.  tf_out = 4^13 x u64 in MPHF order (= the .tf.bin file). */
int aix_count13(aix_ctx *ctx, const aix_mphf *m, const uint8_t *bytes, uint64_t len, int fmt,
                uint64_t *tf_out, aix_count_stats *stats);
/* Streaming form used for sharded / multi-GPU counting:
 *   begin -> add (any number of shards, each starting at a line start)
 *         -> [caller: NCCL reduce-scatter / all-reduce on aix_count13_hist_dev()]
 *         -> finish (apply the MPHF permutation, widen to u64).
 * The device histogram is u32[4^13] in direct-address order (v = 2-bit value); add
 * flushes it into a u64 shadow before any counter could wrap. */
int aix_count13_begin(aix_ctx *ctx);
int aix_count13_add(aix_ctx *ctx, const uint8_t *bytes, uint64_t len, int fmt);
int aix_count13_add_dev(aix_ctx *ctx, const uint8_t *bytes_dev, uint64_t len, int fmt);
/* u64[4^13] direct-address totals in HBM (valid after aix_count13_flush) */
int aix_count13_flush(aix_ctx *ctx);
uint64_t *aix_count13_hist_dev(aix_ctx *ctx);
int aix_count13_stats(aix_ctx *ctx, aix_count_stats *stats);
/* [v_begin, v_end) of the direct-address histogram -> tf_out[perm13[v]] += hist[v];
 * tf_out (host, 4^13 x u64) must be zeroed by the caller before the first slice */
int aix_count13_finish(aix_ctx *ctx, const aix_mphf *m, uint64_t v_begin, uint64_t v_end,
                       uint64_t *tf_out, aix_count_stats *stats);
int aix_count13_finish_dev(aix_ctx *ctx, const aix_mphf *m, uint64_t v_begin, uint64_t v_end,
                           uint64_t *tf_out_dev);
int aix_count13_end(aix_ctx *ctx);
/* Histogram combine over NVLink peer memory instead of widen + NCCL reduce-scatter (one rank per GPU, one node):
 *   export (3 x 64-byte CUDA IPC handles) -> the caller all-gathers the handles -> peers_open(all, n, my_rank)
 *   begin -> add ... -> [caller: barrier on the stream] -> reduce_peers_dev(my range) -> [caller: barrier] -> ...
 * reduce_peers_dev writes out_dev[v - v_begin] = sum over ranks of their counts of k-mer v (u64, direct-address
 * order): what aix_count13_flush + reduce-scatter + slicing would give.  No flush is needed before it. */
int aix_count13_ipc_export(aix_ctx *ctx, void *handles_out /* 192 bytes */);
int aix_count13_peers_open(aix_ctx *ctx, const void *handles /* n_ranks x 192 bytes */, int n_ranks,
                           int my_rank);
int aix_count13_reduce_peers_dev(aix_ctx *ctx, uint64_t v_begin, uint64_t v_end, uint64_t *out_dev);
int aix_count13_peers_close(aix_ctx *ctx);

/* ---- several GPUs of one box from one process (host code stays C++: no torch, no NCCL) ------------------------ */
/* Kmer13Counter::count_kmers_from_file (count_kmers13.cpp:277-353) spawns num_threads workers over one table
 * (:305-309); aix_multi maps the workers to GPUs: one ctx + one host thread per GPU, peer access between all pairs.
 * n_dev <= 0 = every visible GPU; dev_ids == NULL = devices 0 .. n_dev-1. */
typedef struct aix_multi aix_multi;
int aix_multi_create(int n_dev, const int *dev_ids, aix_multi **out);
void aix_multi_destroy(aix_multi *mg);
int aix_multi_size(const aix_multi *mg);
aix_ctx *aix_multi_ctx(aix_multi *mg, int i);
/* 1 when every pair of the GPUs has peer access (the exchange step then runs as one kernel per GPU over NVLink peer
 * pointers), 0 when ranges are moved with cudaMemcpyPeer instead */
int aix_multi_peer_access(const aix_multi *mg);
const char *aix_multi_last_error(const aix_multi *mg);
/* aix_count13 over all GPUs of mg: the file image is cut at record boundaries (one shard per GPU), every GPU counts
 * its shard into its own direct-address histogram, GPU r sums k-mer range r over all GPUs (the reduce-scatter of
 * SURVEY 8(e)) and permutes it into .tf.bin order; tf_out / stats as aix_count13.  m = an MPHF uploaded on any ctx
 * (its host arrays are re-uploaded to every GPU).  The number of GPUs must divide 4^13 (1, 2, 4, 8, 16). */
int aix_count13_multi(aix_multi *mg, const aix_mphf *m, const uint8_t *bytes, uint64_t len, int fmt, uint64_t *tf_out,
                      aix_count_stats *stats);
/* the same with one shard per GPU already resident in that GPU's HBM (shards_dev[r] on the device of aix_multi_ctx(mg, r),
 * each starting at a record start; fmt explicit).  tf_out == NULL stops after the exchange step: GPU r then holds the
 * summed k-mer range r in aix_count13_hist_dev(aix_multi_ctx(mg, r)) -- the state an NCCL reduce-scatter leaves. */
int aix_count13_multi_dev(aix_multi *mg, const aix_mphf *m, const uint8_t *const *shards_dev, const uint64_t *lens, int fmt,
                          uint64_t *tf_out, aix_count_stats *stats);

/* AIndexCompressed::fill_index_from_reads over the GPUs of mg (hash.hpp:407-444 splits the byte range of the reads file over
 * worker threads): ix[r] = the same 23-mer index uploaded on the GPU of aix_multi_ctx(mg, r).  GPU r looks up the windows that
 * start in its byte range; buckets are cut into one range per GPU with equal numbers of slots; keys go to the owner of their
 * bucket by direct peer copies (NVLink), the owner sorts them and owns that slice of positions[].  Outputs are the arrays
 * aix_positions_build23 gives (bit-identical), stats (nullable) the phase times (max over GPUs) and the bytes that crossed GPUs. */
typedef struct aix_multi_build_stats {
    double total_ms, upload_scan_ms, emit_partition_ms, exchange_ms, sort_finalize_ms, download_ms;
    uint64_t keys, peer_bytes, positions;
    double alloc_ms; /* cudaMalloc / cudaFree of the exchange buffers (peer-visible memory outside the pool), not part of exchange_ms */
} aix_multi_build_stats;
int aix_positions_build23_multi(aix_multi *mg, const aix_index23 *const *ix, const uint8_t *reads, uint64_t len,
                                uint64_t *indices_out, uint64_t *positions_out, aix_multi_build_stats *stats);

/* ---- coverage: aindex/core/aindex.py:314-322 ------------------------------------- */
/* n_seq sequences concatenated in `seqs`, sequence s = seqs[offs[s] .. offs[s+1]).
 * out holds sum_s max(0, len_s-k+1) values: out[.] = tf >= cutoff ? tf : 0.
 * k = 23 uses ix23, k = 13 uses ix13 (the other may be NULL). */
int aix_coverage(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13,
                 const uint8_t *seqs, const int64_t *offs, uint64_t n_seq, int k, uint32_t cutoff,
                 uint32_t *out);
/* device form: seqs_dev must be readable for 8 bytes past total_bytes (aligned word loads of the last window) */
int aix_coverage_dev(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13,
                     const uint8_t *seqs_dev, const int64_t *offs_dev, uint64_t n_seq,
                     uint64_t total_bytes, uint64_t total_out, int k, uint32_t cutoff,
                     uint32_t *out_dev);

/* ---- positions index: src/hash.hpp:357-490, src/hash.cpp:960-1060,
 *      src/compute_aindex13.cpp:36-323 --------------------------------------------- */
/* AIndexCompressed ctor + fill_index_from_reads (1 worker = ascending order per bucket,
 * first tf occurrences kept, zero tail) on a .reads file image.
 * indices_out: u64[n+1]; positions_out: u64[indices_out[n]] (query the size first with
 * aix_positions_total23 / _total13). */
int aix_positions_total23(aix_ctx *ctx, const aix_index23 *ix, uint64_t *total);
int aix_positions_build23(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *reads, uint64_t len,
                          uint64_t *indices_out, uint64_t *positions_out);
int aix_positions_total13(aix_ctx *ctx, const aix_index13 *ix, uint64_t *total);
int aix_positions_build13(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *reads, uint64_t len,
                          uint64_t *indices_out, uint64_t *positions_out);
/* same on a .reads image already in HBM (readable for 8 bytes past len); the index stays in
 * HBM as an aix_positions object (C5 at full size: 51 GB of positions never leave the GPU) */
int aix_positions_build23_dev(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *reads_dev,
                              uint64_t len, aix_positions **out);
int aix_positions_build13_dev(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *reads_dev,
                              uint64_t len, aix_positions **out);
/* info[0] = number of indices (n+1), info[1] = number of positions (indices[n]) */
int aix_positions_info(const aix_positions *p, uint64_t info[2]);
int aix_positions_arrays_dev(const aix_positions *p, const uint64_t **indices_dev,
                             const uint64_t **positions_dev);
/* AIndexCompressed::save (hash.hpp:470-486): the arrays as they are written to
 * .indices.bin / .index.bin (either pointer may be NULL) */
int aix_positions_download(aix_ctx *ctx, const aix_positions *p, uint64_t *indices_out,
                           uint64_t *positions_out);
/* AindexWrapper::load_aindex (python_wrapper.cpp:361-402): upload .indices.bin/.index.bin */
int aix_positions_upload(aix_ctx *ctx, const uint64_t *indices, uint64_t n_indices,
                         const uint64_t *positions, uint64_t n_positions, aix_positions **out);
void aix_positions_destroy(aix_ctx *ctx, aix_positions *p);
/* get_positions_23mer / get_positions_13mer (python_wrapper.cpp:800-822, 1070-1101), batched:
 * pass 1 (pos_out == NULL): counts_out[i] = number of positions of query i;
 * pass 2: pos_out filled at offs[i] (exclusive scan of counts), values are 0-based. */
int aix_positions_query(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13,
                        const aix_positions *p, const uint8_t *recs, uint32_t stride,
                        const uint8_t *lens, uint64_t q, int k, uint64_t *counts_out,
                        const uint64_t *offs, uint64_t *pos_out);
/* same, all buffers in HBM; asynchronous on aix_ctx_stream() */
int aix_positions_query_dev(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13,
                            const aix_positions *p, const uint8_t *recs_dev, uint32_t stride,
                            const uint8_t *lens_dev, uint64_t q, int k, uint64_t *counts_dev,
                            const uint64_t *offs_dev, uint64_t *pos_out_dev);

/* ---- canonical 23-mer table (input of the index build) ------------ */
/* distinct canonical 23-mers (min(kmer, revcomp), ACGT-only windows, '\n' and '~'
 * break windows) of a .reads image with their counts, sorted ascending.
 * pass 1: kmers_out == NULL -> *n_out = number of distinct k-mers (result kept in ctx);
 * pass 2: copies them out. */
int aix_canonical23_count(aix_ctx *ctx, const uint8_t *reads, uint64_t len, uint64_t *n_out,
                          uint64_t *kmers_out, uint32_t *counts_out);
/* same on a reads image already in HBM; the table stays in HBM (owned by ctx until the
 * next call) and is exposed by aix_canonical23_result_dev */
int aix_canonical23_count_dev(aix_ctx *ctx, const uint8_t *reads_dev, uint64_t len,
                              uint64_t *n_out);
int aix_canonical23_result_dev(aix_ctx *ctx, const uint64_t **kmers_dev,
                               const uint32_t **counts_dev, uint64_t *n);

/* the counting stage's output files (scripts/compute_aindex.py:140-200; jellyfish dump / kmer_counter format):
 * dat_path = "KMER\tCOUNT\n" lines (input of the reference's compute_index, hash.cpp:696-701), keys_path = "KMER\n" lines
 * (input of compute_mphf_seq); either may be NULL.  kmers / counts are HOST arrays (as returned by aix_canonical23_count). */
int aix_write_dat(aix_ctx *ctx, const uint64_t *kmers, const uint32_t *counts, uint64_t n, const char *dat_path,
                  const char *keys_path);

/* ---- building blocks of the two builders above, exposed for tests and callers that hold keys in HBM ------- */
/* stable LSD radix sort of n u64 keys on bits [begin_bit, end_bit) (hand-written, csrc/radix_sort.cu; replaces the
 * `sort` of scripts/compute_aindex.py:140-182 and the per-bucket ordering of the 1-thread worker, hash.cpp:1006-1051).
 * alt_dev = spare buffer of n keys; *result_in_alt = 1 when the sorted keys ended up in alt_dev. */
int aix_sort_u64_dev(aix_ctx *ctx, uint64_t *keys_dev, uint64_t *alt_dev, uint64_t n, int begin_bit, int end_bit,
                     int *result_in_alt);
/* stable partition of n u64 keys into n_ranges <= 16 key ranges (bounds[r] = first key of range r, HOST array; bounds[0]
 * counts as 0): out_dev = the keys grouped by range in range order, input order kept inside a range; counts_out[r] (HOST) =
 * keys of range r.  The routing step of the multi-GPU positions build (keys to the owner of their bucket range). */
int aix_partition_u64_dev(aix_ctx *ctx, const uint64_t *keys_dev, uint64_t *out_dev, uint64_t n, const uint64_t *bounds,
                          int n_ranges, uint64_t *counts_out);
/* run-length encoding of a sorted array (`uniq -c`): uniq_dev[r], counts_dev[r] for the *n_runs distinct keys */
int aix_rle_u64_dev(aix_ctx *ctx, const uint64_t *sorted_dev, uint64_t n, uint64_t *uniq_dev, uint32_t *counts_dev,
                    uint64_t *n_runs);

#ifdef __cplusplus
}
#endif
#endif /* AINDEX_CUDA_H */
