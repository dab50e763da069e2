/*
 * aindex_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the reference algorithms of ad3002/aindex for the
 * hot path (2-bit codec, emphf MPHF lookup, 23-mer/13-mer tf queries, 13-mer
 * counting, coverage, positions index).  Every function cites the reference
 * file:line it follows.  Nothing here is on the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The CUDA library (aindex_b200/csrc) never links it.
 *
 * Parity of this oracle is PINNED: tests/test_oracle_golden.py checks it
 * against known-answer vectors generated from the unmodified reference
 *  and against fixtures under tests/golden/ produced by the
 * compiled reference (oracle/_ref, recipe oracle/build_ref.sh,
 * generator tests/golden/make_golden.py).
 */
#ifndef AINDEX_ORACLE_H
#define AINDEX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_TOTAL_13MERS 67108864ULL /* 4^13, count_kmers13.cpp:27 */

/* ---- emphf: base_hash.hpp, mphf.hpp, ranked_bitpair_vector.hpp ---------- */
typedef struct orc_mphf {
    uint64_t n;           /* mphf.hpp:116 m_n */
    uint64_t hash_domain; /* mphf.hpp:117 m_hash_domain */
    uint64_t seed;        /* base_hash.hpp:147 m_seed */
    uint64_t bv_size;     /* bitpair_vector.hpp:117 m_size (= 3*hash_domain) */
    uint64_t n_words;     /* (bv_size+31)/32 */
    uint64_t n_blocks;    /* (bv_size+511)/512 */
    uint64_t *words;
    uint64_t *block_ranks;
} orc_mphf;

void orc_jenkins64(uint64_t seed, const uint8_t *s, uint64_t len, uint64_t out[3]);
orc_mphf *orc_mphf_load(const char *pf_path);
orc_mphf *orc_mphf_from_arrays(uint64_t n, uint64_t hash_domain, uint64_t seed,
                               const uint64_t *words, uint64_t n_words,
                               const uint64_t *block_ranks, uint64_t n_blocks);
int orc_mphf_save(const orc_mphf *m, const char *pf_path);
void orc_mphf_free(orc_mphf *m);
uint64_t orc_mphf_lookup(const orc_mphf *m, const uint8_t *s, uint64_t len);
void orc_mphf_lookup_batch(const orc_mphf *m, const uint8_t *recs, uint64_t stride,
                           const uint8_t *lens, uint64_t q, uint64_t *out, int threads);

/* ---- codec: kmers.cpp --------------------------------------------------- */
uint64_t orc_dna23_bitset(const uint8_t *s, uint64_t len);
uint32_t orc_dna13_bitset(const uint8_t *s, uint64_t len);
void orc_bitset_dna23(uint64_t x, uint8_t *out, int k);
void orc_bitset_dna13(uint32_t x, uint8_t *out, int k);
uint64_t orc_reverse_dna23(uint64_t x);
uint32_t orc_reverse_dna13(uint32_t x);
/* dna_bitseq.hpp:22-61 packing (4 bases/byte, MSB first, non-ACGT -> A) */
void orc_dna_bitset_pack(const uint8_t *s, uint64_t len, uint8_t *out);
uint64_t orc_dna_bitset_ukmer(const uint8_t *packed, uint64_t pos, int k);

/* ---- 23-mer index: hash.hpp PHASH_MAP, python_wrapper.cpp --------------- */
typedef struct orc_index23 {
    const orc_mphf *mphf;
    const uint64_t *checker; /* .kmers.bin */
    const uint32_t *tf;      /* .tf.bin    */
    uint64_t n;
} orc_index23;

uint32_t orc_tf23(const orc_index23 *ix, const uint8_t *s, uint64_t len);
uint64_t orc_total_tf23(const orc_index23 *ix, const uint8_t *s, uint64_t len);
void orc_both_tf23(const orc_index23 *ix, const uint8_t *s, uint64_t len, uint32_t out[2]);
uint64_t orc_pfid23(const orc_index23 *ix, const uint8_t *s, uint64_t len);
uint64_t orc_kid23(const orc_index23 *ix, const uint8_t *s, uint64_t len);
uint64_t orc_strand23(const orc_index23 *ix, const uint8_t *s, uint64_t len);
uint32_t orc_get_freq23(const orc_index23 *ix, uint64_t ukmer);
/* mode: 0 tf(u32) 1 total(u64) 2 both(u32x2) 3 pfid(u64) 4 strand(u64) 5 kid(u64) */
void orc_tf23_batch(const orc_index23 *ix, const uint8_t *recs, uint64_t stride,
                    const uint8_t *lens, uint64_t q, int mode, void *out, int threads);

/* ---- 13-mer queries: python_wrapper.cpp:482-608, 938-980 ---------------- */
uint32_t orc_tf13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *s, uint64_t len);
void orc_both_tf13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *s, uint64_t len,
                   uint64_t out[2]);
/* mode: 0 tf(u32) 1 total(u64) 2 both(u64x2) */
void orc_tf13_batch(const orc_mphf *m, const uint64_t *tf64, const uint8_t *recs,
                    uint64_t stride, const uint8_t *lens, uint64_t q, int mode, void *out,
                    int threads);

/* ---- 13-mer counting: count_kmers13.cpp --------------------------------- */
typedef struct orc_count_stats {
    uint64_t sequences; /* total_sequences        count_kmers13.cpp:135 */
    uint64_t windows;   /* total_kmers_processed  :143 */
    uint64_t valid;     /* valid_kmers            :152 */
    uint64_t invalid;   /* invalid_kmers          :155,:158 */
} orc_count_stats;
/* fmt: 0 plain 1 fasta 2 fastq -1 detect (count_kmers13.cpp:194-206) */
int orc_detect_format(const uint8_t *bytes, uint64_t len);
/* direct-address histogram hist[v], v = 2-bit value of the window (no MPHF) */
void orc_count13_direct(const uint8_t *bytes, uint64_t len, int fmt, uint64_t *hist,
                        orc_count_stats *st);
/* reference layout counts[mphf(window)] (count_kmers13.cpp:141-160) */
void orc_count13(const orc_mphf *m, const uint8_t *bytes, uint64_t len, int fmt,
                 uint64_t *counts, orc_count_stats *st);

/* ---- coverage: aindex/core/aindex.py:314-322 ---------------------------- */
/* out has max(0,len-k+1) entries; k==23 -> tf23, k==13 -> tf13 */
void orc_coverage23(const orc_index23 *ix, const uint8_t *seq, uint64_t len, uint32_t cutoff,
                    uint32_t *out);
void orc_coverage13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *seq, uint64_t len,
                    uint32_t cutoff, uint32_t *out);

/* ---- positions index ---------------------------------------------------- */
/* hash.hpp:365-399 (prefix sum) + hash.cpp:960-1060 (1 worker, k=23).
 * indices has n+1 entries; positions has indices[n] entries (zeroed here). */
void orc_positions_build23(const orc_index23 *ix, const uint8_t *reads, uint64_t len,
                           uint64_t *indices, uint64_t *positions);
/* compute_aindex13.cpp:36-86, :125-239 with the uint64 tf (defect 2.3#2 not reproduced) */
void orc_positions_build13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *reads,
                           uint64_t len, uint64_t *indices, uint64_t *positions);
/* python_wrapper.cpp:800-822 (absent k-mer -> 0 results, defect 2.3#6 not reproduced).
 * returns number of positions written (<= cap) */
uint64_t orc_positions_query23(const orc_index23 *ix, const uint64_t *indices,
                               const uint64_t *positions, const uint8_t *s, uint64_t len,
                               uint64_t *out, uint64_t cap);
uint64_t orc_positions_query13(const orc_mphf *m, const uint64_t *indices,
                               const uint64_t *positions, uint64_t n_positions, const uint8_t *s,
                               uint64_t len, uint64_t *out, uint64_t cap);

/* ---- canonical 23-mer table (the counting stage in front of the index build) ------------ */
/* Definition: tests/analyze_kmers.py:25-33, :70-71 and the jellyfish -C branch of
 * scripts/compute_aindex.py:164-182 -- every window of 23 upper-case ACGT characters counts for
 * min(kmer, revcomp(kmer)) (lexicographic = numeric on the 2-bit values, kmers.cpp:376-381).
 * (The reference's own kmer_counter is broken)  Returns the number of distinct
 * k-mers; *kmers_out / *counts_out are malloc'ed arrays sorted by k-mer (free with orc_free). */
uint64_t orc_canonical23_count(const uint8_t *reads, uint64_t len, int threads, uint64_t **kmers_out,
                               uint32_t **counts_out);
void orc_free(void *p);
/* the two text files of scripts/compute_aindex.py:140-200: `.dat` = "KMER\tCOUNT\n" lines (input of
 * compute_index, hash.cpp:696-701) and the key file = "KMER\n" lines (`cut -f1`, input of compute_mphf_seq).
 * Either path may be NULL.  Returns 0 on success. */
int orc_write_dat(const uint64_t *kmers, const uint32_t *counts, uint64_t n, const char *dat_path,
                  const char *keys_path);

#ifdef __cplusplus
}
#endif
#endif
