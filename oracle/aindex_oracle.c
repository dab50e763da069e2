/*
 * aindex_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see aindex_oracle.h).
 *
 * Plain-C restatement of the reference's hot-path algorithms.  Scalar and
 * deliberately literal: it exists to be obviously equal to the reference, not
 * to be fast.  The batch entry points take a `threads` argument (OpenMP) only
 * so the cpu_baseline leg of bench.py can use all host cores.
 *
 * Reference paths are relative to ad3002/aindex (mounted at /root/reference).
 */
#include "aindex_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ========================================================================= */
/* emphf: Jenkins lookup8 triple hash  (src/emphf/base_hash.hpp:38-91, 127-145) */
/* ========================================================================= */

static inline uint64_t load64le(const uint8_t *p) { /* base_hash.hpp:11-17 */
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}

static inline void jenkins_mix(uint64_t *pa, uint64_t *pb, uint64_t *pc) { /* :127-145 */
    uint64_t a = *pa, b = *pb, c = *pc;
    a -= b; a -= c; a ^= (c >> 43);
    b -= c; b -= a; b ^= (a << 9);
    c -= a; c -= b; c ^= (b >> 8);
    a -= b; a -= c; a ^= (c >> 38);
    b -= c; b -= a; b ^= (a << 23);
    c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 35);
    b -= c; b -= a; b ^= (a << 49);
    c -= a; c -= b; c ^= (b >> 11);
    a -= b; a -= c; a ^= (c >> 12);
    b -= c; b -= a; b ^= (a << 18);
    c -= a; c -= b; c ^= (b >> 22);
    *pa = a; *pb = b; *pc = c;
}

void orc_jenkins64(uint64_t seed, const uint8_t *s, uint64_t len, uint64_t out[3]) {
    uint64_t a = seed, b = seed, c = 0x9e3779b97f4a7c13ULL; /* :40 */
    const uint8_t *cur = s;
    uint64_t rem = len;
    while (rem >= 24) { /* :46-55 */
        a += load64le(cur);
        b += load64le(cur + 8);
        c += load64le(cur + 16);
        cur += 24;
        rem -= 24;
        jenkins_mix(&a, &b, &c);
    }
    c += len; /* :57 */
    /* :59-86: the fall-through switch adds byte i of the tail at bit 8*i of
     * (a,b) and at bit 8*(i-16)+8 of c (the low byte of c is the length). */
    for (uint64_t i = 0; i < rem; ++i) {
        uint64_t v = cur[i];
        if (i < 8) a += v << (8 * i);
        else if (i < 16) b += v << (8 * (i - 8));
        else c += v << (8 * (i - 16) + 8);
    }
    jenkins_mix(&a, &b, &c); /* :88 */
    out[0] = a; out[1] = b; out[2] = c;
}

/* ========================================================================= */
/* emphf: bit-pair vector + rank  (bitpair_vector.hpp:46-49,                  */
/*        ranked_bitpair_vector.hpp:47-62, 92-106)                           */
/* ========================================================================= */

static inline uint64_t nonzero_pairs(uint64_t x) { /* ranked_bitpair_vector.hpp:92-106 */
    const uint64_t ones_step_4 = 0x1111111111111111ULL;
    const uint64_t ones_step_8 = 0x0101010101010101ULL;
    x = (x | (x >> 1)) & (0x5 * ones_step_4);
    /* EMPHF_USE_POPCOUNT == 0 (emphf_config.hpp:3-6): SWAR population count */
    x = (x & 3 * ones_step_4) + ((x >> 2) & 3 * ones_step_4);
    x = (x + (x >> 4)) & 0x0f * ones_step_8;
    return (x * ones_step_8) >> 56;
}

static inline uint64_t bv_get(const orc_mphf *m, uint64_t pos) { /* bitpair_vector.hpp:46-49 */
    return (m->words[pos / 32] >> ((pos % 32) * 2)) & 3;
}

static uint64_t bv_rank(const orc_mphf *m, uint64_t pos) { /* ranked_bitpair_vector.hpp:47-62 */
    uint64_t word_idx = pos / 32;
    uint64_t word_offset = pos % 32;
    uint64_t block = pos / 512;
    uint64_t r = m->block_ranks[block];
    for (uint64_t w = block * 512 / 32; w < word_idx; ++w) r += nonzero_pairs(m->words[w]);
    uint64_t mask = ((uint64_t)1 << (word_offset * 2)) - 1;
    r += nonzero_pairs(m->words[word_idx] & mask);
    return r;
}

uint64_t orc_mphf_lookup(const orc_mphf *m, const uint8_t *s, uint64_t len) { /* mphf.hpp:79-89 */
    uint64_t h[3];
    orc_jenkins64(m->seed, s, len, h);
    uint64_t nodes[3] = {h[0] % m->hash_domain, m->hash_domain + (h[1] % m->hash_domain),
                         2 * m->hash_domain + (h[2] % m->hash_domain)};
    uint64_t hidx = (bv_get(m, nodes[0]) + bv_get(m, nodes[1]) + bv_get(m, nodes[2])) % 3;
    return bv_rank(m, nodes[hidx]);
}

void orc_mphf_lookup_batch(const orc_mphf *m, const uint8_t *recs, uint64_t stride,
                           const uint8_t *lens, uint64_t q, uint64_t *out, int threads) {
    (void)threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t i = 0; i < (int64_t)q; ++i)
        out[i] = orc_mphf_lookup(m, recs + (uint64_t)i * stride, lens ? lens[i] : stride);
}

/* .pf layout (mphf.hpp:99-113, base_hash.hpp:111-119, bitpair_vector.hpp:96-107,
 * ranked_bitpair_vector.hpp:70-84): u64 n, u64 hash_domain, u64 seed,
 * u64 bv_size, u64 words[(bv_size+31)/32], u64 block_ranks[(bv_size+511)/512]. */
orc_mphf *orc_mphf_load(const char *pf_path) {
    FILE *f = fopen(pf_path, "rb");
    if (!f) return NULL;
    orc_mphf *m = (orc_mphf *)calloc(1, sizeof(orc_mphf));
    uint64_t hdr[4];
    if (fread(hdr, 8, 4, f) != 4) goto fail;
    m->n = hdr[0]; m->hash_domain = hdr[1]; m->seed = hdr[2]; m->bv_size = hdr[3];
    m->n_words = (m->bv_size + 31) / 32;
    m->n_blocks = (m->bv_size + 511) / 512;
    m->words = (uint64_t *)malloc(8 * (m->n_words ? m->n_words : 1));
    m->block_ranks = (uint64_t *)malloc(8 * (m->n_blocks ? m->n_blocks : 1));
    if (fread(m->words, 8, m->n_words, f) != m->n_words) goto fail;
    if (fread(m->block_ranks, 8, m->n_blocks, f) != m->n_blocks) goto fail;
    fclose(f);
    return m;
fail:
    fclose(f);
    orc_mphf_free(m);
    return NULL;
}

orc_mphf *orc_mphf_from_arrays(uint64_t n, uint64_t hash_domain, uint64_t seed,
                               const uint64_t *words, uint64_t n_words,
                               const uint64_t *block_ranks, uint64_t n_blocks) {
    orc_mphf *m = (orc_mphf *)calloc(1, sizeof(orc_mphf));
    m->n = n; m->hash_domain = hash_domain; m->seed = seed; m->bv_size = 3 * hash_domain;
    m->n_words = n_words; m->n_blocks = n_blocks;
    m->words = (uint64_t *)malloc(8 * (n_words ? n_words : 1));
    m->block_ranks = (uint64_t *)malloc(8 * (n_blocks ? n_blocks : 1));
    memcpy(m->words, words, 8 * n_words);
    memcpy(m->block_ranks, block_ranks, 8 * n_blocks);
    return m;
}

int orc_mphf_save(const orc_mphf *m, const char *pf_path) {
    FILE *f = fopen(pf_path, "wb");
    if (!f) return -1;
    uint64_t hdr[4] = {m->n, m->hash_domain, m->seed, m->bv_size};
    fwrite(hdr, 8, 4, f);
    fwrite(m->words, 8, m->n_words, f);
    fwrite(m->block_ranks, 8, m->n_blocks, f);
    fclose(f);
    return 0;
}

void orc_mphf_free(orc_mphf *m) {
    if (!m) return;
    free(m->words);
    free(m->block_ranks);
    free(m);
}

/* ========================================================================= */
/* codec  (src/kmers.cpp)                                                    */
/* ========================================================================= */

/* kmers.cpp:12-40: always reads 23 chars; anything but ACGT adds 0.  For a
 * shorter std::string the reference reads the terminator (0) at [len]; chars
 * past that are undefined in the reference -- the oracle (and the CUDA path)
 * define them as 0 as well. */
uint64_t orc_dna23_bitset(const uint8_t *s, uint64_t len) {
    uint64_t num = 0;
    for (int n = 0; n < 23; n++) {
        uint8_t c = (uint64_t)n < len ? s[n] : 0;
        num = num << 2;
        if (c == 'A') num += 0;
        if (c == 'C') num += 1;
        if (c == 'G') num += 2;
        if (c == 'T') num += 3;
    }
    return num;
}

uint32_t orc_dna13_bitset(const uint8_t *s, uint64_t len) { /* kmers.cpp:42-55 */
    uint32_t num = 0;
    for (int n = 0; n < 13; n++) {
        uint8_t c = (uint64_t)n < len ? s[n] : 0;
        num = num << 2;
        if (c == 'A') num += 0;
        if (c == 'C') num += 1;
        if (c == 'G') num += 2;
        if (c == 'T') num += 3;
    }
    return num;
}

void orc_bitset_dna23(uint64_t x, uint8_t *out, int k) { /* kmers.cpp:89-114 */
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int i = k - 1; i >= 0; i--) {
        out[i] = (uint8_t)L[x & 3];
        x >>= 2;
    }
}

void orc_bitset_dna13(uint32_t x, uint8_t *out, int k) { /* kmers.cpp:174-199 */
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int i = k - 1; i >= 0; i--) {
        out[i] = (uint8_t)L[x & 3];
        x >>= 2;
    }
}

static uint64_t reverse_pairs64(uint64_t num) { /* kmers.cpp:355-363 */
    uint64_t count = 8 * sizeof num - 2;
    uint64_t reverse_num = num;
    for (num >>= 2; num; num >>= 2) {
        reverse_num <<= 2;
        reverse_num |= num & 3;
        count -= 2;
    }
    return reverse_num << count;
}

static uint32_t reverse_pairs32(uint32_t num) { /* kmers.cpp:365-373 */
    uint32_t count = 8 * sizeof num - 2;
    uint32_t reverse_num = num;
    for (num >>= 2; num; num >>= 2) {
        reverse_num <<= 2;
        reverse_num |= num & 3;
        count -= 2;
    }
    return reverse_num << count;
}

uint64_t orc_reverse_dna23(uint64_t x) { return (~reverse_pairs64(x)) >> 18; } /* :376-381 */
uint32_t orc_reverse_dna13(uint32_t x) { return (~reverse_pairs32(x)) >> 6; }  /* :383-388 */

void orc_dna_bitset_pack(const uint8_t *s, uint64_t len, uint8_t *out) { /* dna_bitseq.hpp:22-61 */
    uint64_t nbytes = (len / 4) + (len % 4 != 0);
    memset(out, 0, nbytes);
    for (uint64_t i = 0; i < len; i++) {
        uint8_t shift = (uint8_t)(6 - 2 * (i % 4));
        uint8_t code = 0; /* default: BASE_A */
        switch (s[i]) {
            case 'A': code = 0; break;
            case 'C': code = 1; break;
            case 'G': code = 2; break;
            case 'T': code = 3; break;
            default: code = 0; break;
        }
        out[i / 4] |= (uint8_t)(code << shift);
    }
}

uint64_t orc_dna_bitset_ukmer(const uint8_t *packed, uint64_t pos, int k) { /* dna_bitseq.hpp:124-151 */
    uint64_t num = 0;
    for (int i = 0; i < k; ++i) {
        uint8_t shift = (uint8_t)(6 - 2 * ((pos + i) % 4));
        uint8_t base = (uint8_t)((packed[(i + pos) / 4] >> shift) & 3);
        num = (num << 2) + base;
    }
    return num;
}

/* ========================================================================= */
/* 23-mer queries  (src/python_wrapper.cpp, src/hash.hpp)                    */
/* ========================================================================= */

/* python_wrapper.cpp:610-627 get_tf_value_23mer; the same probe sequence is
 * used by get_kid_by_kmer (:700-716) and get_strand (:726-742).
 * returns 1 forward hit, 2 reverse hit, 0 not found; *h_out = bucket. */
static int probe23(const orc_index23 *ix, const uint8_t *s, uint64_t len, uint64_t *h_out) {
    uint64_t ukmer = orc_dna23_bitset(s, len);
    uint64_t h1 = orc_mphf_lookup(ix->mphf, s, len); /* hashes the RAW bytes, full length */
    if (h1 >= ix->n || ix->checker[h1] != ukmer) {
        uint8_t rev[23];
        uint64_t urev = orc_reverse_dna23(ukmer);
        orc_bitset_dna23(urev, rev, 23);
        uint64_t h2 = orc_mphf_lookup(ix->mphf, rev, 23);
        if (h2 >= ix->n || ix->checker[h2] != urev) return 0;
        *h_out = h2;
        return 2;
    }
    *h_out = h1;
    return 1;
}

uint32_t orc_tf23(const orc_index23 *ix, const uint8_t *s, uint64_t len) {
    uint64_t h;
    return probe23(ix, s, len, &h) ? ix->tf[h] : 0;
}

uint64_t orc_kid23(const orc_index23 *ix, const uint8_t *s, uint64_t len) { /* :700-716 */
    uint64_t h;
    return probe23(ix, s, len, &h) ? h : 0;
}

uint64_t orc_strand23(const orc_index23 *ix, const uint8_t *s, uint64_t len) { /* :726-742 */
    uint64_t h;
    return (uint64_t)probe23(ix, s, len, &h);
}

uint32_t orc_get_freq23(const orc_index23 *ix, uint64_t kmer) { /* hash.hpp:123-140 */
    uint8_t buf[23];
    orc_bitset_dna23(kmer, buf, 23);
    uint64_t h1 = orc_mphf_lookup(ix->mphf, buf, 23);
    if (h1 < ix->n && ix->checker[h1] == kmer) return ix->tf[h1];
    uint64_t rev = orc_reverse_dna23(kmer);
    orc_bitset_dna23(rev, buf, 23);
    uint64_t h2 = orc_mphf_lookup(ix->mphf, buf, 23);
    if (h2 < ix->n && ix->checker[h2] == rev) return ix->tf[h2];
    return 0;
}

/* python_wrapper.cpp:1230-1246 (len != 23 -> 0; fwd + tf(revcomp string)) */
void orc_both_tf23(const orc_index23 *ix, const uint8_t *s, uint64_t len, uint32_t out[2]) {
    out[0] = out[1] = 0;
    if (len != 23) return; /* :1258-1260 */
    out[0] = orc_tf23(ix, s, len);
    uint8_t rev[23];
    orc_bitset_dna23(orc_reverse_dna23(orc_dna23_bitset(s, len)), rev, 23);
    out[1] = orc_tf23(ix, rev, 23);
}

uint64_t orc_total_tf23(const orc_index23 *ix, const uint8_t *s, uint64_t len) {
    uint32_t b[2];
    orc_both_tf23(ix, s, len, b);
    return (uint64_t)b[0] + (uint64_t)b[1];
}

/* hash.hpp:150-170 get_pfid: bucket of the lexicographically smaller of the
 * RAW string and the decoded reverse complement; n if absent. */
uint64_t orc_pfid23(const orc_index23 *ix, const uint8_t *s, uint64_t len) {
    uint64_t kmer = orc_dna23_bitset(s, len);
    uint8_t rev[23];
    uint64_t rkmer = orc_reverse_dna23(kmer);
    orc_bitset_dna23(rkmer, rev, 23);
    /* std::string_view::compare: memcmp over min length, then length */
    uint64_t ml = len < 23 ? len : 23;
    int c = memcmp(s, rev, ml);
    if (c == 0) c = (len < 23) ? -1 : (len > 23 ? 1 : 0);
    if (c <= 0) {
        uint64_t h1 = orc_mphf_lookup(ix->mphf, s, len);
        return (h1 < ix->n && ix->checker[h1] == kmer) ? h1 : ix->n;
    } else {
        uint64_t h1 = orc_mphf_lookup(ix->mphf, rev, 23);
        return (h1 < ix->n && ix->checker[h1] == rkmer) ? h1 : ix->n;
    }
}

void orc_tf23_batch(const orc_index23 *ix, const uint8_t *recs, uint64_t stride,
                    const uint8_t *lens, uint64_t q, int mode, void *out, int threads) {
    (void)threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t i = 0; i < (int64_t)q; ++i) {
        const uint8_t *s = recs + (uint64_t)i * stride;
        uint64_t len = lens ? lens[i] : stride;
        switch (mode) {
            case 0: ((uint32_t *)out)[i] = orc_tf23(ix, s, len); break;
            case 1: ((uint64_t *)out)[i] = orc_total_tf23(ix, s, len); break;
            case 2: orc_both_tf23(ix, s, len, ((uint32_t *)out) + 2 * i); break;
            case 3: ((uint64_t *)out)[i] = orc_pfid23(ix, s, len); break;
            case 4: ((uint64_t *)out)[i] = orc_strand23(ix, s, len); break;
            case 5: ((uint64_t *)out)[i] = orc_kid23(ix, s, len); break;
            default: break;
        }
    }
}

/* ========================================================================= */
/* 13-mer queries  (src/python_wrapper.cpp:482-608, 938-980)                 */
/* ========================================================================= */

static int all_acgt(const uint8_t *s, uint64_t len) {
    for (uint64_t i = 0; i < len; ++i)
        if (s[i] != 'A' && s[i] != 'T' && s[i] != 'G' && s[i] != 'C') return 0;
    return 1;
}

uint32_t orc_tf13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *s, uint64_t len) {
    if (len != 13) return 0;      /* :483, :954 */
    if (!all_acgt(s, len)) return 0; /* :488-492, :960-966 */
    uint64_t id = orc_mphf_lookup(m, s, len);
    if (id < ORC_TOTAL_13MERS) return (uint32_t)tf64[id]; /* :498-500 (uint64 -> uint32) */
    return 0;
}

/* :505-517 string reverse complement: non-ACGT chars stay as they are */
static void revcomp13_str(const uint8_t *s, uint64_t len, uint8_t *out) {
    for (uint64_t i = 0; i < len; ++i) {
        uint8_t c = s[len - 1 - i];
        switch (c) {
            case 'A': c = 'T'; break;
            case 'T': c = 'A'; break;
            case 'G': c = 'C'; break;
            case 'C': c = 'G'; break;
            default: break;
        }
        out[i] = c;
    }
}

/* :522-545 / :567-590: len != 13 -> 0; NO validity check; an id >= 4^13 reads
 * out of bounds in the reference -- defined as 0 here and in the CUDA path. */
void orc_both_tf13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *s, uint64_t len,
                   uint64_t out[2]) {
    out[0] = out[1] = 0;
    if (len != 13) return;
    uint8_t rc[13];
    uint64_t id = orc_mphf_lookup(m, s, 13);
    out[0] = id < ORC_TOTAL_13MERS ? tf64[id] : 0;
    revcomp13_str(s, 13, rc);
    uint64_t rid = orc_mphf_lookup(m, rc, 13);
    out[1] = rid < ORC_TOTAL_13MERS ? tf64[rid] : 0;
}

void orc_tf13_batch(const orc_mphf *m, const uint64_t *tf64, const uint8_t *recs,
                    uint64_t stride, const uint8_t *lens, uint64_t q, int mode, void *out,
                    int threads) {
    (void)threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t i = 0; i < (int64_t)q; ++i) {
        const uint8_t *s = recs + (uint64_t)i * stride;
        uint64_t len = lens ? lens[i] : stride;
        uint64_t b[2];
        switch (mode) {
            case 0: ((uint32_t *)out)[i] = orc_tf13(m, tf64, s, len); break;
            case 1: orc_both_tf13(m, tf64, s, len, b); ((uint64_t *)out)[i] = b[0] + b[1]; break;
            case 2: orc_both_tf13(m, tf64, s, len, ((uint64_t *)out) + 2 * i); break;
            default: break;
        }
    }
}

/* ========================================================================= */
/* 13-mer counting  (src/count_kmers13.cpp)                                  */
/* ========================================================================= */

int orc_detect_format(const uint8_t *bytes, uint64_t len) { /* :194-206 */
    if (len == 0) return 0;
    if (bytes[0] == '\n') return 0; /* empty first line -> PLAIN */
    if (bytes[0] == '>') return 1;
    if (bytes[0] == '@') return 2;
    return 0;
}

typedef void (*seq_cb)(const uint8_t *seq, uint64_t len, void *ud);

/* Sequence producers: read_plain_file :262-272, read_fastq_file :240-257,
 * read_fasta_file :211-235.  std::getline splits on '\n' only. */
static void for_each_sequence(const uint8_t *bytes, uint64_t len, int fmt, seq_cb cb, void *ud) {
    uint64_t pos = 0;
    uint64_t line_no = 0;
    uint8_t *acc = NULL;
    uint64_t acc_len = 0, acc_cap = 0;
    while (pos < len) {
        uint64_t e = pos;
        while (e < len && bytes[e] != '\n') ++e;
        const uint8_t *line = bytes + pos;
        uint64_t ll = e - pos;
        if (fmt == 0) {
            if (ll) cb(line, ll, ud);
        } else if (fmt == 2) {
            if (line_no % 4 == 1 && ll) cb(line, ll, ud);
        } else {
            if (ll) {
                if (line[0] == '>') {
                    if (acc_len) cb(acc, acc_len, ud);
                    acc_len = 0;
                } else {
                    if (acc_len + ll > acc_cap) {
                        acc_cap = (acc_len + ll) * 2 + 64;
                        acc = (uint8_t *)realloc(acc, acc_cap);
                    }
                    memcpy(acc + acc_len, line, ll);
                    acc_len += ll;
                }
            }
        }
        line_no++;
        pos = e + 1;
    }
    if (fmt == 1 && acc_len) cb(acc, acc_len, ud);
    free(acc);
}

typedef struct {
    const orc_mphf *m; /* NULL -> direct address */
    uint64_t *counts;
    orc_count_stats st;
} count_ud;

static inline uint8_t norm_base(uint8_t c) { /* normalize_sequence :113-126 (C-locale toupper) */
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    if (c == 'A' || c == 'T' || c == 'G' || c == 'C') return c;
    return 'N';
}

static void count_sequence(const uint8_t *seq, uint64_t len, void *ud_) { /* process_sequence :131-161 */
    count_ud *ud = (count_ud *)ud_;
    if (len < 13) return; /* :132 */
    ud->st.sequences++;
    uint8_t km[13];
    for (uint64_t i = 0; i + 13 <= len; ++i) {
        int ok = 1;
        for (int j = 0; j < 13; ++j) {
            km[j] = norm_base(seq[i + j]);
            if (km[j] == 'N') ok = 0;
        }
        ud->st.windows++;
        if (ok) {
            uint64_t idx = ud->m ? orc_mphf_lookup(ud->m, km, 13) : (uint64_t)orc_dna13_bitset(km, 13);
            if (idx < ORC_TOTAL_13MERS) {
                ud->counts[idx]++;
                ud->st.valid++;
            } else {
                ud->st.invalid++;
            }
        } else {
            ud->st.invalid++;
        }
    }
}

void orc_count13(const orc_mphf *m, const uint8_t *bytes, uint64_t len, int fmt,
                 uint64_t *counts, orc_count_stats *st) {
    count_ud ud;
    memset(&ud, 0, sizeof ud);
    ud.m = m;
    ud.counts = counts;
    if (fmt < 0) fmt = orc_detect_format(bytes, len);
    for_each_sequence(bytes, len, fmt, count_sequence, &ud);
    if (st) *st = ud.st;
}

void orc_count13_direct(const uint8_t *bytes, uint64_t len, int fmt, uint64_t *hist,
                        orc_count_stats *st) {
    orc_count13(NULL, bytes, len, fmt, hist, st);
}

/* ========================================================================= */
/* coverage  (aindex/core/aindex.py:314-322)                                 */
/* ========================================================================= */

void orc_coverage23(const orc_index23 *ix, const uint8_t *seq, uint64_t len, uint32_t cutoff,
                    uint32_t *out) {
    if (len < 23) return;
    for (uint64_t i = 0; i + 23 <= len; ++i) {
        uint32_t tf = orc_tf23(ix, seq + i, 23);
        out[i] = tf >= cutoff ? tf : 0;
    }
}

void orc_coverage13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *seq, uint64_t len,
                    uint32_t cutoff, uint32_t *out) {
    if (len < 13) return;
    for (uint64_t i = 0; i + 13 <= len; ++i) {
        uint32_t tf = orc_tf13(m, tf64, seq + i, 13);
        out[i] = tf >= cutoff ? tf : 0;
    }
}

/* ========================================================================= */
/* positions index                                                           */
/* ========================================================================= */

/* worker prologue hash.cpp:973-988 / compute_aindex13.cpp:133-147 */
static uint64_t worker_first_start(const uint8_t *c, uint64_t start, uint64_t end, uint64_t k) {
    while (start + k <= end) { /* start < end-k+1 */
        int found = 0;
        for (uint64_t i = start; i < start + k; ++i) {
            if (c[i] == '\n' || c[i] == '~' || c[i] == '?') {
                start = i + 1;
                found = 1;
                break;
            }
        }
        if (!found) break;
    }
    return start;
}

void orc_positions_build23(const orc_index23 *ix, const uint8_t *reads, uint64_t len,
                           uint64_t *indices, uint64_t *positions) {
    const uint64_t k = 23;
    uint64_t n = ix->n;
    indices[0] = 0; /* hash.hpp:373-378 */
    for (uint64_t i = 1; i < n + 1; ++i) indices[i] = indices[i - 1] + ix->tf[i - 1];
    uint64_t total = indices[n];
    memset(positions, 0, 8 * total);
    uint64_t *cursor = (uint64_t *)calloc(n ? n : 1, 8); /* ppositions, hash.hpp:384 */
    if (len >= k) {
        uint64_t start = worker_first_start(reads, 0, len, k);
        for (uint64_t i = start; i + k <= len; ++i) { /* hash.cpp:990 */
            int skip = 0;
            for (uint64_t j = 0; j < k; ++j) { /* :1006-1015 */
                uint8_t ch = reads[i + j];
                if (ch == '\n' || ch == '~' || ch == 'N') { skip = 1; break; }
            }
            if (skip) continue;
            const uint8_t *kmer = reads + i;
            uint64_t ukmer = orc_dna23_bitset(kmer, 23);
            uint64_t urev = orc_reverse_dna23(ukmer);
            uint64_t h1;
            if (ukmer <= urev) { /* :1032-1041 */
                h1 = orc_mphf_lookup(ix->mphf, kmer, 23);
                if (h1 >= n || ix->checker[h1] != ukmer) continue;
            } else { /* :1042-1051 */
                uint8_t rev[23];
                orc_bitset_dna23(urev, rev, 23);
                h1 = orc_mphf_lookup(ix->mphf, rev, 23);
                if (h1 >= n || ix->checker[h1] != urev) continue;
            }
            uint64_t h2 = cursor[h1]++;
            if (h2 >= ix->tf[h1]) continue;
            positions[indices[h1] + h2] = i + 1;
        }
    }
    free(cursor);
}

void orc_positions_build13(const orc_mphf *m, const uint64_t *tf64, const uint8_t *reads,
                           uint64_t len, uint64_t *indices, uint64_t *positions) {
    const uint64_t k = 13;
    const uint64_t n = ORC_TOTAL_13MERS;
    indices[0] = 0; /* compute_aindex13.cpp:58-64 */
    for (uint64_t i = 1; i < n + 1; ++i) indices[i] = indices[i - 1] + tf64[i - 1];
    uint64_t total = indices[n];
    memset(positions, 0, 8 * total);
    uint64_t *cursor = (uint64_t *)calloc(n, 8);
    if (len >= k) {
        uint64_t start = worker_first_start(reads, 0, len, k);
        for (uint64_t i = start; i + k <= len; ++i) {
            int skip = 0;
            for (uint64_t j = 0; j < k; ++j) { /* :186-193 */
                uint8_t c = reads[i + j];
                if (c != 'A' && c != 'T' && c != 'G' && c != 'C') { skip = 1; break; }
            }
            if (skip) continue;
            uint64_t h = orc_mphf_lookup(m, reads + i, 13);
            if (h < n) { /* :208-216 */
                uint64_t pos_idx = cursor[h]++;
                uint64_t array_idx = indices[h] + pos_idx;
                if (array_idx < total && pos_idx < (indices[h + 1] - indices[h]))
                    positions[array_idx] = i + 1;
            }
        }
    }
    free(cursor);
}

uint64_t orc_positions_query23(const orc_index23 *ix, const uint64_t *indices,
                               const uint64_t *positions, const uint8_t *s, uint64_t len,
                               uint64_t *out, uint64_t cap) {
    if (len != 23) return 0; /* dispatcher python_wrapper.cpp:826-831 */
    uint64_t h1 = orc_pfid23(ix, s, len);
    if (h1 >= ix->n) return 0; /* defect 2.3#6: the reference aborts here */
    uint64_t cnt = 0;
    for (uint64_t p = indices[h1]; p < indices[h1 + 1]; ++p) /* :812-819 */
        if (positions[p] && cnt < cap) out[cnt++] = positions[p] - 1;
    return cnt;
}

uint64_t orc_positions_query13(const orc_mphf *m, const uint64_t *indices,
                               const uint64_t *positions, uint64_t n_positions, const uint8_t *s,
                               uint64_t len, uint64_t *out, uint64_t cap) {
    if (len != 13 || !all_acgt(s, len)) return 0; /* :1073-1082 */
    uint64_t h = orc_mphf_lookup(m, s, 13);
    uint64_t cnt = 0;
    if (h < ORC_TOTAL_13MERS) { /* :1088-1098 */
        for (uint64_t i = indices[h]; i < indices[h + 1] && i < n_positions; ++i)
            if (positions[i] > 0 && cnt < cap) out[cnt++] = positions[i] - 1;
    }
    return cnt;
}

/* ========================================================================= */
/* canonical 23-mer table: tests/analyze_kmers.py:25-33, scripts/compute_aindex.py:164-182 */
/* ========================================================================= */

void orc_free(void *p) { free(p); }

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* canonical values of the windows [lo, hi) of the image, rolled: fwd = (fwd << 2 | code) & mask,
 * rc = rc >> 2 | (3 - code) << 44 (identical to get_dna23_bitset + reverseDNA, kmers.cpp:12-25, 376-381,
 * for ACGT-only windows -- checked against those functions in tests/test_oracle_golden.py);
 * cb(value, arg) is called for every ACGT-only window */
typedef void (*c23_cb)(uint64_t c, void *arg);
static void canonical23_roll(const uint8_t *s, uint64_t lo, uint64_t hi, c23_cb cb, void *arg) {
    const uint64_t mask = (1ULL << 46) - 1;
    uint64_t f = 0, r = 0;
    int run = 0; /* ACGT characters in a row ending at the current one */
    if (lo >= hi) return;
    for (uint64_t i = lo; i < hi + 22; ++i) {
        uint64_t c;
        switch (s[i]) {
        case 'A': c = 0; break;
        case 'C': c = 1; break;
        case 'G': c = 2; break;
        case 'T': c = 3; break;
        default: run = 0; f = 0; r = 0; continue;
        }
        f = ((f << 2) | c) & mask;
        r = (r >> 2) | ((3 - c) << 44);
        if (++run >= 23) cb(f <= r ? f : r, arg);
    }
}
static void c23_count_cb(uint64_t c, void *arg) { ((uint64_t *)arg)[c >> 34]++; }
struct c23_scatter { uint64_t *cur, *all; };
static void c23_scatter_cb(uint64_t c, void *arg) {
    struct c23_scatter *sc = (struct c23_scatter *)arg;
    sc->all[sc->cur[c >> 34]++] = c;
}

#define C23_BINS 4096 /* top 12 of the 46 bits */

uint64_t orc_canonical23_count(const uint8_t *reads, uint64_t len, int threads, uint64_t **kmers_out,
                               uint32_t **counts_out) {
    *kmers_out = NULL;
    *counts_out = NULL;
    if (len < 23) return 0;
    if (threads < 1) threads = 1;
    const uint64_t n_win = len - 22;
    /* pass 1: windows per bin; pass 2: scatter into bins; then every bin is sorted and run-length encoded */
    uint64_t *bin_n = (uint64_t *)calloc((size_t)threads * C23_BINS, 8);
    uint64_t *bin_off = (uint64_t *)calloc(C23_BINS + 1, 8);
    const uint64_t chunk = (n_win + (uint64_t)threads - 1) / (uint64_t)threads;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < n_win ? lo + chunk : n_win;
        canonical23_roll(reads, lo, hi, c23_count_cb, bin_n + (size_t)t * C23_BINS);
    }
    /* bin b: thread 0's part, thread 1's part, ...  (bin_n becomes the write cursor of every (thread, bin)) */
    uint64_t total = 0;
    for (int b = 0; b < C23_BINS; ++b) {
        bin_off[b] = total;
        for (int t = 0; t < threads; ++t) {
            uint64_t c = bin_n[(size_t)t * C23_BINS + b];
            bin_n[(size_t)t * C23_BINS + b] = total;
            total += c;
        }
    }
    bin_off[C23_BINS] = total;
    uint64_t *all = (uint64_t *)malloc((total ? total : 1) * 8);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < n_win ? lo + chunk : n_win;
        struct c23_scatter sc = {bin_n + (size_t)t * C23_BINS, all};
        canonical23_roll(reads, lo, hi, c23_scatter_cb, &sc);
    }
    uint64_t *uniq_n = (uint64_t *)calloc(C23_BINS + 1, 8);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
    for (int b = 0; b < C23_BINS; ++b) {
        uint64_t *a = all + bin_off[b], m = bin_off[b + 1] - bin_off[b], u = 0;
        qsort(a, m, 8, cmp_u64);
        for (uint64_t i = 0; i < m; ++i)
            if (i == 0 || a[i] != a[i - 1]) ++u;
        uniq_n[b] = u;
    }
    uint64_t n = 0;
    for (int b = 0; b < C23_BINS; ++b) {
        uint64_t u = uniq_n[b];
        uniq_n[b] = n;
        n += u;
    }
    uint64_t *kmers = (uint64_t *)malloc((n ? n : 1) * 8);
    uint32_t *counts = (uint32_t *)malloc((n ? n : 1) * 4);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
    for (int b = 0; b < C23_BINS; ++b) {
        const uint64_t *a = all + bin_off[b];
        uint64_t m = bin_off[b + 1] - bin_off[b], w = uniq_n[b];
        for (uint64_t i = 0; i < m;) {
            uint64_t j = i + 1;
            while (j < m && a[j] == a[i]) ++j;
            kmers[w] = a[i];
            counts[w] = (uint32_t)(j - i);
            ++w;
            i = j;
        }
    }
    free(all); free(bin_n); free(bin_off); free(uniq_n);
    *kmers_out = kmers;
    *counts_out = counts;
    return n;
}

int orc_write_dat(const uint64_t *kmers, const uint32_t *counts, uint64_t n, const char *dat_path,
                  const char *keys_path) {
    FILE *fd = dat_path ? fopen(dat_path, "wb") : NULL, *fk = keys_path ? fopen(keys_path, "wb") : NULL;
    if ((dat_path && !fd) || (keys_path && !fk)) {
        if (fd) fclose(fd);
        if (fk) fclose(fk);
        return -1;
    }
    static char big1[1 << 20], big2[1 << 20];
    if (fd) setvbuf(fd, big1, _IOFBF, sizeof big1);
    if (fk) setvbuf(fk, big2, _IOFBF, sizeof big2);
    char line[48];
    int ok = 1;
    for (uint64_t i = 0; i < n && ok; ++i) {
        orc_bitset_dna23(kmers[i], (uint8_t *)line, 23);
        if (fk) {
            line[23] = '\n';
            ok = fwrite(line, 1, 24, fk) == 24;
        }
        if (fd && ok) {
            int m = 23 + snprintf(line + 23, sizeof line - 23, "\t%u\n", counts ? counts[i] : 1u);
            ok = fwrite(line, 1, (size_t)m, fd) == (size_t)m;
        }
    }
    if (fd) ok = (fclose(fd) == 0) && ok;
    if (fk) ok = (fclose(fk) == 0) && ok;
    return ok ? 0 : -1;
}
