// device_common.cuh -- device-side building blocks shared by every kernel of
// libaindex_cuda (sm_100a): 2-bit codec, Jenkins lookup8 triple hash, emphf MPHF
// evaluation on the B200 layout, exact fast modulo.
//
// Reference semantics (ad3002/aindex):
//   codec     src/kmers.cpp:12-85 (encode), :89-257 (decode), :355-388 (reverseDNA)
//   hash      src/emphf/base_hash.hpp:38-91, mix :127-145
//   lookup    src/emphf/mphf.hpp:79-89, bitpair_vector.hpp:46-49,
//             ranked_bitpair_vector.hpp:47-62
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aix {

// ---------------------------------------------------------------------------------
// B200 layout of the MPHF.  The reference stores words[] and a rank sample every 512
// pairs, so rank() walks up to 16 words.  Here the bit-pair vector is cut into 16-byte
// records that carry the rank of their first pair: one 16-byte load yields both the 2-bit
// value and everything rank() needs, so a lookup is three independent 16-byte loads and no
// dependent fourth access.  Two record shapes:
//   compact (bv_size < 2^32, i.e. up to 1.16 G keys): 48 pairs + u32 rank   -> 0.41 B / key
//   wide    (anything larger):                         32 pairs + u64 rank   -> 0.61 B / key
// The compact form keeps the C2 structure (50 M keys) at 20.5 MB, which together with the
// fingerprint tier stays inside the ~60 MB of L2 that random accesses from all SMs can use.
// ---------------------------------------------------------------------------------
struct MphfDev {
    uint64_t n;
    uint64_t hash_domain;
    uint64_t seed;
    uint64_t magic;     // floor(2^64 / hash_domain)
    const ulonglong2 *recs;  // wide:    recs[w] = { words[w], rank of pair 32*w }
    const uint4 *crecs;      // compact: crecs[r] = { pairs 48r..48r+47 (96 bits), rank of pair 48*r }; nullptr = wide
    // fused (owned by an aix_index23, nullptr otherwise): frecs[r] = { pairs 16r..16r+15 (32 bits), 16 x 4-bit
    // fingerprint of the key assigned to each node (64 bits), rank of pair 16*r }.  The membership filter rides in
    // the record the lookup loads anyway: 3 scattered L2 requests per query instead of 4 (L1TEX was the limiter at
    // 85 %, profiles/r01_tf23_ncu.txt), node / 16 is a shift, and rank needs one popcount.  1.23 B / key.
    const uint4 *frecs;
};
// mphf_eval on the fused layout returns the id with the fingerprint of the chosen node in the top byte; probe23
// strips it (its h argument is a reference) before anyone else sees the id.
constexpr uint64_t kFusedTagShift = 56;
constexpr uint64_t kFusedIdMask = (1ULL << kFusedTagShift) - 1;

struct Index23Dev {
    uint64_t n;
    int canonical_only;
    const uint4 *recs;  // recs[h] = { checker lo, checker hi, tf, 0 }: one sector per probe
    // Optional L2-resident filter tier: fp[h] = fingerprint8(checker[h]).  A probe whose
    // fingerprint differs cannot verify, so it never touches the HBM record: on miss-dominated
    // batches (the reference's own stress workload) 255/256 of the random HBM probes disappear.
    // Exact: fp mismatch => checker mismatch; a match is always confirmed on the full record.
    const uint8_t *fp;  // nullptr = tier disabled (index too large to keep it in L2)
    int fp_bits;        // 8: fp[h]; 4: nibble (h & 1) of fp[h >> 1] (half the L2 footprint, 1/16 false positives)
    // Optional membership filter in FRONT of the MPHF (canonical-only indexes): a blocked Bloom filter, one 64-bit word
    // per key (two bits in each half), ~8 bits per key.  A k-mer whose four bits are not all set is in no record: the
    // query is answered with ONE 8-byte L2 request and without the Jenkins hash, the three MPHF record loads and the
    // rank.  No false negatives (every stored k-mer set its bits), so answers do not change.  tf_query.cu decides per
    // batch whether the filter pays (miss-dominated batches) from the pass rate it observes.
    const uint2 *bloom;    // nullptr = no filter
    uint32_t bloom_words;
};

// word and bit masks of a canonical 23-mer code in the front filter (same functions build and test); split in two so that
// a kernel can issue the load of the word before it needs the masks
__device__ __forceinline__ void bloom_word(uint64_t c, uint32_t n_words, uint32_t &word, uint32_t &g) {
    uint32_t h = (uint32_t)c * 0x9E3779B1u + (uint32_t)(c >> 32) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    g = h * 0x297A2D39u;
    g ^= g >> 15;
    word = __umulhi(h, n_words);
}
__device__ __forceinline__ void bloom_masks(uint32_t g, uint32_t &mlo, uint32_t &mhi) {
    mlo = (1u << (g & 31u)) | (1u << ((g >> 5) & 31u));
    mhi = (1u << ((g >> 10) & 31u)) | (1u << ((g >> 15) & 31u));
}
__device__ __forceinline__ void bloom_slot(uint64_t c, uint32_t n_words, uint32_t &word, uint32_t &mlo, uint32_t &mhi) {
    uint32_t g;
    bloom_word(c, n_words, word, g);
    bloom_masks(g, mlo, mhi);
}

__device__ __host__ __forceinline__ uint32_t fingerprint8(uint64_t kmer) {
    uint32_t x = (uint32_t)kmer ^ (uint32_t)(kmer >> 23);
    return (x ^ (x >> 9)) & 0xFFu;
}
__device__ __host__ __forceinline__ uint32_t fingerprint4(uint64_t kmer) {
    uint32_t x = fingerprint8(kmer);
    return (x ^ (x >> 4)) & 0xFu;
}

// ---- cache-policy loads ---------------------------------------------------------------
// The MPHF records (0.46 B/key, ~31 MB for 50 M keys) are hit three times per query and must
// stay in L2; the {checker, tf} records (16 B/key, 0.8 GB) and the query bytes are touched
// once and must not evict them.  Without the hints the random index stream thrashes L2 and
// the MPHF loads go to DRAM (ncu r01a: 16.2 GB read per 100 M queries, L2 hit rate 30 %).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ ulonglong2 ld_evict_last_u64x2(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                 : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
__device__ __forceinline__ uint2 ld_evict_last_u32x2(const uint2 *p) {
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
__device__ __forceinline__ uint4 ld_evict_first_u32x4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_policy_evict_first()));
    return v;
}

// exact h % d for any 64-bit h.  magic = floor(2^64/d) = (2^64 - r0)/d with r0 < d, so
// h*magic/2^64 = h/d - h*r0/(d*2^64) > h/d - 1: q = mulhi(h, magic) is q_true or q_true-1 and
// h - q*d lies in [0, 2d): one conditional subtraction.
__device__ __forceinline__ uint64_t fastmod(uint64_t h, uint64_t d, uint64_t magic) {
    uint64_t q = __umul64hi(h, magic);
    uint64_t r = h - q * d;
    if (r >= d) r -= d;
    return r;
}
// the same for d < 2^31: the remainder before correction is below 2d <= 2^32, so it can be
// formed from the low halves alone (only the low 32 bits of the quotient are needed)
__device__ __forceinline__ uint32_t fastmod32(uint64_t h, uint32_t d, uint64_t magic) {
    uint32_t q = (uint32_t)__umul64hi(h, magic);
    uint32_t r = (uint32_t)h - q * d;
    if (r >= d) r -= d;
    return r;
}

// ---- Jenkins lookup8 (base_hash.hpp:127-145) ---------------------------------------
__device__ __forceinline__ void jenkins_mix(uint64_t &a, uint64_t &b, uint64_t &c) {
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
}

constexpr uint64_t kGolden = 0x9e3779b97f4a7c13ULL;

// hash of a string shorter than 24 bytes given as three little-endian words:
// w0 = bytes 0..7, w1 = bytes 8..15, w2 = bytes 16..22 (unused bytes zero)
__device__ __forceinline__ void jenkins_short(uint64_t seed, uint64_t w0, uint64_t w1, uint64_t w2,
                                              uint32_t len, uint64_t &a, uint64_t &b, uint64_t &c) {
    a = seed + w0;
    b = seed + w1;
    c = kGolden + len + (w2 << 8);  // base_hash.hpp:57-66: low byte of c holds the length
    jenkins_mix(a, b, c);
}

// general length, bytes addressed through p (global or shared); base_hash.hpp:38-91
__device__ __forceinline__ void jenkins_bytes(uint64_t seed, const uint8_t *p, uint32_t len,
                                              uint64_t &a, uint64_t &b, uint64_t &c) {
    a = seed; b = seed; c = kGolden;
    uint32_t rem = len;
    while (rem >= 24) {
        uint64_t w[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 24; ++i) w[i >> 3] |= (uint64_t)p[i] << (8 * (i & 7));
        a += w[0]; b += w[1]; c += w[2];
        jenkins_mix(a, b, c);
        p += 24; rem -= 24;
    }
    uint64_t w0 = 0, w1 = 0, w2 = 0;
    for (uint32_t i = 0; i < rem; ++i) {
        uint64_t v = p[i];
        if (i < 8) w0 |= v << (8 * i);
        else if (i < 16) w1 |= v << (8 * (i - 8));
        else w2 |= v << (8 * (i - 16));
    }
    a += w0; b += w1; c += (uint64_t)len + (w2 << 8);
    jenkins_mix(a, b, c);
}

// ---- emphf lookup on the B200 layout (mphf.hpp:79-89) -------------------------------
__device__ __forceinline__ uint32_t nonzero_pairs64(uint64_t x) {
    x = (x | (x >> 1)) & 0x5555555555555555ULL;
    return (uint32_t)__popcll(x);  // == the SWAR count of ranked_bitpair_vector.hpp:92-106
}

__device__ __forceinline__ uint32_t nonzero_pairs32(uint32_t x) { return (uint32_t)__popc((x | (x >> 1)) & 0x55555555u); }
__device__ __forceinline__ uint4 ld_evict_last_u32x4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}

// compact records: all node indices fit 32 bits (3 * hash_domain < 2^32)
__device__ __forceinline__ uint64_t mphf_eval_compact(const MphfDev &m, uint64_t a, uint64_t b, uint64_t c) {
    const uint32_t d = (uint32_t)m.hash_domain;
    const uint32_t n0 = fastmod32(a, d, m.magic);
    const uint32_t n1 = d + fastmod32(b, d, m.magic);
    const uint32_t n2 = 2u * d + fastmod32(c, d, m.magic);
    const uint32_t q0 = __umulhi(n0, 0xAAAAAAABu) >> 5, q1 = __umulhi(n1, 0xAAAAAAABu) >> 5,
                   q2 = __umulhi(n2, 0xAAAAAAABu) >> 5;  // n / 48
    const uint4 r0 = ld_evict_last_u32x4(&m.crecs[q0]);
    const uint4 r1 = ld_evict_last_u32x4(&m.crecs[q1]);
    const uint4 r2 = ld_evict_last_u32x4(&m.crecs[q2]);
    const uint32_t p0 = n0 - q0 * 48u, p1 = n1 - q1 * 48u, p2 = n2 - q2 * 48u;  // pair inside the record
    const uint32_t w0 = p0 < 16u ? r0.x : (p0 < 32u ? r0.y : r0.z);
    const uint32_t w1 = p1 < 16u ? r1.x : (p1 < 32u ? r1.y : r1.z);
    const uint32_t w2 = p2 < 16u ? r2.x : (p2 < 32u ? r2.y : r2.z);
    const uint32_t s0 = (p0 & 15u) * 2u, s1 = (p1 & 15u) * 2u, s2 = (p2 & 15u) * 2u;
    const uint32_t v = ((w0 >> s0) & 3u) + ((w1 >> s1) & 3u) + ((w2 >> s2) & 3u);
    const uint32_t hidx = (0x24924u >> (2u * v)) & 3u;  // v % 3 for v <= 9, two bits per entry
    const uint4 r = hidx == 0 ? r0 : (hidx == 1 ? r1 : r2);
    const uint32_t p = hidx == 0 ? p0 : (hidx == 1 ? p1 : p2);
    const uint32_t w = hidx == 0 ? w0 : (hidx == 1 ? w1 : w2);
    const uint32_t sh = hidx == 0 ? s0 : (hidx == 1 ? s1 : s2);
    uint32_t rank = r.w + nonzero_pairs32(w & ((1u << sh) - 1u));  // sh <= 30
    if (p >= 16u) rank += nonzero_pairs32(r.x);
    if (p >= 32u) rank += nonzero_pairs32(r.y);
    return (uint64_t)rank;
}

// fused records: the chosen node (global pair index) is returned through *node when asked for (index upload)
__device__ __forceinline__ uint64_t mphf_eval_fused(const MphfDev &m, uint64_t a, uint64_t b, uint64_t c, uint32_t *node = nullptr) {
    const uint32_t d = (uint32_t)m.hash_domain;
    const uint32_t n0 = fastmod32(a, d, m.magic);
    const uint32_t n1 = d + fastmod32(b, d, m.magic);
    const uint32_t n2 = 2u * d + fastmod32(c, d, m.magic);
    const uint4 r0 = ld_evict_last_u32x4(&m.frecs[n0 >> 4]);
    const uint4 r1 = ld_evict_last_u32x4(&m.frecs[n1 >> 4]);
    const uint4 r2 = ld_evict_last_u32x4(&m.frecs[n2 >> 4]);
    const uint32_t v = ((r0.x >> ((n0 & 15u) * 2u)) & 3u) + ((r1.x >> ((n1 & 15u) * 2u)) & 3u) + ((r2.x >> ((n2 & 15u) * 2u)) & 3u);
    const uint32_t hidx = (0x24924u >> (2u * v)) & 3u;  // v % 3 for v <= 9
    const uint4 r = hidx == 0 ? r0 : (hidx == 1 ? r1 : r2);
    const uint32_t nd = hidx == 0 ? n0 : (hidx == 1 ? n1 : n2);
    const uint32_t p = nd & 15u;
    const uint32_t rank = r.w + nonzero_pairs32(r.x & ((1u << (2u * p)) - 1u));  // 2p <= 30
    const uint32_t fp = ((p < 8u ? r.y : r.z) >> ((p & 7u) * 4u)) & 15u;
    if (node) *node = nd;
    return (uint64_t)rank | ((uint64_t)(0x10u | fp) << kFusedTagShift);
}

__device__ __forceinline__ uint64_t mphf_eval(const MphfDev &m, uint64_t a, uint64_t b, uint64_t c) {
    if (m.frecs != nullptr) return mphf_eval_fused(m, a, b, c);
    if (m.crecs != nullptr) return mphf_eval_compact(m, a, b, c);
    const uint64_t d = m.hash_domain;
    uint64_t n0 = fastmod(a, d, m.magic);
    uint64_t n1 = d + fastmod(b, d, m.magic);
    uint64_t n2 = 2 * d + fastmod(c, d, m.magic);
    ulonglong2 r0 = ld_evict_last_u64x2(&m.recs[n0 >> 5]);
    ulonglong2 r1 = ld_evict_last_u64x2(&m.recs[n1 >> 5]);
    ulonglong2 r2 = ld_evict_last_u64x2(&m.recs[n2 >> 5]);
    uint32_t s0 = (uint32_t)(n0 & 31) * 2, s1 = (uint32_t)(n1 & 31) * 2, s2 = (uint32_t)(n2 & 31) * 2;
    uint32_t v = (uint32_t)((r0.x >> s0) & 3) + (uint32_t)((r1.x >> s1) & 3) + (uint32_t)((r2.x >> s2) & 3);
    uint32_t hidx = v % 3;
    uint64_t word = hidx == 0 ? r0.x : (hidx == 1 ? r1.x : r2.x);
    uint64_t base = hidx == 0 ? r0.y : (hidx == 1 ? r1.y : r2.y);
    uint32_t sh = hidx == 0 ? s0 : (hidx == 1 ? s1 : s2);
    uint64_t mask = ((uint64_t)1 << sh) - 1;  // sh <= 62
    return base + nonzero_pairs64(word & mask);
}

// ---- 2-bit codec --------------------------------------------------------------------
// kmers.cpp:17-23: A0 C1 G2 T3, anything else (incl. lower case) 0
__device__ __forceinline__ uint32_t base_code_strict(uint32_t ch) {
    return ch == 'C' ? 1u : (ch == 'G' ? 2u : (ch == 'T' ? 3u : 0u));
}
__device__ __forceinline__ bool is_acgt_upper(uint32_t ch) {
    return ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T';
}
// ((c>>1)^(c>>2))&3 maps A,a->0 C,c->1 G,g->2 T,t->3 (only meaningful for ACGT letters)
__device__ __forceinline__ uint32_t base_code_fast(uint32_t ch) { return ((ch >> 1) ^ (ch >> 2)) & 3u; }

__device__ __forceinline__ uint64_t reverse_pairs64(uint64_t x) {
    x = __brevll(x);                                                          // bit reversal
    return ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);  // un-swap inside pairs
}
// reverseDNA (kmers.cpp:376-381): (~pair_reverse(x)) >> 18; the reference loop is a full
// 64-bit pair reversal for every input, so this closed form is identical for all x.
__device__ __forceinline__ uint64_t revcomp23(uint64_t x) { return (~reverse_pairs64(x)) >> 18; }
__device__ __forceinline__ uint32_t revcomp13(uint32_t x) {  // kmers.cpp:383-388
    uint32_t r = __brev(x);
    r = ((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1);
    return (~r) >> 6;
}

// expand 8 two-bit codes (code j at bits [2j+1:2j]) into 8 ASCII bytes, byte j = "ACGT"[code j]
__device__ __forceinline__ uint64_t ascii8_from_codes_le(uint32_t x16) {
    uint32_t t = (x16 | (x16 << 8)) & 0x00FF00FFu;
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    t = (t | (t << 2)) & 0x33333333u;  // nibble j = code j
    uint32_t lo = __byte_perm(0x54474341u, 0u, t);
    uint32_t hi = __byte_perm(0x54474341u, 0u, t >> 16);
    return ((uint64_t)hi << 32) | lo;
}

// ASCII string of the k-mer whose REVERSE COMPLEMENT is `rc_of_kmer`, as little-endian words.
// pair_reverse(u) == ~revcomp(u) on k pairs, i.e. the codes of u in string order sit in
// ~rc at bits [2j+1:2j]: no bit reversal needed when both strands are at hand.
__device__ __forceinline__ void ascii_words23_from_rc(uint64_t rc_of_kmer, uint64_t &w0, uint64_t &w1,
                                                      uint64_t &w2) {
    uint64_t x = ~rc_of_kmer;
    w0 = ascii8_from_codes_le((uint32_t)(x & 0xFFFF));
    w1 = ascii8_from_codes_le((uint32_t)((x >> 16) & 0xFFFF));
    w2 = ascii8_from_codes_le((uint32_t)((x >> 32) & 0x3FFF)) & 0x00FFFFFFFFFFFFFFULL;  // 7 bytes
}
__device__ __forceinline__ void ascii_words13_from_rc(uint32_t rc_of_kmer, uint64_t &w0, uint64_t &w1) {
    uint32_t x = ~rc_of_kmer;
    w0 = ascii8_from_codes_le(x & 0xFFFF);
    w1 = ascii8_from_codes_le((x >> 16) & 0x3FF) & 0x000000FFFFFFFFFFULL;  // 5 bytes
}

// ASCII words of the reverse complement of a VALID (upper-case ACGT) 23-byte string given as
// ASCII words: complement per byte (A^T = 0x15, C^G = 0x04, bit 1 tells which) and byte reversal by
// one PRMT per output word (byte j = in byte 22-j: the alignment is the same for every word).
__device__ __forceinline__ uint32_t ascii_complement4(uint32_t w) {
    return w ^ 0x15151515u ^ (((w >> 1) & 0x01010101u) * 0x11u);
}
__device__ __forceinline__ void rc_ascii_words23(uint64_t e0, uint64_t e1, uint64_t e2, uint64_t &f0, uint64_t &f1, uint64_t &f2) {
    const uint32_t c0 = ascii_complement4((uint32_t)e0), c1 = ascii_complement4((uint32_t)(e0 >> 32)),
                   c2 = ascii_complement4((uint32_t)e1), c3 = ascii_complement4((uint32_t)(e1 >> 32)),
                   c4 = ascii_complement4((uint32_t)e2), c5 = ascii_complement4((uint32_t)(e2 >> 32));
    const uint32_t o0 = __byte_perm(c4, c5, 0x3456), o1 = __byte_perm(c3, c4, 0x3456), o2 = __byte_perm(c2, c3, 0x3456),
                   o3 = __byte_perm(c1, c2, 0x3456), o4 = __byte_perm(c0, c1, 0x3456), o5 = __byte_perm(0u, c0, 0x3456);
    f0 = ((uint64_t)o1 << 32) | o0;
    f1 = ((uint64_t)o3 << 32) | o2;
    f2 = ((uint64_t)o5 << 32) | o4;
}

// MPHF id of a packed 23-mer u given its reverse complement r (hashes the ASCII string of u)
__device__ __forceinline__ uint64_t mphf_lookup23(const MphfDev &m, uint64_t rc_of_u) {
    uint64_t w0, w1, w2, a, b, c;
    ascii_words23_from_rc(rc_of_u, w0, w1, w2);
    jenkins_short(m.seed, w0, w1, w2, 23u, a, b, c);
    return mphf_eval(m, a, b, c);
}
__device__ __forceinline__ uint64_t mphf_lookup13(const MphfDev &m, uint32_t rc_of_u) {
    uint64_t w0, w1, a, b, c;
    ascii_words13_from_rc(rc_of_u, w0, w1);
    jenkins_short(m.seed, w0, w1, 0, 13u, a, b, c);
    return mphf_eval(m, a, b, c);
}

// checker/tf probe: returns true and *tf when slot h holds `kmer`.  h comes straight from mphf_eval: on the fused
// layout its top byte carries the fingerprint of the chosen node, which is checked and stripped here.
// own_string = the string that was hashed is the ASCII form of `kmer` (then a stored kmer was reached through the node
// the MPHF assigned to it, and the fused fingerprint of that node is its own).  The reference also compares slots
// reached by hashing OTHER bytes (raw strings with non-ACGT characters, python_wrapper.cpp:610-622): those probes
// pass own_string = false and go straight to the full compare, as does every probe of a layout with no filter.
__device__ __forceinline__ bool probe23(const Index23Dev &ix, uint64_t &h, uint64_t kmer, uint32_t &tf, bool own_string = true) {
    const uint32_t tag = (uint32_t)(h >> kFusedTagShift);
    h &= kFusedIdMask;
    if (h >= ix.n) return false;
    if (tag & 0x10u) {
        if (own_string && (tag & 15u) != fingerprint4(kmer)) return false;
    } else if (ix.fp != nullptr) {  // the separate tier is indexed by slot: valid whichever bytes led to the slot
        uint32_t f;
        if (ix.fp_bits == 8) {
            asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(f) : "l"(ix.fp + h), "l"(l2_policy_evict_last()));
            if (f != fingerprint8(kmer)) return false;
        } else {
            asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(f) : "l"(ix.fp + (h >> 1)), "l"(l2_policy_evict_last()));
            if (((f >> (((uint32_t)h & 1u) * 4u)) & 0xFu) != fingerprint4(kmer)) return false;
        }
    }
    uint4 r = ld_evict_first_u32x4(&ix.recs[h]);
    uint64_t chk = ((uint64_t)r.y << 32) | r.x;
    tf = r.z;
    return chk == kmer;
}

}  // namespace aix
