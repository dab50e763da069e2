// query23.cuh -- the per-query device code of the 23-mer lookup (shared by the batch
// query kernels, coverage and the positions index).
//
// Reference: AindexWrapper::get_tf_value_23mer python_wrapper.cpp:610-627 and the users of
// the same probe sequence (:700-742, :1219-1286); PHASH_MAP::get_pfid hash.hpp:150-170.
#pragma once
#include "../../include/aindex_cuda.h"
#include "device_common.cuh"

namespace aix {

// 2-bit value of a 23-byte string (words zero padded past byte 22), the value of its reverse complement, AND whether
// every byte is an upper-case ACGT letter.  Per 4-byte word (7 instructions, one of them on the FMA pipe):
//   x   = (w ^ w >> 1) & 0x06060606       letter code of each byte in its bits 2:1 (A 0, C 1, G 2, T 3)
//   x * (2^5 + 2^11 + 2^17 + 2^23)        top byte = c3 c2 c1 c0, the four codes in REVERSE order: the six top bytes,
//                                         little-endian, are the reversed code string, whose complement is the
//                                         reverse complement r; the forward value is revcomp23(r) (two BREVs)
//   PRMT("A.C.", "G.T.", nibbles 2c)      the one byte value that is valid for each code, compared with the word
// (prmt.b32 by inline PTX: __byte_perm masks the selector with 0x7777 first, one more ALU instruction per word)
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t expect_acgt4(uint32_t c4) {  // c4: letter code in bits 1:0 of each byte
    const uint32_t t = c4 | (c4 >> 4);
    return __byte_perm(0x54474341u, 0u, __byte_perm(t, 0u, 0x4420));
}
__device__ __forceinline__ void encode_validate23_rc(uint64_t r0, uint64_t r1, uint64_t r2, bool &all_acgt, uint64_t &u, uint64_t &r) {
    const uint32_t w[6] = {(uint32_t)r0, (uint32_t)(r0 >> 32), (uint32_t)r1, (uint32_t)(r1 >> 32), (uint32_t)r2, (uint32_t)(r2 >> 32)};
    uint32_t p[6], bad = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const uint32_t x = (w[j] ^ (w[j] >> 1)) & 0x06060606u;
        p[j] = x * 0x00820820u;
        const uint32_t sel = prmt_b32(x + (x >> 4), 0u, 0x4420u);            // nibble i = 2 * code of byte i
        const uint32_t diff = prmt_b32(0x00430041u, 0x00540047u, sel) ^ w[j];
        bad |= j == 5 ? (diff & 0x00FFFFFFu) : diff;  // byte 23 is padding
    }
    all_acgt = bad == 0;
    const uint32_t p01 = prmt_b32(p[0], p[1], 0x0073u), p23 = prmt_b32(p[2], p[3], 0x0073u), p45 = prmt_b32(p[4], p[5], 0x0073u);
    const uint32_t lo = prmt_b32(p01, p23, 0x5410u);
    const uint64_t y = ((uint64_t)(p45 & 0x3FFFu) << 32) | lo;  // reversed code string; the 24th byte is padding
    r = y ^ 0x3FFFFFFFFFFFULL;
    u = reverse_pairs64(y) >> 18;  // == revcomp23(r): the complements cancel
}
__device__ __forceinline__ uint64_t encode_validate23(uint64_t r0, uint64_t r1, uint64_t r2, bool &all_acgt) {
    uint64_t u, r;
    encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
    return u;
}

// strict encoder of get_dna23_bitset (kmers.cpp:12-25): non-ACGT (and missing) bytes -> 0
__device__ __forceinline__ uint64_t encode23_strict(uint64_t r0, uint64_t r1, uint64_t r2) {
    uint64_t u = 0;
#pragma unroll
    for (int j = 0; j < 23; ++j) {
        uint64_t w = j < 8 ? r0 : (j < 16 ? r1 : r2);
        uint32_t ch = (uint32_t)(w >> (8 * (j & 7))) & 0xFFu;
        u = (u << 2) | base_code_strict(ch);
    }
    return u;
}

struct Hit {
    int strand;   // 0 not found, 1 forward, 2 reverse (python_wrapper.cpp:726-742)
    uint64_t h;
    uint32_t tf;
};

// get_tf_value_23mer for a VALID packed 23-mer u (r = revcomp23(u)).
// asc_u* = ASCII words of u if have_asc_u (saves one expansion).
template <bool kCanon>
__device__ __forceinline__ Hit lookup_packed23(const Index23Dev &ix, const MphfDev &m, uint64_t u, uint64_t r,
                                               bool have_asc_u, uint64_t e0, uint64_t e1, uint64_t e2) {
    Hit hit = {0, 0, 0};
    uint64_t a, b, c;
    if (kCanon) {
        // every stored k-mer is canonical: only min(u, r) can be present, and the reference's
        // forward probe of a non-canonical u can never verify -> one probe, same answer.
        const bool fwd = u <= r;
        if (have_asc_u) {
            if (!fwd) rc_ascii_words23(e0, e1, e2, e0, e1, e2);
        } else {
            ascii_words23_from_rc(fwd ? r : u, e0, e1, e2);
        }
        jenkins_short(m.seed, e0, e1, e2, 23u, a, b, c);
        uint64_t h = mphf_eval(m, a, b, c);
        uint32_t tf;
        if (probe23(ix, h, fwd ? u : r, tf)) {
            hit.strand = fwd ? 1 : 2; hit.h = h; hit.tf = tf;
        }
        return hit;
    }
    if (!have_asc_u) ascii_words23_from_rc(r, e0, e1, e2);
    jenkins_short(m.seed, e0, e1, e2, 23u, a, b, c);
    uint64_t h1 = mphf_eval(m, a, b, c);
    uint32_t tf;
    if (probe23(ix, h1, u, tf)) {
        hit.strand = 1; hit.h = h1; hit.tf = tf;
        return hit;
    }
    ascii_words23_from_rc(u, e0, e1, e2);
    jenkins_short(m.seed, e0, e1, e2, 23u, a, b, c);
    uint64_t h2 = mphf_eval(m, a, b, c);
    if (probe23(ix, h2, r, tf)) {
        hit.strand = 2; hit.h = h2; hit.tf = tf;
    }
    return hit;
}

// lexicographic compare of two 23-byte strings held as little-endian words (a < b -> -1)
__device__ __forceinline__ int cmp_words23(uint64_t a0, uint64_t a1, uint64_t a2, uint64_t b0, uint64_t b1, uint64_t b2) {
    auto be = [](uint64_t x) {
        uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
        return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
    };
    uint64_t x, y;
    x = be(a0); y = be(b0); if (x != y) return x < y ? -1 : 1;
    x = be(a1); y = be(b1); if (x != y) return x < y ? -1 : 1;
    x = be(a2); y = be(b2); if (x != y) return x < y ? -1 : 1;
    return 0;
}

// one query, all modes.  r0..r2: first 23 raw bytes (zero padded past len); p: record bytes
template <int kMode, bool kCanon>
__device__ __forceinline__ void query23(const Index23Dev &ix, const MphfDev &m, uint64_t r0, uint64_t r1, uint64_t r2,
                                        uint32_t len, const uint8_t *p, uint64_t i, void *out) {
    // fast encode + validate (a valid query's ASCII words are the raw words themselves)
    bool all_acgt;
    uint64_t u, r;
    encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
    const uint64_t e0 = r0, e1 = r1, e2 = r2;
    const bool valid = (len == 23u) && all_acgt;

    Hit hit = {0, 0, 0};
    uint64_t ustrict = u, rstrict = r;
    if (valid) {
        if (kMode != AIX_Q_PFID) hit = lookup_packed23<kCanon>(ix, m, u, r, true, e0, e1, e2);
    } else {
        // exact reference sequence for odd queries: hash the RAW bytes (full length) forward,
        // the decoded reverse complement backward (python_wrapper.cpp:610-622)
        ustrict = encode23_strict(r0, r1, r2);
        rstrict = revcomp23(ustrict);
        if (kMode != AIX_Q_PFID && !((kMode == AIX_Q_TOTAL || kMode == AIX_Q_BOTH) && len != 23u)) {
            uint64_t a, b, c;
            if (len <= 23u) jenkins_short(m.seed, r0, r1, r2, len, a, b, c);
            else jenkins_bytes(m.seed, p, len, a, b, c);
            uint64_t h1 = mphf_eval(m, a, b, c);
            uint32_t tf;
            if (probe23(ix, h1, ustrict, tf, false)) {  // raw bytes were hashed, not the string of ustrict
                hit.strand = 1; hit.h = h1; hit.tf = tf;
            } else {
                uint64_t h2 = mphf_lookup23(m, ustrict);  // ASCII of rstrict
                if (probe23(ix, h2, rstrict, tf)) {
                    hit.strand = 2; hit.h = h2; hit.tf = tf;
                }
            }
        }
    }

    if (kMode == AIX_Q_TF) {
        ((uint32_t *)out)[i] = hit.tf;
    } else if (kMode == AIX_Q_KID) {
        ((uint64_t *)out)[i] = hit.strand ? hit.h : 0;  // python_wrapper.cpp:700-716
    } else if (kMode == AIX_Q_STRAND) {
        ((uint64_t *)out)[i] = (uint64_t)hit.strand;
    } else if (kMode == AIX_Q_TOTAL || kMode == AIX_Q_BOTH) {
        // python_wrapper.cpp:1230-1275: len != 23 -> 0; second value = tf of the decoded
        // reverse-complement STRING, itself a full forward-then-reverse lookup
        uint32_t fwd = 0, rev = 0;
        if (len == 23u) {
            fwd = hit.tf;
            if (kCanon && valid) {
                rev = fwd;  // both lookups probe the same canonical k-mer
            } else {
                Hit h2 = lookup_packed23<kCanon>(ix, m, rstrict, ustrict, false, 0, 0, 0);
                rev = h2.tf;
            }
        }
        if (kMode == AIX_Q_TOTAL) ((uint64_t *)out)[i] = (uint64_t)fwd + (uint64_t)rev;
        else {
            ((uint32_t *)out)[2 * i] = fwd;
            ((uint32_t *)out)[2 * i + 1] = rev;
        }
    } else if (kMode == AIX_Q_PFID) {
        // hash.hpp:150-170: the lexicographically smaller of the raw string and the decoded
        // reverse complement decides which single probe is made
        uint64_t v0, v1, v2;
        ascii_words23_from_rc(ustrict, v0, v1, v2);  // ASCII of rstrict
        int cmp;
        if (valid) cmp = ustrict <= rstrict ? -1 : 1;
        else {
            cmp = cmp_words23(r0, r1, r2, v0, v1, v2);
            if (cmp == 0 && len > 23u) cmp = 1;
        }
        uint64_t res = ix.n;
        uint32_t tf;
        if (cmp <= 0) {
            uint64_t a, b, c;
            if (len <= 23u) jenkins_short(m.seed, r0, r1, r2, len, a, b, c);
            else jenkins_bytes(m.seed, p, len, a, b, c);
            uint64_t h1 = mphf_eval(m, a, b, c);
            if (probe23(ix, h1, ustrict, tf, valid)) res = h1;  // a valid query's raw bytes are the string of ustrict
        } else {
            uint64_t a, b, c;
            jenkins_short(m.seed, v0, v1, v2, 23u, a, b, c);
            uint64_t h1 = mphf_eval(m, a, b, c);
            if (probe23(ix, h1, rstrict, tf)) res = h1;
        }
        ((uint64_t *)out)[i] = res;
    }
}


// get_tf_value_23mer of one raw window of exactly 23 bytes (r0..r2 little-endian words):
// the coverage / read-scan form of query23<AIX_Q_TF>.
template <bool kCanon>
__device__ __forceinline__ Hit find23_window(const Index23Dev &ix, const MphfDev &m, uint64_t r0, uint64_t r1,
                                             uint64_t r2) {
    bool all_acgt;
    uint64_t u, r;
    encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
    if (all_acgt) return lookup_packed23<kCanon>(ix, m, u, r, true, r0, r1, r2);
    // window with a non-ACGT byte: raw bytes forward, decoded reverse complement backward
    Hit hit = {0, 0, 0};
    uint64_t us = encode23_strict(r0, r1, r2), rs = revcomp23(us), a, b, c;
    jenkins_short(m.seed, r0, r1, r2, 23u, a, b, c);
    uint64_t h1 = mphf_eval(m, a, b, c);
    uint32_t tf;
    if (probe23(ix, h1, us, tf, false)) {  // raw bytes were hashed
        hit.strand = 1; hit.h = h1; hit.tf = tf;
    } else {
        uint64_t h2 = mphf_lookup23(m, us);
        if (probe23(ix, h2, rs, tf)) {
            hit.strand = 2; hit.h = h2; hit.tf = tf;
        }
    }
    return hit;
}

// 23 bytes at an arbitrary (unaligned) global address as three little-endian words
__device__ __forceinline__ void load_window23(const uint8_t *p, uint64_t &r0, uint64_t &r1, uint64_t &r2) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    uint32_t x0 = __ldg(w), x1 = __ldg(w + 1), x2 = __ldg(w + 2), x3 = __ldg(w + 3), x4 = __ldg(w + 4), x5 = __ldg(w + 5);
    // the 7th word holds window bytes only for byte offsets 2 and 3 (then it is in bounds)
    uint32_t x6 = sh >= 16u ? __ldg(w + 6) : 0u;
    uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
             y3 = __funnelshift_r(x3, x4, sh), y4 = __funnelshift_r(x4, x5, sh), y5 = __funnelshift_r(x5, x6, sh);
    r0 = ((uint64_t)y1 << 32) | y0;
    r1 = ((uint64_t)y3 << 32) | y2;
    r2 = (((uint64_t)y5 << 32) | y4) & 0x00FFFFFFFFFFFFFFULL;
}

}  // namespace aix
