// query23.cuh -- the per-query device code of the 23-mer lookup (shared by the batch
// query kernels, coverage and the positions index).
//
// Reference: AindexWrapper::get_tf_value_23mer python_wrapper.cpp:610-627 and the users of
// the same probe sequence (:700-742, :1219-1286); PHASH_MAP::get_pfid hash.hpp:150-170.
#pragma once
#include "../../include/aindex_cuda.h"
#include "device_common.cuh"

namespace aix {

// 2-bit value of a 23-byte string (words zero padded past byte 22) AND whether every byte is an
// upper-case ACGT letter: the letter code of each byte selects the one byte value that is valid
// for it ("ACGT"[code], by PRMT) and the word is compared with that expectation.
__device__ __forceinline__ uint32_t expect_acgt4(uint32_t c4) {  // c4: letter code in bits 1:0 of each byte
    const uint32_t t = c4 | (c4 >> 4);
    return __byte_perm(0x54474341u, 0u, __byte_perm(t, 0u, 0x4420));
}
__device__ __forceinline__ uint64_t encode_validate23(uint64_t r0, uint64_t r1, uint64_t r2, bool &all_acgt) {
    const uint32_t w[6] = {(uint32_t)r0, (uint32_t)(r0 >> 32), (uint32_t)r1, (uint32_t)(r1 >> 32), (uint32_t)r2, (uint32_t)(r2 >> 32)};
    uint32_t p[6], bad = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const uint32_t c4 = ((w[j] >> 1) ^ (w[j] >> 2)) & 0x03030303u;
        p[j] = (c4 * 0x40100401u) >> 24;  // c0<<6 | c1<<4 | c2<<2 | c3
        const uint32_t diff = expect_acgt4(c4) ^ w[j];
        bad |= j == 5 ? (diff & 0x00FFFFFFu) : diff;  // byte 23 is padding
    }
    all_acgt = bad == 0;
    const uint64_t all48 = ((uint64_t)((p[0] << 8) | p[1]) << 32) | ((uint64_t)((p[2] << 8) | p[3]) << 16) | ((p[4] << 8) | p[5]);
    return all48 >> 2;  // 23 codes; the 24th byte is padding
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
    uint64_t u = encode_validate23(r0, r1, r2, all_acgt);
    uint64_t r = revcomp23(u);
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
    uint64_t u = encode_validate23(r0, r1, r2, all_acgt);
    uint64_t r = revcomp23(u);
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
