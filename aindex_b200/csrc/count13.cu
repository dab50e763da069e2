// count13.cu -- 13-mer counting: rolling 2-bit windows over raw read bytes and a
// direct-address 4^13 histogram in HBM.
//
// Reference: Kmer13Counter (src/count_kmers13.cpp): process_sequence :131-161,
// normalize_sequence :113-126, is_valid_kmer :101-108, readers :211-272,
// save_counts :358-388.  The reference evaluates the MPHF once per k-mer occurrence
// (:148) and bumps counts[mphf(kmer)]; for a fixed .pf that is a fixed permutation of the
// 2-bit value v of the window, so the GPU counts in direct-address space hist[v] (no hash
// in the inner loop) and applies perm13 once at the end (aix_count13_finish).
//
// K1/K2 kernel shape: one thread per 16 input bytes (one aligned 16-byte load + the
// previous 16 bytes from the neighbouring lane by shuffle), 29 two-bit codes and two
// 29-bit masks (valid letter / newline) per thread, window predicates by shift-and
// doubling, then up to 16 RED.ADD.U32 into the 256 MiB histogram.
#include "aix_internal.cuh"

namespace aix {

constexpr int kCntBlock = 256;
constexpr uint32_t kMask26 = (1u << 26) - 1;
// statistics buffer (u64): [3] out-of-range ids, [4] FASTQ line counter, [kStatBase + 3*slot + j] =
// sequences / windows / valid of slot `slot` (summed on the host in aix_count13_stats)
constexpr uint32_t kStatBase = 8, kStatSlots = 64, kStatWords = kStatBase + 3 * kStatSlots;
constexpr uint32_t kStatFlushed = 5;  // non-zero once part of the counts has been moved to the u64 histogram

// ---- SIMD classification of 16 input bytes -------------------------------------------------
// codes: 16 two-bit codes, byte 0 in bits 31:30 ... byte 15 in bits 1:0 (A0 C1 G2 T3 via
//        ((c>>1)^(c>>2))&3, meaningful where the byte is valid)
// masks: bit i      = byte i is an ACGT letter of either case (std::toupper, count_kmers13.cpp:118)
//        bit 16 + i = byte i is '\n'
struct Sum16 {
    uint32_t codes;
    uint32_t masks;
};

// bit 7 of every non-zero byte of x
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) {
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// bits 7/15/23/31 -> bits 0/1/2/3
__device__ __forceinline__ uint32_t gather_msb4(uint32_t x) { return ((x >> 7) * 0x01020408u) >> 24; }

template <bool kNeedNl>
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &codes8, uint32_t &valid4, uint32_t &nl4) {
    uint32_t c4 = ((w >> 1) ^ (w >> 2)) & 0x03030303u;  // letter code of every byte
    codes8 = (c4 * 0x40100401u) >> 24;                   // c0<<6 | c1<<4 | c2<<2 | c3
    uint32_t t = c4 | (c4 >> 4);
    uint32_t sel = __byte_perm(t, 0u, 0x4420);           // nibble j = code of byte j
    uint32_t expect = __byte_perm(0x54474341u, 0u, sel); // "ACGT"[code]: the only byte that is valid
    valid4 = gather_msb4(nonzero_bytes(expect ^ (w & 0xDFDFDFDFu))) ^ 0xFu;
    nl4 = kNeedNl ? (gather_msb4(nonzero_bytes(w ^ 0x0A0A0A0Au)) ^ 0xFu) : 0u;
}

template <bool kNeedNl>
__device__ __forceinline__ Sum16 classify16(uint4 v) {
    uint32_t c0, c1, c2, c3, v0, v1, v2, v3, n0, n1, n2, n3;
    classify4<kNeedNl>(v.x, c0, v0, n0);
    classify4<kNeedNl>(v.y, c1, v1, n1);
    classify4<kNeedNl>(v.z, c2, v2, n2);
    classify4<kNeedNl>(v.w, c3, v3, n3);
    Sum16 s;
    s.codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    s.masks = v0 | (v1 << 4) | (v2 << 8) | (v3 << 12) | (n0 << 16) | (n1 << 20) | (n2 << 24) | (n3 << 28);
    return s;
}

__device__ __forceinline__ uint32_t run13(uint32_t v) {  // bit i = v[i..i+12] all set
    uint32_t a1 = v & (v >> 1);
    uint32_t a2 = a1 & (a1 >> 2);
    uint32_t a3 = a2 & (a2 >> 4);
    return a3 & (a2 >> 8) & (v >> 12);
}

__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// base: 16-byte aligned; owned bytes are [own_begin, own_end) (own_begin multiple of 16).
// Bytes before own_begin are lookback if own_begin > 0, otherwise the shard start (virtual
// newline); bytes at or past own_end read as newline.
//
// Each thread classifies its own 16 bytes once (SIMD over 32-bit words) and receives the
// summary of the previous 16 bytes from the neighbouring lane (shuffle), the neighbouring
// warp (shared memory) or, for the first thread of a CTA, by classifying them itself.
// Positions 0..12 = lookback bytes 3..15, positions 13..28 = own bytes; the window starting at
// position s (1..16) ends on an own byte and is emitted by this thread.
//
// smask/sval: multi-pass mode -- only k-mers whose byte offset satisfies (off & smask) == sval
// are counted in this launch, so the histogram slice being updated stays L2 resident
// (smask = 0: single pass).  kStats: accumulate the count_kmers13 statistics (first pass only).
// kVariant: 0 one RED per window + warp merge of low-complexity windows (default), 1 thread-local run-length
// merge, 2 warp match_any merge of every window, 3 one RED per window, nothing merged.
template <bool kStats, int kVariant>
__global__ void __launch_bounds__(kCntBlock) count13_kernel(const uint8_t *__restrict__ base, uint64_t own_begin,
                                                          uint64_t own_end, uint32_t *__restrict__ hist,
                                                          unsigned long long *__restrict__ stats, uint32_t smask,
                                                          uint32_t sval) {
    __shared__ uint2 edge[kCntBlock / 32];
    __shared__ uint32_t lc_cnt[16], lc_off[16];  // low-complexity windows of the CTA (variant 0)
    if (kVariant == 0 && threadIdx.x < 16) lc_cnt[threadIdx.x] = 0;  // visible after the barrier below
    const uint64_t t = (uint64_t)blockIdx.x * kCntBlock + threadIdx.x;
    const uint64_t pos = own_begin + t * 16;
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    Sum16 own;
    own.codes = 0;
    own.masks = 0xFFFF0000u;  // inactive: 16 newlines
    if (pos < own_end) {
        own = classify16<kStats>(__ldcs(reinterpret_cast<const uint4 *>(base + pos)));
        if (pos + 16 > own_end) {  // tail: bytes past the end are newlines
            uint32_t keep = (1u << (uint32_t)(own_end - pos)) - 1u;
            own.masks = (own.masks & (keep | (keep << 16))) | ((~keep & 0xFFFFu) << 16);
        }
    }
    Sum16 prev;
    prev.codes = __shfl_up_sync(0xFFFFFFFFu, own.codes, 1);
    prev.masks = __shfl_up_sync(0xFFFFFFFFu, own.masks, 1);
    if (lane == 31) edge[wid] = make_uint2(own.codes, own.masks);
    __syncthreads();
    if (lane == 0) {
        if (wid > 0) {
            uint2 e = edge[wid - 1];
            prev.codes = e.x;
            prev.masks = e.y;
        } else if (pos >= 16 && pos - 16 < own_end) {  // CTA boundary: the 16 bytes before are whole
            prev = classify16<kStats>(__ldg(reinterpret_cast<const uint4 *>(base + pos - 16)));
        } else {
            prev.codes = 0;
            prev.masks = 0xFFFF0000u;  // shard start = virtual newline
        }
    }
    const uint32_t valid29 = ((prev.masks & 0xFFFFu) >> 3) | ((own.masks & 0xFFFFu) << 13);
    const uint32_t wvalid = run13(valid29) & 0x1FFFEu;
    // byte offsets into the histogram: (prev:own codes) << 2, window s at bits [2(16-s)+2 ...]
    const uint32_t lo2 = own.codes << 2, hi2 = __funnelshift_l(own.codes, prev.codes, 2);
    constexpr uint32_t kOffMask = kMask26 << 2;
    // Low-complexity windows (period 1 or 2: poly-A, (CA)n ... -- 16 k-mers that real genomes hit millions of
    // times) would serialise on a handful of L2 counters.  They are found per thread with a few 64-bit operations
    // on the 29 codes (code[q] == code[q-2] for 11 consecutive q) and counted in 16 shared-memory counters per CTA
    // that are flushed with at most 16 REDs; warps of uniform-random reads never enter that path.
    uint32_t lowc = 0;  // bit (32 - 2s): window s has period <= 2
    if (kVariant == 0) {
        const uint64_t w64 = ((uint64_t)(prev.codes & kMask26) << 32) | own.codes;  // position p at bits [57-2p, 56-2p]
        const uint64_t d = w64 ^ (w64 >> 4);
        const uint64_t z = ~(d | (d >> 1)) & 0x0015555555555555ULL;  // field q (>= 2): code[q] == code[q-2]
        const uint64_t r1 = z & (z >> 2), r2 = r1 & (r1 >> 4), r3 = r2 & (r2 >> 8);
        lowc = (uint32_t)(r3 & (r1 >> 16) & (z >> 20));  // field q: the 11 comparisons q-10 .. q all hold
    }
    uint32_t pend_off = 0, pend_c = 0;
    // warp-uniform choice: warps without a low-complexity window (all of them on random reads) run the plain loop
    if (kVariant == 0 && __any_sync(0xFFFFFFFFu, lowc != 0)) {
#pragma unroll
        for (int s = 1; s <= 16; ++s) {
            const uint32_t off = __funnelshift_r(lo2, hi2, 2 * (16 - s)) & kOffMask;
            const bool ok = ((wvalid >> s) & 1u) && ((off & smask) == sval);
            const bool sp = (lowc >> (32 - 2 * s)) & 1u;
            if (ok && !sp) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + off), 1u);
            if (ok && sp) {
                // a period-<=2 13-mer is determined by its last two bases: 16 CTA-level counters in shared memory
                const uint32_t cls = (off >> 2) & 15u;
                atomicAdd(&lc_cnt[cls], 1u);
                lc_off[cls] = off;  // every writer stores the same value
            }
        }
    } else {
#pragma unroll
        for (int s = 1; s <= 16; ++s) {
            const uint32_t off = __funnelshift_r(lo2, hi2, 2 * (16 - s)) & kOffMask;
            const bool ok = ((wvalid >> s) & 1u) && ((off & smask) == sval);
            if (kVariant == 0 || kVariant == 3) {
                if (ok) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + off), 1u);
            } else if (kVariant == 1) {
                if (ok && pend_c && off == pend_off) { ++pend_c; continue; }
                if (pend_c) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + pend_off), pend_c);
                pend_c = ok ? 1u : 0u;
                pend_off = off;
            } else {
                const uint32_t key = ok ? off : 0xFFFFFFFFu;
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
                if (ok && lane == (unsigned)(__ffs(peers) - 1))
                    atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + off), (uint32_t)__popc(peers));
            }
        }
    }
    if (kVariant == 1 && pend_c) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + pend_off), pend_c);
    if (kVariant == 0) {
        __syncthreads();
        if (threadIdx.x < 16) {
            const uint32_t c = lc_cnt[threadIdx.x];
            if (c) atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + lc_off[threadIdx.x]), c);
        }
    }
    if (kStats) {
        // count_kmers13 statistics: per thread <= 16 of each kind, so one packed 32-bit REDUX per
        // warp, one shared-memory word per warp and three global atomics per CTA on one of
        // kStatSlots slot triples (every warp hitting the same three addresses serialises in L2
        // and made this pass 4.6x slower than the others: profiles/r01_count13_ncu.txt)
        __shared__ uint32_t wstat[kCntBlock / 32];
        const uint32_t nl29 = (prev.masks >> 19) | ((own.masks >> 16) << 13);
        const uint32_t wline = run13(~nl29 & 0x1FFFFFFFu) & 0x1FFFEu;
        const uint32_t n_win = __popc(wline), n_valid = __popc(wvalid);
        const uint32_t n_seq = __popc((nl29 << 1) & wline);  // first window of a line with >= 13 characters
        const uint32_t packed = __reduce_add_sync(0xFFFFFFFFu, (n_seq << 20) | (n_win << 10) | n_valid);
        if (lane == 0) wstat[wid] = packed;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t s_seq = 0, s_win = 0, s_valid = 0;
#pragma unroll
            for (int w = 0; w < kCntBlock / 32; ++w) {
                const uint32_t x = wstat[w];
                s_seq += x >> 20; s_win += (x >> 10) & 0x3FFu; s_valid += x & 0x3FFu;
            }
            if (s_win) {
                unsigned long long *slot = stats + kStatBase + 3u * (blockIdx.x & (kStatSlots - 1u));
                atomicAdd(slot + 0, (unsigned long long)s_seq);
                atomicAdd(slot + 1, (unsigned long long)s_win);
                atomicAdd(slot + 2, (unsigned long long)s_valid);
            }
        }
    }
}

// hist64[v] += hist32[v]; hist32[v] = 0
__global__ void flush_hist_kernel(uint32_t *__restrict__ h32, uint64_t *__restrict__ h64) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (uint32_t)AIX_TOTAL_13MERS) return;
    uint32_t c = h32[v];
    if (c) {
        h64[v] += c;
        h32[v] = 0;
    }
}

// tf_out[mphf(v)] += hist64[v] for v in [v_begin, v_end); ids >= 4^13 are the reference's
// "hash index out of range" branch (count_kmers13.cpp:153-156): counted invalid, not stored
__global__ void permute13_kernel(MphfDev m, const uint64_t *__restrict__ hist64, uint32_t v_begin, uint32_t v_end,
                                 unsigned long long *__restrict__ tf_out, unsigned long long *__restrict__ out_of_range) {
    uint32_t v = v_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= v_end) return;
    uint64_t c = hist64[v];
    if (!c) return;
    uint64_t id = mphf_lookup13(m, revcomp13(v));
    if (id < AIX_TOTAL_13MERS) atomicAdd(tf_out + id, (unsigned long long)c);
    else atomicAdd(out_of_range, (unsigned long long)c);
}

// ---- FASTQ: blank every byte that is not on a sequence line (line_no % 4 == 1,
// count_kmers13.cpp:240-257) so the buffer can be counted as plain text ----------------
constexpr int kTileBytes = kCntBlock * 16;

__device__ __forceinline__ uint32_t count_nl16(uint4 v) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w}, n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t x = w[i] ^ 0x0A0A0A0Au;  // zero byte <=> newline
        uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
        n += __popc(z);
    }
    return n;
}

__device__ __forceinline__ uint4 load_own16(const uint8_t *base, uint64_t pos, uint64_t own_end) {
    uint4 own = make_uint4(0, 0, 0, 0);
    if (pos < own_end) {
        own = *reinterpret_cast<const uint4 *>(base + pos);
        if (pos + 16 > own_end) {
            uint32_t w[4] = {own.x, own.y, own.z, own.w};
            uint32_t keep = (uint32_t)(own_end - pos);
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if ((uint32_t)b >= keep) w[b >> 2] &= ~(0xFFu << (8 * (b & 3)));
            own = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    return own;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem /*>=33*/, uint32_t &total) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) smem[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if (lane >= (unsigned)o) s += y;
        }
        smem[lane] = s;  // inclusive per-warp totals
    }
    __syncthreads();
    total = smem[(blockDim.x >> 5) - 1];
    uint32_t warp_off = wid ? smem[wid - 1] : 0;
    __syncthreads();
    return warp_off + x - v;
}

__global__ void __launch_bounds__(kCntBlock) nl_count_kernel(const uint8_t *__restrict__ base, uint64_t own_begin,
                                                           uint64_t own_end, uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t sm[33];
    const uint64_t pos = own_begin + ((uint64_t)blockIdx.x * kCntBlock + threadIdx.x) * 16;
    uint32_t n = count_nl16(load_own16(base, pos, own_end));
    uint32_t total;
    block_exclusive_scan(n, sm, total);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// single block: tile_base[i] = line_base + exclusive scan; line_base += total
__global__ void nl_scan_kernel(uint32_t *__restrict__ tile_cnt, uint64_t *__restrict__ tile_base, uint32_t n_tiles,
                               unsigned long long *__restrict__ line_base) {
    __shared__ uint32_t sm[33];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = *line_base;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n_tiles; i0 += blockDim.x) {
        uint32_t i = i0 + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_cnt[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, sm, total);
        if (i < n_tiles) tile_base[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *line_base = carry;
}

// in place: bytes whose line number % 4 != 1 become '\n'.  The (up to 16) lookback bytes
// in front of own_begin are handled by thread 0 of block 0, walking backwards.
__global__ void __launch_bounds__(kCntBlock) fastq_mask_kernel(uint8_t *__restrict__ base, uint64_t own_begin, uint64_t own_end,
                                                             const uint64_t *__restrict__ tile_base) {
    __shared__ uint32_t sm[33];
    const uint64_t pos = own_begin + ((uint64_t)blockIdx.x * kCntBlock + threadIdx.x) * 16;
    uint4 own = load_own16(base, pos, own_end);
    uint32_t n = count_nl16(own), total;
    uint32_t ex = block_exclusive_scan(n, sm, total);
    uint64_t line = tile_base[blockIdx.x] + ex;
    if (pos < own_end) {
        uint32_t w[4] = {own.x, own.y, own.z, own.w};
        uint32_t lim = (uint32_t)(own_end - pos < 16 ? own_end - pos : 16);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            uint32_t c = (w[b >> 2] >> (8 * (b & 3))) & 0xFFu;
            bool is_nl = (c == '\n') && (uint32_t)b < lim;
            if ((line & 3u) != 1u) w[b >> 2] = (w[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | (0x0Au << (8 * (b & 3)));
            if (is_nl) ++line;
        }
        if (lim == 16) *reinterpret_cast<uint4 *>(base + pos) = make_uint4(w[0], w[1], w[2], w[3]);
        else
            for (uint32_t b = 0; b < lim; ++b) base[pos + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && own_begin > 0) {
        uint64_t ln = tile_base[0];  // line number of the byte at own_begin
        for (uint64_t p = own_begin; p-- > 0;) {
            if (base[p] == '\n') --ln;  // byte p is the newline that ends line ln-1... (see below)
            // line number of byte p = newlines strictly before p = ln after the decrement
            if (base[p] != '\n' && (ln & 3u) != 1u) base[p] = '\n';
        }
    }
}


// ---- histogram combine over NVLink peer memory (one rank per GPU) ---------------------------------------
// Rank r owns the k-mer range [v_begin, v_end).  Instead of widening its whole 256 MiB u32 histogram to u64
// and handing 512 MiB to an NCCL reduce-scatter, every rank reads its range straight out of the u32 histograms
// of all ranks (peer loads through NVLink / NVSwitch, 16 bytes per lane), adds them up in u64 and writes its
// slice: widen + reduce-scatter in one kernel, a quarter of the bytes on the wire.  A peer that already moved
// counts to its u64 histogram (flag word) contributes that slice too.
constexpr int kMaxPeers = 16;
struct PeerTable {
    const uint32_t *h32[kMaxPeers];
    const unsigned long long *h64[kMaxPeers];
    const unsigned long long *stats[kMaxPeers];
    int n;
};

__global__ void __launch_bounds__(256) count13_reduce_peers_kernel(PeerTable t, uint32_t v_begin, uint32_t count,
                                                                 unsigned long long *__restrict__ out) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (i >= count) return;  // count is a multiple of 4 (4^13 / world)
    unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int p = 0; p < t.n; ++p) {
        const uint4 x = *reinterpret_cast<const uint4 *>(t.h32[p] + v_begin + i);
        a0 += x.x; a1 += x.y; a2 += x.z; a3 += x.w;
        if (t.stats[p][kStatFlushed]) {
            const ulonglong2 y0 = *reinterpret_cast<const ulonglong2 *>(t.h64[p] + v_begin + i);
            const ulonglong2 y1 = *reinterpret_cast<const ulonglong2 *>(t.h64[p] + v_begin + i + 2);
            a0 += y0.x; a1 += y0.y; a2 += y1.x; a3 += y1.y;
        }
    }
    *reinterpret_cast<ulonglong2 *>(out + i) = make_ulonglong2(a0, a1);
    *reinterpret_cast<ulonglong2 *>(out + i + 2) = make_ulonglong2(a2, a3);
}

// ---- FASTA: concatenate the lines of a record, drop header lines (count_kmers13.cpp:211-235) ----------
// Output byte stream (counted as plain text afterwards): sequence-line bytes without their newlines, and
// one '\n' for every header line (it ends the record before it; empty records are empty lines, which the
// plain reader skips).  A byte is a line start iff the byte before it is '\n' (or it is the first byte);
// a line is a header iff its first byte is '>'.  Three kernels: per-tile summary (header state leaving
// the tile, kept-byte count for either entering state), one single-block scan over tiles, compaction.
__device__ __forceinline__ uint32_t last_nonzero_scan_warp(uint32_t v) {  // inclusive: last non-zero value at or before the lane
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= (unsigned)o && v == 0) v = y;
    }
    return v;
}

// per-thread walk over its 16 bytes.  state_in: 0 sequence line, 1 header line.  Returns the state after
// the last byte; *ls = 2 + header flag of the last line start in the chunk (0 if none); kept bytes are
// appended to out (when not null) and counted.
__device__ __forceinline__ uint32_t fasta_walk16(uint4 own, uint32_t n_bytes, uint32_t prev, uint32_t state_in, uint32_t *ls,
                                                 uint32_t *count, uint8_t *out) {
    const uint32_t w[4] = {own.x, own.y, own.z, own.w};
    uint32_t state = state_in, last = 0, cnt = 0;
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        if ((uint32_t)b < n_bytes) {
            const uint32_t c = (w[b >> 2] >> (8 * (b & 3))) & 0xFFu;
            const bool line_start = prev == '\n';
            if (line_start) {
                state = c == '>' ? 1u : 0u;
                last = 2u + state;
            }
            if (state) {
                if (line_start) {
                    if (out) out[cnt] = '\n';
                    ++cnt;
                }
            } else if (c != '\n') {
                if (out) out[cnt] = (uint8_t)c;
                ++cnt;
            }
            prev = c;
        }
    }
    *ls = last;
    *count = cnt;
    return state;
}

// state entering every thread relative to the tile: 0 = inherits the tile's entering state, else 2 + header flag
__device__ __forceinline__ uint32_t fasta_block_prefix_state(uint32_t my_last, uint32_t *smem /* >= 32 */, uint32_t *tile_last) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = last_nonzero_scan_warp(my_last);
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = lane < nw ? smem[lane] : 0u;
        t = last_nonzero_scan_warp(t);
        smem[lane] = t;
    }
    __syncthreads();
    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) excl = 0;
    const uint32_t warp_in = wid ? smem[wid - 1] : 0u;
    *tile_last = smem[nw - 1];
    __syncthreads();
    return excl ? excl : warp_in;
}

__device__ __forceinline__ uint32_t fasta_prev_byte(const uint8_t *base, uint64_t pos) { return pos ? base[pos - 1] : (uint32_t)'\n'; }

struct FastaTile {
    uint32_t last;   // 0: no line start in the tile, else 2 + header flag of its last line start
    uint32_t cnt[2];  // kept bytes if the tile is entered in state 0 / 1
};

__global__ void __launch_bounds__(kCntBlock) fasta_summary_kernel(const uint8_t *__restrict__ base, uint64_t len,
                                                                FastaTile *__restrict__ tiles) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t sums[2];
    if (threadIdx.x < 2) sums[threadIdx.x] = 0;
    const uint64_t pos = ((uint64_t)blockIdx.x * kCntBlock + threadIdx.x) * 16;
    const uint4 own = load_own16(base, pos, len);
    const uint32_t nb = pos < len ? (uint32_t)(len - pos < 16 ? len - pos : 16) : 0u;
    const uint32_t prev = pos < len ? fasta_prev_byte(base, pos) : 0u;
    uint32_t ls0, ls1, c0, c1;
    fasta_walk16(own, nb, prev, 0u, &ls0, &c0, nullptr);
    fasta_walk16(own, nb, prev, 1u, &ls1, &c1, nullptr);  // ls1 == ls0: line starts do not depend on the state
    uint32_t tile_last;
    const uint32_t in = fasta_block_prefix_state(ls0, sm, &tile_last);
    // entered through an earlier line start of this tile: the count is fixed; otherwise it depends on the tile's state
    const uint32_t k0 = in ? ((in & 1u) ? c1 : c0) : c0, k1 = in ? ((in & 1u) ? c1 : c0) : c1;
    const uint32_t w0 = __reduce_add_sync(0xFFFFFFFFu, k0), w1 = __reduce_add_sync(0xFFFFFFFFu, k1);
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(&sums[0], w0);
        atomicAdd(&sums[1], w1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        FastaTile t;
        t.last = tile_last; t.cnt[0] = sums[0]; t.cnt[1] = sums[1];
        tiles[blockIdx.x] = t;
    }
}

// single block: entering state and output offset of every tile; total[0] = kept bytes
__global__ void fasta_scan_kernel(const FastaTile *__restrict__ tiles, uint32_t n_tiles, uint32_t *__restrict__ tile_state,
                                  unsigned long long *__restrict__ tile_off, unsigned long long *__restrict__ total) {
    __shared__ uint32_t sm[33];
    __shared__ unsigned long long carry_off;
    __shared__ uint32_t carry_state;
    if (threadIdx.x == 0) { carry_off = 0; carry_state = 0; }  // the file starts on a sequence-or-header line start
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ unsigned long long wsum[33];
    for (uint32_t i0 = 0; i0 < n_tiles; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        FastaTile t = {0u, {0u, 0u}};
        if (i < n_tiles) t = tiles[i];
        uint32_t tile_last;
        uint32_t in = fasta_block_prefix_state(t.last, sm, &tile_last);
        const uint32_t st = in ? (in & 1u) : carry_state;
        unsigned long long c = i < n_tiles ? t.cnt[st] : 0ull, x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long v = lane < nw ? wsum[lane] : 0ull;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= (unsigned)o) v += y;
            }
            wsum[lane] = v;
        }
        __syncthreads();
        const unsigned long long off = carry_off + (wid ? wsum[wid - 1] : 0ull) + x - c;
        if (i < n_tiles) {
            tile_state[i] = st;
            tile_off[i] = off;
        }
        const unsigned long long batch_total = wsum[nw - 1];
        __syncthreads();
        if (threadIdx.x == 0) {
            carry_off += batch_total;
            if (tile_last) carry_state = tile_last & 1u;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry_off;
}

__global__ void __launch_bounds__(kCntBlock) fasta_compact_kernel(const uint8_t *__restrict__ base, uint64_t len,
                                                                const uint32_t *__restrict__ tile_state,
                                                                const unsigned long long *__restrict__ tile_off,
                                                                uint8_t *__restrict__ out) {
    __shared__ uint32_t sm[33];
    const uint64_t pos = ((uint64_t)blockIdx.x * kCntBlock + threadIdx.x) * 16;
    const uint4 own = load_own16(base, pos, len);
    const uint32_t nb = pos < len ? (uint32_t)(len - pos < 16 ? len - pos : 16) : 0u;
    const uint32_t prev = pos < len ? fasta_prev_byte(base, pos) : 0u;
    uint32_t ls, cnt;
    fasta_walk16(own, nb, prev, 0u, &ls, &cnt, nullptr);
    uint32_t tile_last;
    const uint32_t in = fasta_block_prefix_state(ls, sm, &tile_last);
    const uint32_t st = in ? (in & 1u) : tile_state[blockIdx.x];
    uint8_t local[16];
    fasta_walk16(own, nb, prev, st, &ls, &cnt, local);
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(cnt, sm, total);
    uint8_t *dst = out + tile_off[blockIdx.x] + ex;
    for (uint32_t b = 0; b < cnt; ++b) dst[b] = local[b];
}

}  // namespace aix

using namespace aix;

// streaming chunk size for host/device staged input (multiple of kTileBytes)
static const uint64_t kChunkBytes = 256ull << 20;

static int c13_alloc(aix_ctx *ctx) {
    if (!ctx->c13_hist32) {
        cudaError_t e = cudaMalloc(&ctx->c13_hist32, AIX_TOTAL_13MERS * 4);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->c13_hist64, AIX_TOTAL_13MERS * 8);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->c13_stats_dev, kStatWords * 8);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(AIX_ERR_NOMEM, "count13 buffers: %s", cudaGetErrorString(e));
        }
    }
    return AIX_OK;
}

static int c13_flush_on(aix_ctx *ctx, cudaStream_t st) {
    flush_hist_kernel<<<aix_grid(AIX_TOTAL_13MERS, 256), 256, 0, st>>>(ctx->c13_hist32, ctx->c13_hist64);
    AIX_LAUNCH_CHECK(ctx);
    AIX_CUDA(ctx, cudaMemsetAsync(ctx->c13_stats_dev + kStatFlushed, 1, 8, st));
    ctx->c13_pending_windows = 0;
    return AIX_OK;
}

// 0 one RED per window, low-complexity windows merged across the warp (default), 1 thread-local run-length merge,
// 2 warp match_any merge of every window, 3 nothing merged (profiles/r01_count13_sweep.txt)
static int count_variant() {
    const char *e = getenv("AIX_COUNT13_VARIANT");
    return e ? atoi(e) : 0;
}

static int launch_count(aix_ctx *ctx, cudaStream_t st, const uint8_t *base, uint64_t own_begin, uint64_t own_end) {
    if (own_end <= own_begin) return AIX_OK;
    uint64_t threads = (own_end - own_begin + 15) / 16;
    unsigned grid = aix_grid(threads, kCntBlock);
    unsigned long long *stats = (unsigned long long *)ctx->c13_stats_dev;
    // 4 passes over 64 MiB histogram slices: random REDs run at 210 G/s while the slice is L2
    // resident, against 32 G/s on the whole 256 MiB table (profiles/atomic_roofline.txt)
    int passes_log2 = 2;
    if (const char *e = getenv("AIX_COUNT13_PASSES_LOG2")) passes_log2 = atoi(e);
    if (passes_log2 < 0) passes_log2 = 0;
    if (passes_log2 > 6) passes_log2 = 6;
    const int variant = count_variant();
    const uint32_t smask = passes_log2 ? (((1u << passes_log2) - 1u) << (28 - passes_log2)) : 0u;
    for (uint32_t p = 0; p < (1u << passes_log2); ++p) {
        const uint32_t sval = passes_log2 ? (p << (28 - passes_log2)) : 0u;
#define AIX_C13_LAUNCH(STATS, VAR) \
    count13_kernel<STATS, VAR><<<grid, kCntBlock, 0, st>>>(base, own_begin, own_end, ctx->c13_hist32, stats, smask, sval)
        if (p == 0) {
            if (variant == 1) AIX_C13_LAUNCH(true, 1);
            else if (variant == 2) AIX_C13_LAUNCH(true, 2);
            else if (variant == 3) AIX_C13_LAUNCH(true, 3);
            else AIX_C13_LAUNCH(true, 0);
        } else {
            if (variant == 1) AIX_C13_LAUNCH(false, 1);
            else if (variant == 2) AIX_C13_LAUNCH(false, 2);
            else if (variant == 3) AIX_C13_LAUNCH(false, 3);
            else AIX_C13_LAUNCH(false, 0);
        }
#undef AIX_C13_LAUNCH
        AIX_LAUNCH_CHECK(ctx);
    }
    return AIX_OK;
}

// FASTA image in HBM (16-byte aligned) -> plain-text image in *out_dev (caller frees), *out_len bytes
static int fasta_compact_dev(aix_ctx *ctx, cudaStream_t st, const uint8_t *in_dev, uint64_t len, uint8_t **out_dev, uint64_t *out_len) {
    *out_dev = nullptr;
    *out_len = 0;
    const uint32_t n_tiles = (uint32_t)((len + kTileBytes - 1) / kTileBytes);
    FastaTile *tiles = nullptr;
    uint32_t *tstate = nullptr;
    unsigned long long *toff = nullptr, *total = nullptr;
    uint8_t *out = nullptr;
    auto cleanup = [&]() { cudaFree(tiles); cudaFree(tstate); cudaFree(toff); cudaFree(total); };
    cudaError_t e = cudaMalloc(&tiles, (size_t)n_tiles * sizeof(FastaTile));
    if (e == cudaSuccess) e = cudaMalloc(&tstate, (size_t)n_tiles * 4);
    if (e == cudaSuccess) e = cudaMalloc(&toff, (size_t)n_tiles * 8);
    if (e == cudaSuccess) e = cudaMalloc(&total, 8);
    if (e == cudaSuccess) e = cudaMalloc(&out, len + 64);  // every input byte yields at most one output byte
    if (e != cudaSuccess) {
        cudaGetLastError(); cleanup(); cudaFree(out);
        return ctx->fail(AIX_ERR_NOMEM, "FASTA staging (%llu bytes): %s", (unsigned long long)len, cudaGetErrorString(e));
    }
    fasta_summary_kernel<<<n_tiles, kCntBlock, 0, st>>>(in_dev, len, tiles);
    fasta_scan_kernel<<<1, 1024, 0, st>>>(tiles, n_tiles, tstate, toff, total);
    fasta_compact_kernel<<<n_tiles, kCntBlock, 0, st>>>(in_dev, len, tstate, toff, out);
    ctx->launches += 3;
    unsigned long long h_total = 0;
    e = cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaMemsetAsync(out + h_total, '\n', 64, st);  // the last record ends here
    cleanup();
    if (e != cudaSuccess) {
        cudaGetLastError(); cudaFree(out);
        return ctx->fail(AIX_ERR_CUDA, "FASTA compaction: %s", cudaGetErrorString(e));
    }
    *out_dev = out;
    *out_len = h_total + 1;
    return AIX_OK;
}

static int detect_format_host(const uint8_t *b, uint64_t len) {  // count_kmers13.cpp:194-206
    if (len == 0 || b[0] == '\n') return AIX_FMT_PLAIN;
    if (b[0] == '>') return AIX_FMT_FASTA;
    if (b[0] == '@') return AIX_FMT_FASTQ;
    return AIX_FMT_PLAIN;
}

// shared by aix_count13_add (host source) and aix_count13_add_dev (device source)
static int c13_add_impl(aix_ctx *ctx, const uint8_t *src, uint64_t len, int fmt, bool src_is_device) {
    if (!ctx) return AIX_ERR_ARG;
    if (!ctx->c13_active) return ctx->fail(AIX_ERR_STATE, "aix_count13_begin not called");
    if (len == 0) return AIX_OK;
    if (!src) return ctx->fail(AIX_ERR_ARG, "null input");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (fmt == AIX_FMT_DETECT) {
        uint8_t first = 0;
        if (src_is_device) AIX_CUDA(ctx, cudaMemcpy(&first, src, 1, cudaMemcpyDeviceToHost));
        else first = src[0];
        fmt = detect_format_host(&first, 1);
    }
    uint8_t *fasta_in = nullptr, *fasta_out = nullptr;  // device staging of a FASTA image (freed on return)
    struct Free2 {
        uint8_t *&a, *&b;
        ~Free2() { cudaFree(a); cudaFree(b); }
    } free2{fasta_in, fasta_out};
    if (fmt == AIX_FMT_FASTA) {
        // records are concatenated on the device; the plain-text image is then counted in place
        const uint8_t *in_dev = src;
        // in place only when every 16-byte vector of the image lies inside the caller's buffer
        if (!src_is_device || ((uintptr_t)src & 15) != 0 || (len & 15) != 0) {
            cudaError_t e = cudaMalloc(&fasta_in, len + 64);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return ctx->fail(AIX_ERR_NOMEM, "FASTA image (%llu bytes): %s", (unsigned long long)len, cudaGetErrorString(e));
            }
            AIX_CUDA(ctx, cudaMemcpyAsync(fasta_in, src, len, cudaMemcpyDefault, ctx->stream));
            in_dev = fasta_in;
        }
        uint64_t plain_len = 0;
        AIX_TRY(fasta_compact_dev(ctx, ctx->stream, in_dev, len, &fasta_out, &plain_len));
        src = fasta_out;
        len = plain_len;
        src_is_device = true;
        fmt = AIX_FMT_PLAIN;
    }
    if (fmt != AIX_FMT_PLAIN && fmt != AIX_FMT_FASTQ) return ctx->fail(AIX_ERR_ARG, "unknown format %d", fmt);

    // every input byte starts at most one window; flush the u32 histogram before it can wrap
    auto account = [&](uint64_t nbytes, cudaStream_t st) -> int {
        if (ctx->c13_pending_windows + nbytes >= 0xFFFF0000ull) {
            // the flush is not atomic with respect to running count kernels: quiesce first
            AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[0]));
            AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[1]));
            AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            AIX_TRY(c13_flush_on(ctx, st));
            AIX_CUDA(ctx, cudaStreamSynchronize(st));
        }
        ctx->c13_pending_windows += nbytes;
        return AIX_OK;
    };

    // resident plain text: count in place, no copy.  Only whole 16-byte vectors are read in place; a ragged
    // tail (len % 16 bytes) goes through the staged path below with its 16 bytes of lookback, so nothing past
    // the caller's buffer is ever touched.
    uint64_t done = 0;
    if (src_is_device && fmt == AIX_FMT_PLAIN && ((uintptr_t)src & 15) == 0) {
        const uint64_t whole = len & ~15ull;
        const uint64_t step = 1ull << 31;  // keep each launch below the flush threshold
        while (done < whole) {
            uint64_t n = whole - done < step ? whole - done : step;
            AIX_TRY(account(n, ctx->stream));
            AIX_TRY(launch_count(ctx, ctx->stream, src, done, done + n));
            done += n;
        }
        if (done == len) {
            if (fasta_out) AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the staging buffers are released on return
            return AIX_OK;
        }
    }

    // staged path: chunk c holds [16 lookback bytes][chunk bytes] (the first chunk has no lookback)
    uint64_t chunk = kChunkBytes;
    if (const char *e = getenv("AIX_COUNT13_CHUNK")) {  // test hook: force the multi-chunk path on small inputs
        uint64_t v = strtoull(e, nullptr, 10);
        if (v >= 64) chunk = v & ~15ull;
    }
    if (chunk > len - done) chunk = (len - done + 15) & ~15ull;
    const uint32_t max_tiles = (uint32_t)((chunk + kTileBytes - 1) / kTileBytes) + 1;
    void *buf[2] = {nullptr, nullptr}, *tcnt[2] = {nullptr, nullptr}, *tbase[2] = {nullptr, nullptr};
    for (int s = 0; s < 2; ++s) {
        AIX_TRY(ctx->reserve(SCR_IN0 + s, chunk + 64, &buf[s]));
        if (fmt == AIX_FMT_FASTQ) {
            AIX_TRY(ctx->reserve(SCR_LEN0 + s, (size_t)max_tiles * 4, &tcnt[s]));
            AIX_TRY(ctx->reserve(SCR_OUT0 + s, (size_t)max_tiles * 8, &tbase[s]));
        }
        if (len - done <= chunk) break;
    }
    unsigned long long *line_base = (unsigned long long *)(ctx->c13_stats_dev + 4);
    if (fmt == AIX_FMT_FASTQ) AIX_CUDA(ctx, cudaMemsetAsync(line_base, 0, 8, ctx->stream));
    AIX_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    AIX_CUDA(ctx, cudaStreamWaitEvent(ctx->xfer[0], ctx->ev[0], 0));
    AIX_CUDA(ctx, cudaStreamWaitEvent(ctx->xfer[1], ctx->ev[0], 0));
    int c = 0;
    while (done < len) {
        uint64_t n = len - done < chunk ? len - done : chunk;
        int s = c & 1;
        cudaStream_t st = ctx->xfer[s];
        uint64_t lead = done ? 16 : 0;
        uint8_t *b = (uint8_t *)buf[s];
        if (fmt == AIX_FMT_FASTQ && c > 0) {
            // the line counter is carried on the device: chunks must be processed in order
            AIX_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[1 + ((c - 1) & 1)], 0));
        }
        AIX_CUDA(ctx, cudaMemcpyAsync(b, src + done - lead, n + lead, cudaMemcpyDefault, st));
        AIX_TRY(account(n, st));
        if (fmt == AIX_FMT_FASTQ) {
            uint32_t tiles = (uint32_t)((n + kTileBytes - 1) / kTileBytes);
            nl_count_kernel<<<tiles, kCntBlock, 0, st>>>(b, lead, lead + n, (uint32_t *)tcnt[s]);
            AIX_LAUNCH_CHECK(ctx);
            nl_scan_kernel<<<1, 1024, 0, st>>>((uint32_t *)tcnt[s], (uint64_t *)tbase[s], tiles, line_base);
            AIX_LAUNCH_CHECK(ctx);
            AIX_CUDA(ctx, cudaEventRecord(ctx->ev[1 + (c & 1)], st));
            fastq_mask_kernel<<<tiles, kCntBlock, 0, st>>>(b, lead, lead + n, (const uint64_t *)tbase[s]);
            AIX_LAUNCH_CHECK(ctx);
        }
        AIX_TRY(launch_count(ctx, st, b, lead, lead + n));
        done += n;
        ++c;
    }
    // later work on ctx->stream (flush / finish) must see the counts
    AIX_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->xfer[0]));
    AIX_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->xfer[1]));
    AIX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[3], 0));
    AIX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[4], 0));
    // host buffers (and the local FASTA copy) may be released by the caller on return
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[0]));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[1]));
    return AIX_OK;
}

extern "C" {

int aix_count13_begin(aix_ctx *ctx) {
    if (!ctx) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_TRY(c13_alloc(ctx));
    ctx->l2_unpin();
    AIX_CUDA(ctx, cudaMemsetAsync(ctx->c13_hist32, 0, AIX_TOTAL_13MERS * 4, ctx->stream));
    AIX_CUDA(ctx, cudaMemsetAsync(ctx->c13_hist64, 0, AIX_TOTAL_13MERS * 8, ctx->stream));
    AIX_CUDA(ctx, cudaMemsetAsync(ctx->c13_stats_dev, 0, kStatWords * 8, ctx->stream));
    ctx->c13_pending_windows = 0;
    ctx->c13_active = true;
    return AIX_OK;
}

int aix_count13_add(aix_ctx *ctx, const uint8_t *bytes, uint64_t len, int fmt) {
    return c13_add_impl(ctx, bytes, len, fmt, false);
}

int aix_count13_add_dev(aix_ctx *ctx, const uint8_t *bytes_dev, uint64_t len, int fmt) {
    return c13_add_impl(ctx, bytes_dev, len, fmt, true);
}

int aix_count13_flush(aix_ctx *ctx) {
    if (!ctx || !ctx->c13_active) return ctx ? ctx->fail(AIX_ERR_STATE, "count13 not active") : AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return c13_flush_on(ctx, ctx->stream);
}

uint64_t *aix_count13_hist_dev(aix_ctx *ctx) { return ctx ? ctx->c13_hist64 : nullptr; }

int aix_count13_stats(aix_ctx *ctx, aix_count_stats *stats) {
    if (!ctx || !stats || !ctx->c13_active) return AIX_ERR_ARG;
    uint64_t raw[kStatWords], h[3] = {0, 0, 0};
    AIX_CUDA(ctx, cudaMemcpyAsync(raw, ctx->c13_stats_dev, sizeof raw, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t s = 0; s < kStatSlots; ++s)
        for (int j = 0; j < 3; ++j) h[j] += raw[kStatBase + 3 * s + j];
    stats->sequences = h[0];
    stats->windows = h[1];
    stats->valid = h[2];
    stats->invalid = h[1] - h[2];
    return AIX_OK;
}

int aix_count13_finish_dev(aix_ctx *ctx, const aix_mphf *m, uint64_t v_begin, uint64_t v_end, uint64_t *tf_out_dev) {
    if (!ctx || !m || !tf_out_dev) return AIX_ERR_ARG;
    if (!ctx->c13_active) return ctx->fail(AIX_ERR_STATE, "count13 not active");
    if (v_begin > v_end || v_end > AIX_TOTAL_13MERS) return ctx->fail(AIX_ERR_ARG, "bad k-mer range");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (v_end == v_begin) return AIX_OK;
    permute13_kernel<<<aix_grid(v_end - v_begin, 256), 256, 0, ctx->stream>>>(
        m->dev(), ctx->c13_hist64, (uint32_t)v_begin, (uint32_t)v_end, (unsigned long long *)tf_out_dev,
        (unsigned long long *)(ctx->c13_stats_dev + 3));
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_count13_finish(aix_ctx *ctx, const aix_mphf *m, uint64_t v_begin, uint64_t v_end, uint64_t *tf_out,
                       aix_count_stats *stats) {
    if (!ctx || !m || !tf_out) return AIX_ERR_ARG;
    AIX_TRY(aix_count13_flush(ctx));
    void *tf_dev;
    AIX_TRY(ctx->reserve(SCR_TMP0, AIX_TOTAL_13MERS * 8, &tf_dev));
    AIX_CUDA(ctx, cudaMemsetAsync(tf_dev, 0, AIX_TOTAL_13MERS * 8, ctx->stream));
    AIX_TRY(aix_count13_finish_dev(ctx, m, v_begin, v_end, (uint64_t *)tf_dev));
    const bool whole = (v_begin == 0 && v_end == AIX_TOTAL_13MERS);
    if (whole) {
        AIX_CUDA(ctx, cudaMemcpyAsync(tf_out, tf_dev, AIX_TOTAL_13MERS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        // a slice lands on arbitrary ids: accumulate into the caller's array on the host
        std::vector<uint64_t> tmp(AIX_TOTAL_13MERS);
        AIX_CUDA(ctx, cudaMemcpyAsync(tmp.data(), tf_dev, AIX_TOTAL_13MERS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = 0; i < AIX_TOTAL_13MERS; ++i) tf_out[i] += tmp[i];
    }
    if (stats) {
        AIX_TRY(aix_count13_stats(ctx, stats));
        uint64_t oor = 0;
        AIX_CUDA(ctx, cudaMemcpy(&oor, ctx->c13_stats_dev + 3, 8, cudaMemcpyDeviceToHost));
        stats->valid -= oor;  // count_kmers13.cpp:153-156
        stats->invalid += oor;
    }
    return AIX_OK;
}


// ---- peer-memory combine: handles are exchanged by the caller (any transport), 3 x 64 bytes per rank --------
int aix_count13_ipc_export(aix_ctx *ctx, void *handles_out) {
    if (!ctx || !handles_out) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_TRY(c13_alloc(ctx));
    cudaIpcMemHandle_t h[3];
    AIX_CUDA(ctx, cudaIpcGetMemHandle(&h[0], ctx->c13_hist32));
    AIX_CUDA(ctx, cudaIpcGetMemHandle(&h[1], ctx->c13_hist64));
    AIX_CUDA(ctx, cudaIpcGetMemHandle(&h[2], ctx->c13_stats_dev));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles_out, h, sizeof h);
    return AIX_OK;
}

int aix_count13_peers_close(aix_ctx *ctx) {
    if (!ctx) return AIX_ERR_ARG;
    cudaSetDevice(ctx->device);
    for (int p = 0; p < ctx->c13_n_peers && ctx->c13_peer_ipc; ++p) {
        if (p == ctx->c13_my_rank) continue;
        for (int k = 0; k < 3; ++k)
            if (ctx->c13_peer[p][k]) cudaIpcCloseMemHandle(ctx->c13_peer[p][k]);
    }
    memset(ctx->c13_peer, 0, sizeof ctx->c13_peer);
    ctx->c13_n_peers = 0;
    ctx->c13_peer_ipc = true;
    return AIX_OK;
}

int aix_count13_peers_open(aix_ctx *ctx, const void *handles, int n_ranks, int my_rank) {
    if (!ctx || !handles || n_ranks < 1 || n_ranks > kMaxPeers || my_rank < 0 || my_rank >= n_ranks) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_TRY(c13_alloc(ctx));
    aix_count13_peers_close(ctx);
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)handles;
    ctx->c13_n_peers = n_ranks;
    ctx->c13_my_rank = my_rank;
    for (int p = 0; p < n_ranks; ++p) {
        if (p == my_rank) {
            ctx->c13_peer[p][0] = ctx->c13_hist32; ctx->c13_peer[p][1] = ctx->c13_hist64; ctx->c13_peer[p][2] = ctx->c13_stats_dev;
            continue;
        }
        for (int k = 0; k < 3; ++k) {
            cudaError_t e = cudaIpcOpenMemHandle(&ctx->c13_peer[p][k], h[3 * p + k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                ctx->c13_peer[p][k] = nullptr;
                aix_count13_peers_close(ctx);
                return ctx->fail(AIX_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
            }
        }
    }
    return AIX_OK;
}

int aix_count13_reduce_peers_dev(aix_ctx *ctx, uint64_t v_begin, uint64_t v_end, uint64_t *out_dev) {
    if (!ctx || !out_dev) return AIX_ERR_ARG;
    if (!ctx->c13_active) return ctx->fail(AIX_ERR_STATE, "count13 not active");
    if (ctx->c13_n_peers < 1) return ctx->fail(AIX_ERR_STATE, "aix_count13_peers_open not called");
    if (v_begin > v_end || v_end > AIX_TOTAL_13MERS || ((v_end - v_begin) & 3) || (v_begin & 3)) return ctx->fail(AIX_ERR_ARG, "bad k-mer range");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (v_end == v_begin) return AIX_OK;
    PeerTable t;
    t.n = ctx->c13_n_peers;
    for (int p = 0; p < t.n; ++p) {
        t.h32[p] = (const uint32_t *)ctx->c13_peer[p][0];
        t.h64[p] = (const unsigned long long *)ctx->c13_peer[p][1];
        t.stats[p] = (const unsigned long long *)ctx->c13_peer[p][2];
    }
    const uint32_t count = (uint32_t)(v_end - v_begin);
    count13_reduce_peers_kernel<<<aix_grid(count / 4, 256), 256, 0, ctx->stream>>>(t, (uint32_t)v_begin, count, (unsigned long long *)out_dev);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_count13_end(aix_ctx *ctx) {
    if (!ctx) return AIX_ERR_ARG;
    ctx->c13_active = false;
    return AIX_OK;
}

int aix_count13(aix_ctx *ctx, const aix_mphf *m, const uint8_t *bytes, uint64_t len, int fmt, uint64_t *tf_out,
                aix_count_stats *stats) {
    if (!ctx || !m || !tf_out) return AIX_ERR_ARG;
    AIX_TRY(aix_count13_begin(ctx));
    int rc = aix_count13_add(ctx, bytes, len, fmt);
    if (rc == AIX_OK) rc = aix_count13_finish(ctx, m, 0, AIX_TOTAL_13MERS, tf_out, stats);
    aix_count13_end(ctx);
    return rc;
}

}  // extern "C"
