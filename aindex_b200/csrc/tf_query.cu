// tf_query.cu -- batched term-frequency queries: 23-mer (MPHF + checker verify + tf
// gather) and 13-mer (direct-address gather), plus the index uploads.
//
// Reference: AindexWrapper::get_tf_value_23mer python_wrapper.cpp:610-627 (and its users
// :653-664, :700-742, :1219-1286), PHASH_MAP::get_freq/get_pfid hash.hpp:123-170,
// 13-mer queries python_wrapper.cpp:482-608, :938-980, loaders hash.cpp:367-450,
// python_wrapper.cpp:404-437.
//
// Kernel shape (K3): one query per thread.  Fixed 23-byte records take tf23_stream_kernel (a TMA ring
// of 32-query tiles per warp, see below); tf23_fixed_kernel is its predecessor (256 queries per CTA
// staged with 16-byte streaming loads behind a CTA barrier) and serves the last q mod 32 queries.
// Either way the bytes of a query are re-read from shared memory as 7 aligned words + funnel shifts,
// and a query is ~400 integer instructions, three independent 16-byte L2 loads (MPHF record), one
// L2 byte (fingerprint) and one 16-byte HBM load ({checker, tf} record) when the fingerprint matches.
// Latency is hidden by occupancy, not by intra-thread pipelining: the loads of a query depend on its hash.
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "aix_internal.cuh"
#include "batch_pipeline.cuh"
#include "query23.cuh"
#include "tf23_ring.cuh"
#include "tf23_filter.cuh"

namespace aix {

constexpr int kQBlock = 256;

// K3, fixed 23-byte records (the batch path of get_tf_values)
template <int kMode, bool kCanon>
__global__ void __launch_bounds__(kQBlock) tf23_fixed_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                           uint64_t q, void *__restrict__ out) {
    __shared__ __align__(16) uint32_t tile[(kQBlock * 23 + 16) / 4 + 4];
    const uint64_t q0 = (uint64_t)blockIdx.x * kQBlock;
    const uint64_t byte0 = q0 * 23;  // multiple of 16 (256*23 = 16*368)
    const uint64_t total = q * 23;
    constexpr int kVec = kQBlock * 23 / 16;  // 368
    const uint4 *src = reinterpret_cast<const uint4 *>(recs + byte0);
    uint4 *dst = reinterpret_cast<uint4 *>(tile);
    for (int v = threadIdx.x; v < kVec + 1; v += kQBlock) {
        uint4 x = make_uint4(0, 0, 0, 0);
        const uint64_t off = byte0 + (uint64_t)v * 16;
        if (off + 16 <= total) {
            x = __ldcs(src + v);
        } else if (off < total) {  // last, partial vector of the batch: never read past the buffer
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint32_t b = 0; off + b < total; ++b) w[b >> 2] |= (uint32_t)recs[off + b] << (8 * (b & 3));
            x = make_uint4(w[0], w[1], w[2], w[3]);
        }
        dst[v] = x;
    }
    __syncthreads();
    const uint64_t i = q0 + threadIdx.x;
    if (i >= q) return;
    const uint32_t base = threadIdx.x * 23u;
    const uint32_t w = base >> 2, sh = (base & 3u) * 8u;
    uint32_t x0 = tile[w], x1 = tile[w + 1], x2 = tile[w + 2], x3 = tile[w + 3], x4 = tile[w + 4], x5 = tile[w + 5],
             x6 = tile[w + 6];
    uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
             y3 = __funnelshift_r(x3, x4, sh), y4 = __funnelshift_r(x4, x5, sh), y5 = __funnelshift_r(x5, x6, sh);
    uint64_t r0 = ((uint64_t)y1 << 32) | y0, r1 = ((uint64_t)y3 << 32) | y2,
             r2 = (((uint64_t)y5 << 32) | y4) & 0x00FFFFFFFFFFFFFFULL;
    query23<kMode, kCanon>(ix, m, r0, r1, r2, 23u, recs + i * 23, i, out);
}


template <int kMode, bool kCanon, int kMinBlocks>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_stream_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                              uint64_t n_tiles, void *__restrict__ out) {
    __shared__ __align__(128) uint8_t slots[kStWarps][kStStages][kStSlot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WarpRing<kStTileBytes, kStSlot> ring;
    ring.init(&slots[wid][0][0], &bars[wid][0], lane);
    uint64_t tile0;
    const uint32_t my_tiles = warp_tiles<kStTilesPerWarp>(n_tiles, wid, tile0);
    if (my_tiles == 0) return;
    ring.start(recs + tile0 * kStTileBytes, my_tiles, lane);
    uint64_t i = tile0 * 32u + lane;  // query index of this lane in the current tile
    for (uint32_t it = 0; it < my_tiles; ++it) {
        const uint32_t tile = ring.acquire(it, lane);
        uint32_t x[7];
        lds_words7(tile, lane * 23u, x);
        __syncwarp();
        uint64_t r0, r1, r2;
        words23(x, lane * 23u, r0, r1, r2);
        query23<kMode, kCanon>(ix, m, r0, r1, r2, 23u, recs + i * 23, i, out);
        i += (uint64_t)kStWarps * 32u;
        ring.advance();
    }
}

// get_freq for 23-mers held as 6-byte dna_bitset records (dna_bitseq.hpp:22-61: 4 bases per byte, first base in
// bits 7:6; 23 bases = 46 bits + 2 zero bits): the smallest form a query batch can cross PCIe in (6 B instead of 23).
// The 46-bit value is get_dna23_bitset of the string, so this is PHASH_MAP::get_freq(uint64_t) (hash.hpp:123-140).
template <bool kCanon>
__global__ void __launch_bounds__(256) get_freq23_packed6_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ packed,
                                                               uint64_t q, uint32_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    const uint16_t *p = reinterpret_cast<const uint16_t *>(packed + i * 6);  // 6 i is even: 2-byte aligned
    const uint32_t h0 = __ldcs(p), h1 = __ldcs(p + 1), h2 = __ldcs(p + 2);
    // big-endian 48-bit number >> 2
    const uint64_t be = ((uint64_t)__byte_perm(h0, 0, 0x3301) << 32) | ((uint64_t)__byte_perm(h1, 0, 0x3301) << 16) |
                        (uint64_t)__byte_perm(h2, 0, 0x3301);
    const uint64_t u = be >> 2, r = revcomp23(u);
    uint32_t res = 0, tf;
    if (kCanon) {
        const bool fwd = u <= r;
        uint64_t h = mphf_lookup23(m, fwd ? r : u);  // hashes the ASCII string of min(u, r)
        if (probe23(ix, h, fwd ? u : r, tf)) res = tf;
    } else {
        uint64_t ha = mphf_lookup23(m, r);
        if (probe23(ix, ha, u, tf)) res = tf;
        else {
            uint64_t hb = mphf_lookup23(m, u);
            if (probe23(ix, hb, r, tf)) res = tf;
        }
    }
    __stcs(out + i, res);
}

// K3, generic records (any stride, per-record lengths)
template <int kMode, bool kCanon>
__global__ void __launch_bounds__(kQBlock) tf23_generic_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                             uint32_t stride, const uint8_t *__restrict__ lens, uint64_t q,
                                                             void *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * kQBlock + threadIdx.x;
    if (i >= q) return;
    uint32_t len = lens ? lens[i] : stride;
    if (len > stride) len = stride;
    const uint8_t *p = recs + i * stride;
    uint64_t w[3] = {0, 0, 0};
    const uint32_t nb = len < 23u ? len : 23u;
    for (uint32_t j = 0; j < nb; ++j) w[j >> 3] |= (uint64_t)__ldg(p + j) << (8 * (j & 7));
    query23<kMode, kCanon>(ix, m, w[0], w[1], w[2], len, p, i, out);
}

// PHASH_MAP::get_freq(uint64_t) (hash.hpp:123-140)
template <bool kCanon>
__global__ void get_freq23_kernel(Index23Dev ix, MphfDev m, const uint64_t *__restrict__ ukmers, uint64_t q,
                                  uint32_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    // get_bitset_dna23 decodes the low 46 bits only, so bits above do not reach the hash, but
    // the checker comparison uses the full 64-bit value
    uint64_t k = ukmers[i];
    uint64_t lo = k & ((1ULL << 46) - 1);
    uint64_t rl = revcomp23(lo);
    uint32_t res = 0, tf;
    if (kCanon) {
        // every stored value is canonical and below 2^46: the forward probe can only verify k == lo <= rl, the
        // reverse probe only rl < lo -- one probe of min(lo, rl), same answer
        const bool fwd = lo <= rl;
        if (!fwd || k == lo) {
            uint64_t h = mphf_lookup23(m, fwd ? rl : lo);  // hashes the ASCII string of min(lo, rl)
            if (probe23(ix, h, fwd ? lo : rl, tf)) res = tf;
        }
        out[i] = res;
        return;
    }
    uint64_t h1 = mphf_lookup23(m, rl);
    if (probe23(ix, h1, k, tf)) res = tf;
    else {
        uint64_t rk = revcomp23(k);       // reverseDNA: bits above 46 of k fall off, rk == rl
        uint64_t h2 = mphf_lookup23(m, lo);  // hashes the ASCII string of rk
        if (probe23(ix, h2, rk, tf)) res = tf;
    }
    out[i] = res;
}

// ---- index split by hash-id range over several GPUs (north star: "split by hash range when it exceeds HBM") ------
// The MPHF is small and replicated; the {checker, tf} records are cut into contiguous id ranges, one per GPU.
// A rank turns its queries into probes {id, packed k-mer that must be stored at id}, the caller routes every probe
// to the rank that owns the id (one all-to-all), the owner verifies it against its records and answers {hit, tf}.
// get_tf_value_23mer needs at most two probes per query (forward, then reverse: python_wrapper.cpp:610-622); both
// are produced eagerly and combined afterwards (first hit wins), which gives exactly the sequential answer.
constexpr unsigned long long kNoProbe = ~0ull;

template <bool kCanon>
__global__ void __launch_bounds__(kQBlock) tf23_probes_kernel(MphfDev m, uint64_t n_total, const uint8_t *__restrict__ recs,
                                                            uint32_t stride, const uint8_t *__restrict__ lens, uint64_t q,
                                                            ulonglong2 *__restrict__ probes /* [2q] */) {
    const uint64_t i = (uint64_t)blockIdx.x * kQBlock + threadIdx.x;
    if (i >= q) return;
    uint32_t len = lens ? lens[i] : stride;
    if (len > stride) len = stride;
    const uint8_t *p = recs + i * stride;
    uint64_t w[3] = {0, 0, 0};
    const uint32_t nb = len < 23u ? len : 23u;
    for (uint32_t j = 0; j < nb; ++j) w[j >> 3] |= (uint64_t)__ldg(p + j) << (8 * (j & 7));
    bool all_acgt;
    uint64_t u, r;
    encode_validate23_rc(w[0], w[1], w[2], all_acgt, u, r);
    uint64_t h1 = kNoProbe, k1 = 0, h2 = kNoProbe, k2 = 0, a, b, c, f0, f1, f2;
    if (len == 23u && all_acgt) {
        if (kCanon) {
            const bool fwd = u <= r;
            f0 = w[0]; f1 = w[1]; f2 = w[2];
            if (!fwd) rc_ascii_words23(w[0], w[1], w[2], f0, f1, f2);
            jenkins_short(m.seed, f0, f1, f2, 23u, a, b, c);
            h1 = mphf_eval(m, a, b, c);
            k1 = fwd ? u : r;
        } else {
            jenkins_short(m.seed, w[0], w[1], w[2], 23u, a, b, c);
            h1 = mphf_eval(m, a, b, c);
            k1 = u;
            rc_ascii_words23(w[0], w[1], w[2], f0, f1, f2);
            jenkins_short(m.seed, f0, f1, f2, 23u, a, b, c);
            h2 = mphf_eval(m, a, b, c);
            k2 = r;
        }
    } else {
        const uint64_t us = encode23_strict(w[0], w[1], w[2]), rs = revcomp23(us);
        if (len <= 23u) jenkins_short(m.seed, w[0], w[1], w[2], len, a, b, c);
        else jenkins_bytes(m.seed, p, len, a, b, c);
        h1 = mphf_eval(m, a, b, c);
        k1 = us;
        h2 = mphf_lookup23(m, us);  // hashes the ASCII string of rs
        k2 = rs;
    }
    if (h1 >= n_total) h1 = kNoProbe;
    if (h2 >= n_total) h2 = kNoProbe;
    probes[2 * i] = make_ulonglong2(h1, k1);
    probes[2 * i + 1] = make_ulonglong2(h2, k2);
}

// owner side: probes {local id, k-mer} -> (hit << 32) | tf
__global__ void probe23_kernel(const uint4 *__restrict__ recs, uint64_t n_local, const ulonglong2 *__restrict__ probes,
                               uint64_t cnt, unsigned long long *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const ulonglong2 pr = probes[i];
    unsigned long long res = 0;
    if (pr.x < n_local) {
        const uint4 rec = ld_evict_first_u32x4(&recs[pr.x]);
        if ((((uint64_t)rec.y << 32) | rec.x) == pr.y) res = (1ull << 32) | rec.z;
    }
    out[i] = res;
}

// probes -> per-owner buckets (a counting sort with `world` <= 16 buckets): send[slot] = {id - lo[owner], k-mer},
// tag[slot] = index of the probe, bucket o occupying [offs[o], offs[o] + counts[o]).  Order inside a bucket is arbitrary.
constexpr int kMaxRanks = 16;
struct OwnerBounds {
    unsigned long long b[kMaxRanks + 1];
    int world;
};
__device__ __forceinline__ int owner_of(const OwnerBounds &ob, unsigned long long id) {
    int o = 0;
#pragma unroll
    for (int r = 1; r < kMaxRanks; ++r)
        if (r < ob.world && id >= ob.b[r]) o = r;
    return o;
}

__global__ void __launch_bounds__(256) probes_count_kernel(const ulonglong2 *__restrict__ probes, uint64_t n, OwnerBounds ob,
                                                         unsigned long long *__restrict__ counts) {
    __shared__ unsigned int sc[kMaxRanks];
    if (threadIdx.x < kMaxRanks) sc[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned long long id = probes[i].x;
        if (id != kNoProbe) atomicAdd(&sc[owner_of(ob, id)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < ob.world && sc[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
}

// offs[o] = exclusive prefix of counts; cursor[o] = 0
__global__ void probes_offsets_kernel(const unsigned long long *__restrict__ counts, int world, unsigned long long *__restrict__ offs,
                                      unsigned long long *__restrict__ cursor) {
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int o = 0; o < world; ++o) {
            offs[o] = run;
            cursor[o] = 0;
            run += counts[o];
        }
    }
}

__global__ void __launch_bounds__(256) probes_scatter_kernel(const ulonglong2 *__restrict__ probes, uint64_t n, OwnerBounds ob,
                                                           const unsigned long long *__restrict__ offs,
                                                           unsigned long long *__restrict__ cursor, ulonglong2 *__restrict__ send,
                                                           uint32_t *__restrict__ tag) {
    __shared__ unsigned int sc[kMaxRanks];
    __shared__ unsigned long long sbase[kMaxRanks];
    if (threadIdx.x < kMaxRanks) sc[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ulonglong2 pr = make_ulonglong2(kNoProbe, 0);
    int o = -1;
    unsigned int local = 0;
    if (i < n) {
        pr = probes[i];
        if (pr.x != kNoProbe) {
            o = owner_of(ob, pr.x);
            local = atomicAdd(&sc[o], 1u);  // rank of this probe among the CTA's probes for owner o
        }
    }
    __syncthreads();
    if (threadIdx.x < ob.world && sc[threadIdx.x])  // one reservation per owner and CTA
        sbase[threadIdx.x] = offs[threadIdx.x] + atomicAdd(cursor + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
    __syncthreads();
    if (o >= 0) {
        const unsigned long long slot = sbase[o] + local;
        send[slot] = make_ulonglong2(pr.x - ob.b[o], pr.y);
        tag[slot] = (uint32_t)i;
    }
}

// ---- index upload helpers ----------------------------------------------------------------
__global__ void index23_pack_kernel(const uint64_t *__restrict__ checker, const uint32_t *__restrict__ tf, uint64_t n,
                                    uint4 *__restrict__ recs, uint8_t *__restrict__ fp, int fp_bits,
                                    int *__restrict__ non_canonical) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t c = checker[i];
    recs[i] = make_uint4((uint32_t)c, (uint32_t)(c >> 32), tf[i], 0u);
    if (fp) {
        if (fp_bits == 8) fp[i] = (uint8_t)fingerprint8(c);
        else if ((i & 1) == 0) {  // one thread writes both nibbles of a byte
            uint32_t lo = fingerprint4(c), hi = i + 1 < n ? fingerprint4(checker[i + 1]) : 0u;
            fp[i >> 1] = (uint8_t)(lo | (hi << 4));
        }
    }
    if ((c >> 46) != 0 || c > revcomp23(c)) *non_canonical = 1;
}

// fused records, step 1: pair values and ranks (record r = half-word r of the bit-pair vector; fingerprints zero)
__global__ void fused_layout_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ block_ranks,
                                    uint64_t n_words, uint64_t n_recs, uint4 *__restrict__ frecs) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_recs) return;
    const uint64_t w = r >> 1;
    uint32_t half = 0, rank = 0;
    if (w < n_words) {
        const uint64_t blk = w >> 4;
        uint64_t rk = block_ranks[blk];
        for (uint64_t i = blk << 4; i < w; ++i) rk += nonzero_pairs64(words[i]);
        const uint64_t x = words[w];
        if (r & 1) { rk += nonzero_pairs32((uint32_t)x); half = (uint32_t)(x >> 32); }
        else half = (uint32_t)x;
        rank = (uint32_t)rk;
    }
    frecs[r] = make_uint4(half, 0u, 0u, rank);
}

// step 2: the 4-bit fingerprint of every stored k-mer goes to the node the MPHF assigns to it.  A k-mer whose own
// hash does not lead back to its slot (inconsistent files) can never be found by a lookup, so it gets none.
__global__ void fused_fingerprint_kernel(MphfDev m, const uint64_t *__restrict__ checker, uint64_t n, uint4 *__restrict__ frecs) {
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    const uint64_t k = checker[h];
    // the string a lookup hashes for a stored value is the ASCII form of its low 46 bits (get_bitset_dna23); values with
    // bits above that are only ever compared by get_freq(uint64_t), which hashes the same string (hash.hpp:123-140)
    uint64_t w0, w1, w2, a, b, c;
    ascii_words23_from_rc(revcomp23(k & ((1ULL << 46) - 1)), w0, w1, w2);
    jenkins_short(m.seed, w0, w1, w2, 23u, a, b, c);
    uint32_t node = 0;
    const uint64_t id = mphf_eval_fused(m, a, b, c, &node) & kFusedIdMask;
    if (id != h) return;
    uint32_t *words = reinterpret_cast<uint32_t *>(frecs + (node >> 4));
    const uint32_t p = node & 15u;
    atomicOr(words + (p < 8u ? 1 : 2), fingerprint4(k) << ((p & 7u) * 4u));
}

__global__ void tf13_direct_kernel(MphfDev m, const uint64_t *__restrict__ tf_mphf, uint64_t *__restrict__ tf_direct) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (uint32_t)AIX_TOTAL_13MERS) return;
    uint64_t id = mphf_lookup13(m, revcomp13(v));
    tf_direct[v] = id < AIX_TOTAL_13MERS ? tf_mphf[id] : 0;  // python_wrapper.cpp:498-500
}

// ---- K4: 13-mer queries ----------------------------------------------------------------------
// python_wrapper.cpp:505-517: reverse, complement ACGT, leave anything else as it is
__device__ __forceinline__ uint32_t comp13_char(uint32_t c) {
    return c == 'A' ? 'T' : (c == 'T' ? 'A' : (c == 'G' ? 'C' : (c == 'C' ? 'G' : c)));
}

// one 13-mer query; w0 = bytes 0..7, w1 = bytes 8..12 of the record (zero padded), len = record length
template <int kMode>
__device__ __forceinline__ void query13(const MphfDev &m, const uint64_t *__restrict__ tf_mphf, const uint64_t *__restrict__ tf_direct,
                                        uint64_t w0, uint64_t w1, uint32_t len, uint64_t i, void *__restrict__ out) {
    // SIMD encode + validity (upper-case ACGT only), as encode_validate23
    const uint32_t a[4] = {(uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32)};
    uint32_t p[4], bad = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t c4 = ((a[j] >> 1) ^ (a[j] >> 2)) & 0x03030303u;
        p[j] = j == 3 ? (c4 & 3u) : ((c4 * 0x40100401u) >> 24);
        const uint32_t diff = expect_acgt4(c4) ^ a[j];
        bad |= j == 3 ? (diff & 0xFFu) : diff;
    }
    const bool valid = len == 13u && bad == 0;
    const uint32_t v = (p[0] << 18) | (p[1] << 10) | (p[2] << 2) | p[3];
    if (kMode == AIX_Q_TF) {
        // :482-503 / :938-980: len == 13 and upper-case ACGT only, value narrowed to u32
        ((uint32_t *)out)[i] = valid ? (uint32_t)__ldg(tf_direct + v) : 0u;
        return;
    }
    uint64_t fwd = 0, rev = 0;
    if (len == 13u) {
        if (valid) {
            fwd = __ldg(tf_direct + v);
            rev = __ldg(tf_direct + revcomp13(v));
        } else {
            // :533-542: no validity check -- the raw bytes are hashed; an id of 4^13 (possible
            // for non-keys) is out of bounds in the reference and defined as 0 here
            uint64_t x0 = 0, x1 = 0, ha, hb, hc;
#pragma unroll
            for (int j = 0; j < 13; ++j) {
                const int k = 12 - j;  // python_wrapper.cpp:505-517: reverse, complement ACGT, keep the rest
                const uint64_t g = comp13_char((uint32_t)((k < 8 ? w0 >> (8 * k) : w1 >> (8 * (k - 8))) & 0xFFu));
                if (j < 8) x0 |= g << (8 * j);
                else x1 |= g << (8 * (j - 8));
            }
            jenkins_short(m.seed, w0, w1, 0, 13u, ha, hb, hc);
            uint64_t id = mphf_eval(m, ha, hb, hc);
            fwd = id < AIX_TOTAL_13MERS ? tf_mphf[id] : 0;
            jenkins_short(m.seed, x0, x1, 0, 13u, ha, hb, hc);
            id = mphf_eval(m, ha, hb, hc);
            rev = id < AIX_TOTAL_13MERS ? tf_mphf[id] : 0;
        }
    }
    if (kMode == AIX_Q_TOTAL) ((uint64_t *)out)[i] = fwd + rev;
    else {
        ((uint64_t *)out)[2 * i] = fwd;
        ((uint64_t *)out)[2 * i + 1] = rev;
    }
}

template <int kMode>
__global__ void __launch_bounds__(kQBlock) tf13_kernel(MphfDev m, const uint64_t *__restrict__ tf_mphf,
                                                     const uint64_t *__restrict__ tf_direct, const uint8_t *__restrict__ recs,
                                                     uint32_t stride, const uint8_t *__restrict__ lens, uint64_t q,
                                                     void *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * kQBlock + threadIdx.x;
    if (i >= q) return;
    uint32_t len = lens ? lens[i] : stride;
    if (len > stride) len = stride;
    const uint8_t *p = recs + i * stride;
    uint64_t w0 = 0, w1 = 0;
    if (len == 13u) {
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const uint64_t ch = __ldg(p + j);
            if (j < 8) w0 |= ch << (8 * j);
            else w1 |= ch << (8 * (j - 8));
        }
    }
    query13<kMode>(m, tf_mphf, tf_direct, w0, w1, len, i, out);
}

// K4 streaming form for uint8[q, 13] batches: the TMA ring of tf23_stream_kernel with 416-byte tiles
// (32 x 13 = 26 x 16).  The per-query work is an encode and one 8-byte gather, so staging the records
// (13 single-byte loads per thread in the generic kernel) is what bounds it.
constexpr uint32_t kSt13TileBytes = 32u * 13u;
constexpr int kSt13Slot = 448;

template <int kMode>
__global__ void __launch_bounds__(kStWarps * 32) tf13_stream_kernel(MphfDev m, const uint64_t *__restrict__ tf_mphf,
                                                                  const uint64_t *__restrict__ tf_direct,
                                                                  const uint8_t *__restrict__ recs, uint64_t n_tiles,
                                                                  void *__restrict__ out) {
    __shared__ __align__(128) uint8_t slots[kStWarps][kStStages][kSt13Slot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WarpRing<kSt13TileBytes, kSt13Slot> ring;
    ring.init(&slots[wid][0][0], &bars[wid][0], lane);
    uint64_t tile0;
    const uint32_t my_tiles = warp_tiles<kStTilesPerWarp>(n_tiles, wid, tile0);
    if (my_tiles == 0) return;
    ring.start(recs + tile0 * kSt13TileBytes, my_tiles, lane);
    uint64_t i = tile0 * 32u + lane;
    for (uint32_t it = 0; it < my_tiles; ++it) {
        const uint32_t tile = ring.acquire(it, lane);
        const uint32_t base = lane * 13u, sh = (base & 3u) * 8u;
        uint32_t x0, x1, x2, x3;  // bytes base .. base+12 end inside the fourth word
        asm volatile("ld.shared.u32 %0, [%4];\nld.shared.u32 %1, [%4+4];\nld.shared.u32 %2, [%4+8];\nld.shared.u32 %3, [%4+12];"
                     : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(tile + (base & ~3u)) : "memory");
        __syncwarp();
        const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
                       y3 = (x3 >> sh) & 0xFFu;
        query13<kMode>(m, tf_mphf, tf_direct, ((uint64_t)y1 << 32) | y0, ((uint64_t)y3 << 32) | y2, 13u, i, out);
        i += (uint64_t)kStWarps * 32u;
        ring.advance();
    }
}

static size_t out_bytes23(int mode) {
    switch (mode) {
        case AIX_Q_TF: return 4;
        case AIX_Q_BOTH: return 8;
        default: return 8;
    }
}

// 1: streaming kernel (default), 0: one CTA per 256 queries (AIX_TF23_KERNEL, kept for A/B runs)
static int tf23_kernel_choice() {
    const char *e = getenv("AIX_TF23_KERNEL");
    return e ? atoi(e) : 1;
}

// register budget of the streaming kernel: what ptxas picks (56 registers, 4 resident CTAs, no spills;
// default) or 6 resident CTAs (40 registers, two spilled words).  The kernel is ALU-pipe bound, not
// latency bound: the unconstrained build is 8 % faster (profiles/r01_tf23_sweep.txt).
// AIX_TF23_MINBLOCKS selects for A/B runs.
static int tf23_min_blocks() {
    const char *e = getenv("AIX_TF23_MINBLOCKS");
    return e ? atoi(e) : 1;
}

template <int kMode, bool kCanon, int kMinBlocks>
static void launch_stream(const aix_ctx *ctx, cudaStream_t st, const Index23Dev &id, const MphfDev &md, const uint8_t *recs,
                          uint64_t n_tiles, void *out) {
    (void)ctx;
    const uint64_t grid = (n_tiles + kStTilesPerCta - 1) / kStTilesPerCta;
    tf23_stream_kernel<kMode, kCanon, kMinBlocks><<<(unsigned)grid, kStWarps * 32, 0, st>>>(id, md, recs, n_tiles, out);
}


// Does the next batch go through the front filter?  AIX_INDEX23_FILTER=on|off forces; otherwise the pass rate the filter
// kernel reported decides: unknown -> yes (the first batches are the probe), <= 25 % passed -> yes, else the direct
// kernel with one filtered batch in 32 to notice when the traffic changes.  (With more than a quarter of the queries
// passing, the filter's extra request and the second pass over the passing queries cost more than the filter saves.)
static bool filter_wanted(const aix_index23 *ix) {
    if (!ix->bloom_dev || !ix->canonical_only) return false;
    static int env_mode = -1;  // 0 auto, 1 on, 2 off
    if (env_mode < 0) {
        const char *e = getenv("AIX_INDEX23_FILTER");
        env_mode = !e ? 0 : (!strcmp(e, "on") ? 1 : (!strcmp(e, "off") ? 2 : 0));
    }
    const int forced = ix->filter_mode ? ix->filter_mode : env_mode;
    if (forced) return forced == 1;
    const unsigned long long q = ix->qstats_host[0], p = ix->qstats_host[1];
    if (q > ix->seen_q && p >= ix->seen_p && p - ix->seen_p <= q - ix->seen_q) {
        ix->pass_rate = (double)(p - ix->seen_p) / (double)(q - ix->seen_q);
        ix->seen_q = q;
        ix->seen_p = p;
    }
    if (ix->pass_rate < 0 || ix->pass_rate <= 0.25) return true;
    if (++ix->direct_since_probe >= 32) {
        ix->direct_since_probe = 0;
        return true;
    }
    return false;
}

// Access-policy window over the front filter for the launches that follow on `st`: its lines are kept in the persisting
// part of L2 (set aside once per device, as much as the filter needs and the device allows); everything else the kernel
// touches streams past them.  Returns false when the device cannot do it.
static bool l2_window_on(const aix_ctx *ctx, cudaStream_t st, const void *base, size_t bytes) {
    static int max_persist = -1, max_window = 0;
    if (max_persist < 0) {
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    }
    if (max_persist <= 0 || max_window <= 0 || bytes == 0) return false;
    size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
    const char *esa = getenv("AIX_FILTER_SETASIDE_MB"), *ehr = getenv("AIX_FILTER_HITRATIO");  // A/B knobs
    if (esa && atol(esa) > 0) want = (size_t)atol(esa) << 20 < (size_t)max_persist ? (size_t)atol(esa) << 20 : (size_t)max_persist;
    AixL2State &l2 = g_aix_l2[ctx->device & 63];
    if (l2.set_aside.load() < want) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        l2.set_aside = want;
    }
    const size_t set_aside = l2.set_aside.load();
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    v.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    v.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    v.accessPolicyWindow.hitRatio = ehr ? (float)atof(ehr) : (bytes <= set_aside ? 1.0f : (float)((double)set_aside / (double)bytes));
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    l2.pinned = true;
    return true;
}
static void l2_window_off(cudaStream_t st) {
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
}

template <int kMode>
static void launch23_mode(const aix_ctx *ctx, const aix_index23 *ix, cudaStream_t st, const uint8_t *recs, uint32_t stride,
                          const uint8_t *lens, uint64_t q, void *out) {
    Index23Dev id = ix->dev();
    MphfDev md = ix->mphf_dev();
    const bool fixed = (stride == 23 && lens == nullptr && ((uintptr_t)recs & 15) == 0);
    if (fixed && tf23_kernel_choice() == 1 && q >= 32) {
        const uint64_t n_tiles = q / 32;
        uint64_t done = n_tiles * 32;  // queries answered by the streaming kernel
        if (kMode == AIX_Q_TF && q >= 4096 && filter_wanted(ix)) {
            // Which filter kernel (tf23_filter.cuh; profiles/r02_filter_sweep.txt has every number):
            //   3 (default)  filter word LOADED a tile ahead, loop unrolled by two, ptxas held to 64 registers / 4 resident CTAs:
            //                99.1 G q/s, and 102.7 with the filter's lines held in persisting L2 by an access-policy window
            //                (AIX_FILTER_PERSIST=0 switches the window off; a larger set-aside or a hit ratio < 1 are slower)
            //   1            word prefetched into L1 a tile ahead: 91.4 (39 % of the stall samples on the word's first use)
            //   2            two queries per lane and iteration (needs an 8-byte aligned `out`): 17 % fewer instructions,
            //                72 registers or 60 with ptxas held to 4 CTAs, 83.8 / 87.5 -- the kernel lives on resident warps
            // AIX_FILTER_MINBLOCKS=1 lifts the register cap of 2 and 3.  Read per launch (the parity test switches kernels
            // inside one process).
            const char *ek = getenv("AIX_FILTER_KERNEL"), *em = getenv("AIX_FILTER_MINBLOCKS"), *ep = getenv("AIX_FILTER_PERSIST");
            const int fker = ek ? atoi(ek) : 3, fmin = em ? atoi(em) : 4;
            uint32_t *out32 = (uint32_t *)out;
            unsigned long long *qs = ix->qstats_dev;
            const unsigned g = (unsigned)((n_tiles + kStTilesPerCta - 1) / kStTilesPerCta), threads = kStWarps * 32;
            if (fker == 3) {
                if (fmin >= 4 && !(ep && atoi(ep) == 0) && l2_window_on(ctx, st, ix->bloom_dev, (size_t)ix->bloom_words * 8)) {
                    tf23_filter3_kernel<4, kStTilesPerWarp, true><<<g, threads, 0, st>>>(id, md, recs, n_tiles, out32, qs);
                    l2_window_off(st);
                } else if (fmin >= 4) tf23_filter3_kernel<4, kStTilesPerWarp><<<g, threads, 0, st>>>(id, md, recs, n_tiles, out32, qs);
                else tf23_filter3_kernel<1, kStTilesPerWarp><<<g, threads, 0, st>>>(id, md, recs, n_tiles, out32, qs);
            } else if (fker == 2 && ((uintptr_t)out & 7) == 0) {
                const uint64_t n64 = q / 64;  // tiles of 64 queries, half as many per warp
                const unsigned g2 = (unsigned)((n64 + kStTilesPerCta / 2 - 1) / (kStTilesPerCta / 2));
                if (fmin >= 4) tf23_filter2_kernel<4, kStTilesPerWarp / 2><<<g2, threads, 0, st>>>(id, md, recs, n64, out32, qs);
                else tf23_filter2_kernel<1, kStTilesPerWarp / 2><<<g2, threads, 0, st>>>(id, md, recs, n64, out32, qs);
                done = n64 * 64;
            } else {
                tf23_filter_kernel<1><<<g, threads, 0, st>>>(id, md, recs, n_tiles, out32, qs);
            }
            // the counts travel back behind the kernel; the next launches read whatever has arrived
            cudaMemcpyAsync((void *)ix->qstats_host, ix->qstats_dev, 16, cudaMemcpyDeviceToHost, st);
            ix->launches_filter++;
        } else if (ix->canonical_only) {
            ix->launches_direct++;
            if (tf23_min_blocks() >= 6) launch_stream<kMode, true, 6>(ctx, st, id, md, recs, n_tiles, out);
            else if (tf23_min_blocks() == 5) launch_stream<kMode, true, 5>(ctx, st, id, md, recs, n_tiles, out);
            else launch_stream<kMode, true, 1>(ctx, st, id, md, recs, n_tiles, out);
        } else {
            launch_stream<kMode, false, 6>(ctx, st, id, md, recs, n_tiles, out);
        }
        // the last q % 32 (% 64) queries: the same per-query code through the block kernel (its tile start stays 16-byte aligned)
        if (done == q) return;
        recs += done * 23;
        out = (char *)out + done * out_bytes23(kMode);
        q -= done;
    }
    unsigned grid = aix_grid(q, kQBlock);
    if (fixed) {
        if (ix->canonical_only) tf23_fixed_kernel<kMode, true><<<grid, kQBlock, 0, st>>>(id, md, recs, q, out);
        else tf23_fixed_kernel<kMode, false><<<grid, kQBlock, 0, st>>>(id, md, recs, q, out);
    } else {
        if (ix->canonical_only) tf23_generic_kernel<kMode, true><<<grid, kQBlock, 0, st>>>(id, md, recs, stride, lens, q, out);
        else tf23_generic_kernel<kMode, false><<<grid, kQBlock, 0, st>>>(id, md, recs, stride, lens, q, out);
    }
}

int launch_tf23(aix_ctx *ctx, const aix_index23 *ix, cudaStream_t st, const uint8_t *recs, uint32_t stride,
                const uint8_t *lens, uint64_t q, int mode, void *out) {
    if (q == 0) return AIX_OK;
    switch (mode) {
        case AIX_Q_TF: launch23_mode<AIX_Q_TF>(ctx, ix, st, recs, stride, lens, q, out); break;
        case AIX_Q_TOTAL: launch23_mode<AIX_Q_TOTAL>(ctx, ix, st, recs, stride, lens, q, out); break;
        case AIX_Q_BOTH: launch23_mode<AIX_Q_BOTH>(ctx, ix, st, recs, stride, lens, q, out); break;
        case AIX_Q_PFID: launch23_mode<AIX_Q_PFID>(ctx, ix, st, recs, stride, lens, q, out); break;
        case AIX_Q_STRAND: launch23_mode<AIX_Q_STRAND>(ctx, ix, st, recs, stride, lens, q, out); break;
        case AIX_Q_KID: launch23_mode<AIX_Q_KID>(ctx, ix, st, recs, stride, lens, q, out); break;
        default: return ctx->fail(AIX_ERR_ARG, "unknown query mode %d", mode);
    }
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

template <int kMode>
static void launch13_mode(const aix_index13 *ix, cudaStream_t st, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                          uint64_t q, void *out, size_t out_bytes) {
    MphfDev md = ix->mphf->dev();
    const bool fixed = (stride == 13 && lens == nullptr && ((uintptr_t)recs & 15) == 0);
    if (fixed && tf23_kernel_choice() == 1 && q >= 32) {
        const uint64_t n_tiles = q / 32;
        const uint64_t grid = (n_tiles + kStTilesPerCta - 1) / kStTilesPerCta;
        tf13_stream_kernel<kMode><<<(unsigned)grid, kStWarps * 32, 0, st>>>(md, ix->tf_mphf_dev, ix->tf_direct_dev, recs, n_tiles, out);
        const uint64_t done = n_tiles * 32;
        if (done == q) return;
        recs += done * 13;
        out = (char *)out + done * out_bytes;
        q -= done;
    }
    tf13_kernel<kMode><<<aix_grid(q, kQBlock), kQBlock, 0, st>>>(md, ix->tf_mphf_dev, ix->tf_direct_dev, recs, stride, lens, q, out);
}

int launch_tf13(aix_ctx *ctx, const aix_index13 *ix, cudaStream_t st, const uint8_t *recs, uint32_t stride,
                const uint8_t *lens, uint64_t q, int mode, void *out) {
    if (q == 0) return AIX_OK;
    switch (mode) {
        case AIX_Q_TF: launch13_mode<AIX_Q_TF>(ix, st, recs, stride, lens, q, out, 4); break;
        case AIX_Q_TOTAL: launch13_mode<AIX_Q_TOTAL>(ix, st, recs, stride, lens, q, out, 8); break;
        case AIX_Q_BOTH: launch13_mode<AIX_Q_BOTH>(ix, st, recs, stride, lens, q, out, 16); break;
        default: return ctx->fail(AIX_ERR_ARG, "unknown 13-mer query mode %d", mode);
    }
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}


// ---- single-query mailbox ---------------------------------------------------------------------------------
// The reference's scripts loop over get_tf_value(kmer) (python_wrapper.cpp:644-650, 0.5 us per call on the CPU).
// A kernel launch plus a stream synchronisation per call costs 13 us; here a ONE-THREAD kernel stays resident for as
// long as single calls keep coming, polls a request slot in mapped host memory, answers through the same slot and the
// host spins on the answer: two PCIe traversals per call.  The kernel bounds its own life (idle timeout on
// %globaltimer AND a hard cap of empty polls), so an implicit device synchronisation elsewhere waits a millisecond at
// most; the host relaunches it on demand and falls back to the launch path if it ever fails to answer.
// Slot layout (192 bytes of mapped host memory).  Request: up to six 16-byte chunks, each = 12 payload bytes + the
// sequence number; the host writes every chunk with ONE 16-byte store.  Payload bytes 0..1 = length, kind; bytes 2.. = the
// query.  The device polls chunk 0 ALONE: reads of host memory queue behind each other on the bus (measured on this box,
// profiles/r02_mailbox.txt: one 16-byte read in flight 3.8 us per echo, three 4.6 us, twelve 10 us), so one read per poll
// it is.  An upper-case ACGT 13- or 23-mer -- every query that can hit -- travels 2-bit packed inside chunk 0 (kinds
// kMboxPacked23 / 13) and is answered after that one read; any other string needs chunks 1..2 (3..5 beyond 34 bytes),
// fetched once chunk 0's tag has changed and accepted when their tags agree, so no ordering between chunks is needed.
// Response: ONE 64-bit word {tf, sequence number} written with one store -- no fence, no second word.  `alive` is cleared
// by the kernel's last store.
struct MboxSlot {
    uint32_t req[6][4];      // [c][0..2] payload, [c][3] tag
    unsigned long long resp; // (seq << 32) | tf
    uint32_t quit, alive;
    uint8_t fill[80];
};
static_assert(sizeof(MboxSlot) == 192, "mailbox slot layout");
constexpr uint32_t kMboxPayload = 6 * 12 - 2;  // query bytes a request can carry

__device__ __forceinline__ uint4 ld_relaxed_sys_u32x4(const void *p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr uint32_t kMboxEcho = 0xEEu;      // answered with its length at once (transport latency probe)
constexpr uint32_t kMboxPacked23 = 0xB7u;  // payload bytes 2..9 = the 23-mer, code of character j ("ACGT" order) at bits [2j+1:2j]
constexpr uint32_t kMboxPacked13 = 0xBDu;  // same for a 13-mer

__global__ void __launch_bounds__(32) mailbox_kernel(Index23Dev ix, MphfDev m23, MphfDev m13, const uint64_t *tf13_mphf,
                                                   const uint64_t *tf13_direct, MboxSlot *slot, uint32_t seen,
                                                   unsigned long long idle_ns, uint32_t max_empty_polls) {
    if (threadIdx.x != 0) return;
    uint64_t t_last = global_timer_ns();
    uint32_t empty = 0;
    for (;;) {
        const uint4 c0 = ld_relaxed_sys_u32x4(slot->req[0]);
        const uint32_t s = c0.w;
        if (s == seen) {  // nothing new
            if ((++empty & 15u) == 0u) {
                if (ld_relaxed_sys_u32(&slot->quit) != 0u) break;
                if (empty > max_empty_polls || global_timer_ns() - t_last > idle_ns) break;
            }
            continue;
        }
        uint32_t b[18];  // payload words
#pragma unroll
        for (int j = 3; j < 18; ++j) b[j] = 0;
        b[0] = c0.x; b[1] = c0.y; b[2] = c0.z;
        const uint32_t len0 = b[0] & 0xFFu;
        uint32_t kind = (b[0] >> 8) & 0xFFu;
        uint32_t len = len0 > kMboxPayload ? kMboxPayload : len0;
        const bool one_chunk = kind == kMboxEcho || kind == kMboxPacked23 || kind == kMboxPacked13;
        if (!one_chunk) {  // a string: chunks 1..2 (their tags must agree; the host stored them before chunk 0)
            uint4 c1, c2;
            do {
                c1 = ld_relaxed_sys_u32x4(slot->req[1]); c2 = ld_relaxed_sys_u32x4(slot->req[2]);
            } while (c1.w != s || c2.w != s);
            b[3] = c1.x; b[4] = c1.y; b[5] = c1.z; b[6] = c2.x; b[7] = c2.y; b[8] = c2.z;
            if (len > 34u) {
                uint4 c3, c4, c5;
                do {
                    c3 = ld_relaxed_sys_u32x4(slot->req[3]); c4 = ld_relaxed_sys_u32x4(slot->req[4]); c5 = ld_relaxed_sys_u32x4(slot->req[5]);
                } while (c3.w != s || c4.w != s || c5.w != s);
                b[9] = c3.x; b[10] = c3.y; b[11] = c3.z; b[12] = c4.x; b[13] = c4.y; b[14] = c4.z; b[15] = c5.x; b[16] = c5.y; b[17] = c5.z;
            }
        }
        // the query bytes start at payload byte 2: realign into words
        uint32_t q[17];
#pragma unroll
        for (int j = 0; j < 17; ++j) q[j] = __funnelshift_r(b[j], b[j + 1], 16);
        if (kind == kMboxPacked23 || kind == kMboxPacked13) {  // back to the string the caller passed: same path from here on
            const uint64_t w0 = ascii8_from_codes_le(q[0] & 0xFFFFu), w1 = ascii8_from_codes_le(q[0] >> 16),
                           w2 = ascii8_from_codes_le(q[1] & 0xFFFFu);
            len = kind == kMboxPacked23 ? 23u : 13u;
            kind = len;
            q[0] = (uint32_t)w0; q[1] = (uint32_t)(w0 >> 32); q[2] = (uint32_t)w1; q[3] = (uint32_t)(w1 >> 32);
            q[4] = (uint32_t)w2; q[5] = (uint32_t)(w2 >> 32);
        }
        uint64_t res[2] = {0, 0};
        if (kind == kMboxEcho) {
            res[0] = len0;
        } else if (kind == 23u) {
            const uint64_t mask2 = len >= 23u ? 0x00FFFFFFFFFFFFFFull : (len > 16u ? ((1ull << (8 * (len - 16u))) - 1) : 0ull);
            const uint64_t mask1 = len >= 16u ? ~0ull : (len > 8u ? ((1ull << (8 * (len - 8u))) - 1) : 0ull);
            const uint64_t mask0 = len >= 8u ? ~0ull : ((1ull << (8 * len)) - 1);
            const uint64_t r0 = (((uint64_t)q[1] << 32) | q[0]) & mask0, r1 = (((uint64_t)q[3] << 32) | q[2]) & mask1,
                           r2 = (((uint64_t)q[5] << 32) | q[4]) & mask2;
            const uint8_t *p = reinterpret_cast<const uint8_t *>(q);
            if (ix.canonical_only) query23<AIX_Q_TF, true>(ix, m23, r0, r1, r2, len, p, 0, res);
            else query23<AIX_Q_TF, false>(ix, m23, r0, r1, r2, len, p, 0, res);
        } else {
            const uint64_t mask1 = len >= 13u ? 0x000000FFFFFFFFFFull : (len > 8u ? ((1ull << (8 * (len - 8u))) - 1) : 0ull);
            const uint64_t mask0 = len >= 8u ? ~0ull : ((1ull << (8 * len)) - 1);
            const uint64_t w0 = (((uint64_t)q[1] << 32) | q[0]) & mask0, w1 = (((uint64_t)q[3] << 32) | q[2]) & mask1;
            query13<AIX_Q_TF>(m13, tf13_mphf, tf13_direct, w0, w1, len, 0, res);
        }
        const unsigned long long word = ((unsigned long long)s << 32) | (res[0] & 0xFFFFFFFFull);
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&slot->resp), "l"(word) : "memory");
        seen = s;
        empty = 0;
        t_last = global_timer_ns();
    }
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(&slot->alive), "r"(0u) : "memory");
}

void mbox_stop(aix_ctx *ctx) {
    if (!ctx || !ctx->mbox_launched) return;
    MboxSlot *slot = (MboxSlot *)ctx->mbox_host;
    __atomic_store_n(&slot->quit, 1u, __ATOMIC_RELEASE);
    cudaStreamSynchronize(ctx->mbox_stream);
    __atomic_store_n(&slot->quit, 0u, __ATOMIC_RELEASE);
    ctx->mbox_launched = false;
    ctx->mbox_owner = nullptr;
}

static bool mbox_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("AIX_MAILBOX");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on == 1;
}

// one TF query (k = 23 on ix23, or k = 13 on ix13) through the mailbox; false = not answered (caller takes the launch path)
static bool mbox_query(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13, const uint8_t *rec, uint32_t len, uint32_t *out,
                       bool echo = false) {
    if (!mbox_enabled() || ctx->mbox_broken || len > kMboxPayload) return false;
    if (!ctx->mbox_host) {
        if (cudaSetDevice(ctx->device) != cudaSuccess) return false;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaHostAlloc(&ctx->mbox_host, sizeof(MboxSlot), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&ctx->mbox_dev, ctx->mbox_host, 0) != cudaSuccess ||
            cudaStreamCreateWithPriority(&ctx->mbox_stream, cudaStreamNonBlocking, lo) != cudaSuccess) {
            cudaGetLastError();
            ctx->mbox_broken = true;
            return false;
        }
        memset(ctx->mbox_host, 0, sizeof(MboxSlot));
    }
    MboxSlot *slot = (MboxSlot *)ctx->mbox_host;
    const void *owner = ix23 ? (const void *)ix23 : (const void *)ix13;
    auto launch = [&](uint32_t seen) -> bool {
        if (cudaSetDevice(ctx->device) != cudaSuccess) return false;
        if (ctx->mbox_launched) cudaStreamSynchronize(ctx->mbox_stream);  // the previous kernel has said it is leaving
        Index23Dev id = {};
        MphfDev m23 = {}, m13 = {};
        const uint64_t *t_m = nullptr, *t_d = nullptr;
        if (ix23) { id = ix23->dev(); m23 = ix23->mphf_dev(); }
        else { m13 = ix13->mphf->dev(); t_m = ix13->tf_mphf_dev; t_d = ix13->tf_direct_dev; }
        __atomic_store_n(&slot->alive, 1u, __ATOMIC_RELEASE);
        mailbox_kernel<<<1, 32, 0, ctx->mbox_stream>>>(id, m23, m13, t_m, t_d, (MboxSlot *)ctx->mbox_dev, seen,
                                                       2000000ull /* 2 ms idle */, 400000u);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) return false;
        ctx->mbox_launched = true;
        ctx->mbox_owner = owner;
        return true;
    };
    if (ctx->mbox_launched && ctx->mbox_owner != owner) mbox_stop(ctx);
    if (!ctx->mbox_launched || __atomic_load_n(&slot->alive, __ATOMIC_ACQUIRE) == 0u) {
        if (!launch(ctx->mbox_seq)) {
            ctx->mbox_broken = true;
            return false;
        }
    }
    // request: 16-byte chunks {12 payload bytes, sequence number}.  An upper-case ACGT k-mer of the index's k goes 2-bit
    // packed in chunk 0 alone; anything else as its bytes in three (six beyond 34 bytes) chunks.
    const uint32_t seq = ++ctx->mbox_seq;
    alignas(16) uint8_t pay[72] = {0};
    int n_chunks;
    uint64_t packed = 0;
    bool acgt = !echo && len == (ix23 ? 23u : 13u);
    for (uint32_t j = 0; acgt && j < len; ++j) {
        const uint8_t ch = rec[j];
        const uint32_t code = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
        if (code > 3u) acgt = false;
        packed |= (uint64_t)(code & 3u) << (2 * j);
    }
    pay[0] = (uint8_t)len;
    if (acgt) {
        pay[1] = (uint8_t)(ix23 ? kMboxPacked23 : kMboxPacked13);
        memcpy(pay + 2, &packed, 8);
        n_chunks = 1;
    } else {
        pay[1] = echo ? (uint8_t)kMboxEcho : (ix23 ? 23 : 13);
        memcpy(pay + 2, rec, len);
        n_chunks = echo ? 1 : (len > 34u ? 6 : 3);
    }
    for (int c = n_chunks - 1; c >= 0; --c) {  // chunk 0 (the one the device polls) last
        alignas(16) uint32_t w[4];
        memcpy(w, pay + 12 * c, 12);
        w[3] = seq;
#if defined(__x86_64__)
        _mm_store_si128((__m128i *)slot->req[c], _mm_load_si128((const __m128i *)w));  // one 16-byte store
#else
        memcpy(slot->req[c], w, 12);
        __atomic_store_n(&slot->req[c][3], seq, __ATOMIC_RELEASE);  // tag after its payload
#endif
    }
    __atomic_thread_fence(__ATOMIC_RELEASE);
    const double t0 = AixTrace::now();
    for (uint32_t spin = 1;; ++spin) {
        const unsigned long long r = __atomic_load_n(&slot->resp, __ATOMIC_ACQUIRE);
        if ((uint32_t)(r >> 32) == seq) {
            *out = (uint32_t)r;
            return true;
        }
        if ((spin & 0xFFu) == 0) {
            if (__atomic_load_n(&slot->alive, __ATOMIC_ACQUIRE) == 0u) {  // the kernel left (its idle timeout raced with this request)
                const unsigned long long r2 = __atomic_load_n(&slot->resp, __ATOMIC_ACQUIRE);
                if ((uint32_t)(r2 >> 32) == seq) { *out = (uint32_t)r2; return true; }
                if (!launch(seq - 1)) { ctx->mbox_broken = true; return false; }
            }
            if ((spin & 0xFFFFu) == 0 && AixTrace::now() - t0 > 2.0) {  // never hang the caller: give the mailbox up for this ctx
                mbox_stop(ctx);
                ctx->mbox_broken = true;
                return false;
            }
        }
    }
}

}  // namespace aix

using namespace aix;

static int read_file(aix_ctx *ctx, const std::string &path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return ctx->fail(AIX_ERR_IO, "required file not found: %s", path.c_str());
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    bool ok = buf.empty() || fread(buf.data(), 1, buf.size(), f) == buf.size();
    fclose(f);
    return ok ? AIX_OK : ctx->fail(AIX_ERR_IO, "short read: %s", path.c_str());
}

extern "C" {

int aix_index23_upload_dev(aix_ctx *ctx, const aix_mphf *m, const uint64_t *checker_dev, const uint32_t *tf_dev,
                           uint64_t n, aix_index23 **out) {
    if (!ctx || !m || !out) return AIX_ERR_ARG;
    *out = nullptr;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    aix_index23 *ix = new aix_index23();
    ix->n = n; ix->mphf = m;
    int *flag_dev = nullptr;
    cudaError_t e = cudaMalloc(&ix->recs_dev, (n ? n : 1) * sizeof(uint4));
    if (e == cudaSuccess) e = cudaMalloc(&flag_dev, sizeof(int));
    if (e != cudaSuccess) {
        cudaGetLastError();
        aix_index23_destroy(ctx, ix);
        return ctx->fail(AIX_ERR_NOMEM, "index23 upload: %s", cudaGetErrorString(e));
    }
    // fingerprint tier: must stay in L2 next to the MPHF records.  Random accesses from all SMs see
    // about half of the 126 MB L2 (two partitions, lines replicated on the far side): 8-bit
    // fingerprints when both structures fit in the budget, else 4-bit, else no tier.
    uint64_t fp_budget = 72ull << 20;  // measured: C2 (50 M keys, 70.5 MB with 8-bit fingerprints) still runs best with 8 bits (profiles/r01_tf23_sweep.txt)
    if (const char *e = getenv("AIX_FP_TIER_MAX_BYTES")) fp_budget = strtoull(e, nullptr, 10);
    uint64_t fp_bytes = 0;
    // fused layout (fingerprints inside the MPHF records, 3 scattered requests per query instead of 4): when the
    // 16-byte-per-16-nodes records fit the L2 budget.  AIX_INDEX23_LAYOUT=tier|fused overrides (A/B runs, tests).
    bool fused = n && n == m->n && m->crecs_dev != nullptr && m->bv_size + 32 <= (64ull << 20);  // (a slice of a sharded index keeps the slot-indexed tier)
    if (const char *e = getenv("AIX_INDEX23_LAYOUT")) {
        if (!strcmp(e, "tier")) fused = false;
        else if (!strcmp(e, "fused")) fused = n && n == m->n && m->bv_size < (1ull << 32) && m->hash_domain < (1ull << 31) && m->n < (1ull << 32);
    }
    if (getenv("AIX_FP_TIER_BITS")) fused = false;  // an explicit tier request is a tier request
    if (fused) {
        const uint64_t n_recs = (m->bv_size + 15) / 16 + 1;
        uint64_t *words_dev = nullptr, *ranks_dev = nullptr;
        cudaError_t ef = cudaMalloc(&ix->frecs_dev, n_recs * sizeof(uint4));
        if (ef == cudaSuccess) ef = cudaMalloc(&words_dev, (m->n_words ? m->n_words : 1) * 8);
        if (ef == cudaSuccess) ef = cudaMalloc(&ranks_dev, (m->n_blocks ? m->n_blocks : 1) * 8);
        if (ef == cudaSuccess) ef = cudaMemcpyAsync(words_dev, m->words.data(), m->n_words * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (ef == cudaSuccess) ef = cudaMemcpyAsync(ranks_dev, m->block_ranks.data(), m->n_blocks * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (ef == cudaSuccess) {
            fused_layout_kernel<<<aix_grid(n_recs, 256), 256, 0, ctx->stream>>>(words_dev, ranks_dev, m->n_words, n_recs, ix->frecs_dev);
            MphfDev md = m->dev();
            md.frecs = ix->frecs_dev;
            fused_fingerprint_kernel<<<aix_grid(n, 256), 256, 0, ctx->stream>>>(md, checker_dev, n, ix->frecs_dev);
            ctx->launches += 2;
            ef = cudaStreamSynchronize(ctx->stream);
        }
        cudaFree(words_dev);
        cudaFree(ranks_dev);
        if (ef != cudaSuccess) {  // no room: keep the separate tier logic below
            cudaGetLastError();
            cudaFree(ix->frecs_dev);
            ix->frecs_dev = nullptr;
            fused = false;
        } else {
            ix->frecs_bytes = n_recs * sizeof(uint4);
        }
    }
    if (fused) { /* no separate tier */ }
    else if (n && n + m->layout_bytes <= fp_budget) { ix->fp_bits = 8; fp_bytes = n; }
    else if (n && (n + 1) / 2 + m->layout_bytes <= fp_budget) { ix->fp_bits = 4; fp_bytes = (n + 1) / 2; }
    if (const char *e = fused ? nullptr : getenv("AIX_FP_TIER_BITS")) {  // test / experiment hook: 0, 4 or 8
        int b = atoi(e);
        ix->fp_bits = (b == 4 || b == 8) ? b : 0;
        fp_bytes = ix->fp_bits == 8 ? n : (ix->fp_bits == 4 ? (n + 1) / 2 : 0);
    }
    if (fp_bytes) {
        if (cudaMalloc(&ix->fp_dev, fp_bytes) != cudaSuccess) {
            cudaGetLastError();
            ix->fp_dev = nullptr;
            ix->fp_bits = 0;
        }
    } else {
        ix->fp_bits = 0;
    }
    int flag = 0;
    cudaMemsetAsync(flag_dev, 0, sizeof(int), ctx->stream);
    if (n) {
        index23_pack_kernel<<<aix_grid(n, 256), 256, 0, ctx->stream>>>(checker_dev, tf_dev, n, ix->recs_dev, ix->fp_dev, ix->fp_bits, flag_dev);
        ctx->launches++;
    }
    cudaMemcpyAsync(&flag, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(flag_dev);
    if (e != cudaSuccess) {
        aix_index23_destroy(ctx, ix);
        return ctx->fail(AIX_ERR_CUDA, "index23 upload: %s", cudaGetErrorString(e));
    }
    ix->canonical_only = flag ? 0 : 1;
    // front filter of the batch path: canonical-only whole indexes, AIX_BLOOM_BITS per key (default 8, 0 = none)
    int bloom_bits = 8;
    if (const char *eb = getenv("AIX_BLOOM_BITS")) bloom_bits = atoi(eb);
    if (ix->canonical_only && n >= 1024 && n == m->n && bloom_bits > 0 && bloom_bits <= 64) {
        const uint64_t words = (n * (uint64_t)bloom_bits + 63) / 64;
        if (words < (1ull << 32)) {
            cudaError_t eb = cudaMalloc(&ix->bloom_dev, words * 8);
            if (eb == cudaSuccess) eb = cudaMalloc(&ix->qstats_dev, 16);
            if (eb == cudaSuccess) eb = cudaHostAlloc((void **)&ix->qstats_host, 16, cudaHostAllocDefault);
            if (eb == cudaSuccess) {
                ix->qstats_host[0] = ix->qstats_host[1] = 0;
                ix->bloom_words = (uint32_t)words;
                cudaMemsetAsync(ix->bloom_dev, 0, words * 8, ctx->stream);
                cudaMemsetAsync(ix->qstats_dev, 0, 16, ctx->stream);
                bloom_build_kernel<<<aix_grid(n, 256), 256, 0, ctx->stream>>>(ix->recs_dev, n, (unsigned long long *)ix->bloom_dev, ix->bloom_words);
                ctx->launches++;
                eb = cudaStreamSynchronize(ctx->stream);
            }
            if (eb != cudaSuccess) {  // no room: the index works without the filter
                cudaGetLastError();
                if (ix->bloom_dev) cudaFree(ix->bloom_dev);
                if (ix->qstats_dev) cudaFree(ix->qstats_dev);
                if (ix->qstats_host) cudaFreeHost((void *)ix->qstats_host);
                ix->bloom_dev = nullptr; ix->qstats_dev = nullptr; ix->qstats_host = nullptr; ix->bloom_words = 0;
            }
        }
    }
    *out = ix;
    return AIX_OK;
}

int aix_index23_upload(aix_ctx *ctx, const aix_mphf *m, const uint64_t *checker, const uint32_t *tf, uint64_t n,
                       aix_index23 **out) {
    if (!ctx || !m || !out || (n && (!checker || !tf))) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *c_dev = nullptr;
    uint32_t *t_dev = nullptr;
    cudaError_t e = cudaMalloc(&c_dev, (n ? n : 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&t_dev, (n ? n : 1) * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (c_dev) cudaFree(c_dev);
        return ctx->fail(AIX_ERR_NOMEM, "index23 staging: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(c_dev, checker, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(t_dev, tf, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    int rc = aix_index23_upload_dev(ctx, m, c_dev, t_dev, n, out);
    cudaFree(c_dev);
    cudaFree(t_dev);
    return rc;
}

int aix_index23_load_prefix(aix_ctx *ctx, const char *prefix, aix_mphf **mphf_out, aix_index23 **out) {
    if (!ctx || !prefix || !mphf_out || !out) return AIX_ERR_ARG;
    *mphf_out = nullptr; *out = nullptr;
    std::string p(prefix);
    std::vector<uint8_t> kb, tb;
    AIX_TRY(read_file(ctx, p + ".kmers.bin", kb));
    AIX_TRY(read_file(ctx, p + ".tf.bin", tb));
    uint64_t n = kb.size() / 8;  // hash.cpp:388-392
    if (tb.size() / 4 < n) return ctx->fail(AIX_ERR_IO, "%s.tf.bin shorter than %s.kmers.bin", prefix, prefix);
    aix_mphf *m = nullptr;
    AIX_TRY(aix_mphf_load_pf(ctx, (p + ".pf").c_str(), &m));
    int rc = aix_index23_upload(ctx, m, (const uint64_t *)kb.data(), (const uint32_t *)tb.data(), n, out);
    if (rc != AIX_OK) {
        aix_mphf_destroy(ctx, m);
        return rc;
    }
    *mphf_out = m;
    return AIX_OK;
}

void aix_index23_destroy(aix_ctx *ctx, aix_index23 *ix) {
    if (!ix) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (ctx && ctx->mbox_owner == ix) mbox_stop(ctx);
    if (ix->recs_dev) cudaFree(ix->recs_dev);
    if (ix->fp_dev) cudaFree(ix->fp_dev);
    if (ix->frecs_dev) cudaFree(ix->frecs_dev);
    if (ix->bloom_dev) cudaFree(ix->bloom_dev);
    if (ix->qstats_dev) cudaFree(ix->qstats_dev);
    if (ix->qstats_host) cudaFreeHost((void *)ix->qstats_host);
    delete ix;
}

int aix_index23_set_filter(aix_index23 *ix, int mode) {
    if (!ix || mode < 0 || mode > 2) return AIX_ERR_ARG;
    ix->filter_mode = mode;
    return AIX_OK;
}

int aix_index23_filter_stats(const aix_index23 *ix, uint64_t info[6]) {
    if (!ix || !info) return AIX_ERR_ARG;
    info[0] = (uint64_t)ix->bloom_words * 8;
    info[1] = ix->qstats_host ? ix->qstats_host[0] : 0;
    info[2] = ix->qstats_host ? ix->qstats_host[1] : 0;
    info[3] = ix->launches_filter;
    info[4] = ix->launches_direct;
    info[5] = ix->pass_rate < 0 ? ~0ull : (uint64_t)(ix->pass_rate * 1e6);
    return AIX_OK;
}

int aix_index23_info(const aix_index23 *ix, uint64_t info[2]) {
    if (!ix || !info) return AIX_ERR_ARG;
    info[0] = ix->n;
    info[1] = (uint64_t)ix->canonical_only;
    return AIX_OK;
}

int aix_index23_layout(const aix_index23 *ix, uint64_t info[4]) {
    if (!ix || !info) return AIX_ERR_ARG;
    info[0] = ix->frecs_dev ? 4u : (uint64_t)ix->fp_bits;
    info[1] = ix->frecs_dev ? 0 : (ix->fp_bits == 8 ? ix->n : (ix->fp_bits == 4 ? (ix->n + 1) / 2 : 0));
    info[2] = ix->frecs_dev ? ix->frecs_bytes : ix->mphf->layout_bytes;
    info[3] = ix->frecs_dev ? 2 : (ix->mphf->crecs_dev ? 1 : 0);
    return AIX_OK;
}

int aix_tf23_batch_dev(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *recs_dev, uint32_t stride,
                       const uint8_t *lens_dev, uint64_t q, int mode, void *out_dev) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    if (q && (!recs_dev || !out_dev || !stride)) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_tf23(ctx, ix, ctx->stream, recs_dev, stride, lens_dev, q, mode, out_dev);
}

int aix_tf23_single_call_latency(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *recs, uint32_t stride, uint64_t n, uint32_t *tf_out,
                                 double *echo_ns, double *query_ns) {
    if (!ctx || !ix || !recs || !stride || !echo_ns || !query_ns) return AIX_ERR_ARG;
    *echo_ns = *query_ns = 0;
    if (n == 0) return AIX_OK;
    const uint32_t len = stride > kMboxPayload ? kMboxPayload : stride;
    uint32_t v = 0;
    for (int phase = 0; phase < 2; ++phase) {
        if (!mbox_query(ctx, ix, nullptr, recs, len, &v, phase == 0)) return ctx->fail(AIX_ERR_CUDA, "the mailbox kernel is not available");
        const double t0 = AixTrace::now();
        for (uint64_t i = 0; i < n; ++i) {
            if (!mbox_query(ctx, ix, nullptr, recs + i * stride, len, &v, phase == 0))
                return ctx->fail(AIX_ERR_CUDA, "the mailbox kernel stopped answering");
            if (phase == 1 && tf_out) tf_out[i] = v;
        }
        (phase == 0 ? *echo_ns : *query_ns) = (AixTrace::now() - t0) * 1e9 / (double)n;
    }
    return AIX_OK;
}

int aix_tf23_batch(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                   uint64_t q, int mode, void *out) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    if (mode < AIX_Q_TF || mode > AIX_Q_KID) return ctx->fail(AIX_ERR_ARG, "unknown query mode %d", mode);
    if (q == 1 && mode == AIX_Q_TF && recs && out && stride) {  // a single get_tf_value call: the resident mailbox kernel
        uint32_t len = lens ? lens[0] : stride;
        if (len > stride) len = stride;
        if (mbox_query(ctx, ix, nullptr, recs, len, (uint32_t *)out)) return AIX_OK;
    }
    return run_record_batches(ctx, recs, stride, lens, q, out, out_bytes23(mode),
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *l, uint64_t nq, void *o) {
                                  return launch_tf23(ctx, ix, st, r, stride, l, nq, mode, o);
                              });
}


int aix_tf23_probes_dev(aix_ctx *ctx, const aix_mphf *m, uint64_t n_total, int canonical_only, const uint8_t *recs_dev,
                        uint32_t stride, const uint8_t *lens_dev, uint64_t q, uint64_t *probes_dev) {
    if (!ctx || !m) return AIX_ERR_ARG;
    if (q == 0) return AIX_OK;
    if (!recs_dev || !probes_dev || !stride) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (canonical_only)
        tf23_probes_kernel<true><<<aix_grid(q, kQBlock), kQBlock, 0, ctx->stream>>>(m->dev(), n_total, recs_dev, stride, lens_dev, q, (ulonglong2 *)probes_dev);
    else
        tf23_probes_kernel<false><<<aix_grid(q, kQBlock), kQBlock, 0, ctx->stream>>>(m->dev(), n_total, recs_dev, stride, lens_dev, q, (ulonglong2 *)probes_dev);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}


int aix_probes_bucket_dev(aix_ctx *ctx, const uint64_t *probes_dev, uint64_t n_probes, const uint64_t *bounds, int world,
                          uint64_t *counts_dev, uint64_t *send_dev, uint32_t *tag_dev) {
    if (!ctx || !bounds || world < 1 || world > kMaxRanks) return AIX_ERR_ARG;
    if (n_probes >= (1ull << 32)) return ctx->fail(AIX_ERR_ARG, "at most 2^32-1 probes per call");
    if (!counts_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_CUDA(ctx, cudaMemsetAsync(counts_dev, 0, (size_t)world * 8, ctx->stream));
    if (n_probes == 0) return AIX_OK;
    if (!probes_dev || !send_dev || !tag_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    OwnerBounds ob;
    ob.world = world;
    for (int r = 0; r <= world; ++r) ob.b[r] = bounds[r];  // HOST array of world + 1 ascending ids, bounds[0] = 0
    for (int r = world + 1; r <= kMaxRanks; ++r) ob.b[r] = ~0ull;
    void *scratch;
    AIX_TRY(ctx->reserve(SCR_LEN1, 2 * kMaxRanks * 8, &scratch));
    unsigned long long *offs = (unsigned long long *)scratch, *cursor = offs + kMaxRanks;
    const unsigned grid = aix_grid(n_probes, 256);
    probes_count_kernel<<<grid, 256, 0, ctx->stream>>>((const ulonglong2 *)probes_dev, n_probes, ob, (unsigned long long *)counts_dev);
    AIX_LAUNCH_CHECK(ctx);
    probes_offsets_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)counts_dev, world, offs, cursor);
    AIX_LAUNCH_CHECK(ctx);
    probes_scatter_kernel<<<grid, 256, 0, ctx->stream>>>((const ulonglong2 *)probes_dev, n_probes, ob, offs, cursor, (ulonglong2 *)send_dev, tag_dev);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_probe23_dev(aix_ctx *ctx, const aix_index23 *shard, const uint64_t *probes_dev, uint64_t cnt, uint64_t *out_dev) {
    if (!ctx || !shard) return AIX_ERR_ARG;
    if (cnt == 0) return AIX_OK;
    if (!probes_dev || !out_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    probe23_kernel<<<aix_grid(cnt, 256), 256, 0, ctx->stream>>>(shard->recs_dev, shard->n, (const ulonglong2 *)probes_dev, cnt, (unsigned long long *)out_dev);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_get_freq23(aix_ctx *ctx, const aix_index23 *ix, const uint64_t *ukmers, uint64_t q, uint32_t *out) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    Index23Dev id = ix->dev();
    MphfDev md = ix->mphf_dev();
    return run_record_batches(ctx, (const uint8_t *)ukmers, 8, nullptr, q, out, 4,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *, uint64_t nq, void *o) {
                                  if (ix->canonical_only) get_freq23_kernel<true><<<aix_grid(nq, 256), 256, 0, st>>>(id, md, (const uint64_t *)r, nq, (uint32_t *)o);
                                  else get_freq23_kernel<false><<<aix_grid(nq, 256), 256, 0, st>>>(id, md, (const uint64_t *)r, nq, (uint32_t *)o);
                                  AIX_LAUNCH_CHECK(ctx);
                                  return AIX_OK;
                              });
}

static int launch_packed6(aix_ctx *ctx, const aix_index23 *ix, cudaStream_t st, const uint8_t *r, uint64_t nq, void *o) {
    Index23Dev id = ix->dev();
    MphfDev md = ix->mphf_dev();
    if (ix->canonical_only) get_freq23_packed6_kernel<true><<<aix_grid(nq, 256), 256, 0, st>>>(id, md, r, nq, (uint32_t *)o);
    else get_freq23_packed6_kernel<false><<<aix_grid(nq, 256), 256, 0, st>>>(id, md, r, nq, (uint32_t *)o);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_get_freq23_packed(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *packed, uint64_t q, uint32_t *out) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    return run_record_batches(ctx, packed, 6, nullptr, q, out, 4,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *, uint64_t nq, void *o) {
                                  return launch_packed6(ctx, ix, st, r, nq, o);
                              });
}

int aix_get_freq23_packed_dev(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *packed_dev, uint64_t q, uint32_t *out_dev) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    if (q == 0) return AIX_OK;
    if (!packed_dev || !out_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_packed6(ctx, ix, ctx->stream, packed_dev, q, out_dev);
}

int aix_index13_upload(aix_ctx *ctx, const aix_mphf *m, const uint64_t *tf64, aix_index13 **out) {
    if (!ctx || !m || !tf64 || !out) return AIX_ERR_ARG;
    *out = nullptr;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    aix_index13 *ix = new aix_index13();
    ix->mphf = m;
    cudaError_t e = cudaMalloc(&ix->tf_mphf_dev, AIX_TOTAL_13MERS * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ix->tf_direct_dev, AIX_TOTAL_13MERS * 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        aix_index13_destroy(ctx, ix);
        return ctx->fail(AIX_ERR_NOMEM, "index13 upload: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(ix->tf_mphf_dev, tf64, AIX_TOTAL_13MERS * 8, cudaMemcpyHostToDevice, ctx->stream);
    tf13_direct_kernel<<<aix_grid(AIX_TOTAL_13MERS, 256), 256, 0, ctx->stream>>>(m->dev(), ix->tf_mphf_dev, ix->tf_direct_dev);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        aix_index13_destroy(ctx, ix);
        return ctx->fail(AIX_ERR_CUDA, "index13 upload: %s", cudaGetErrorString(e));
    }
    *out = ix;
    return AIX_OK;
}

int aix_index13_tf_direct(aix_ctx *ctx, const aix_index13 *ix, uint64_t *out) {
    if (!ctx || !ix || !out) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_CUDA(ctx, cudaMemcpyAsync(out, ix->tf_direct_dev, AIX_TOTAL_13MERS * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

void aix_index13_destroy(aix_ctx *ctx, aix_index13 *ix) {
    if (!ix) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (ctx && ctx->mbox_owner == ix) mbox_stop(ctx);
    if (ix->tf_mphf_dev) cudaFree(ix->tf_mphf_dev);
    if (ix->tf_direct_dev) cudaFree(ix->tf_direct_dev);
    delete ix;
}

int aix_tf13_batch_dev(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *recs_dev, uint32_t stride,
                       const uint8_t *lens_dev, uint64_t q, int mode, void *out_dev) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    if (q && (!recs_dev || !out_dev || !stride)) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->l2_unpin();
    return launch_tf13(ctx, ix, ctx->stream, recs_dev, stride, lens_dev, q, mode, out_dev);
}

int aix_tf13_batch(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                   uint64_t q, int mode, void *out) {
    if (!ctx || !ix) return AIX_ERR_ARG;
    if (mode < AIX_Q_TF || mode > AIX_Q_BOTH) return ctx->fail(AIX_ERR_ARG, "unknown 13-mer query mode %d", mode);
    size_t ob = mode == AIX_Q_TF ? 4 : (mode == AIX_Q_TOTAL ? 8 : 16);
    if (q == 1 && mode == AIX_Q_TF && recs && out && stride) {
        uint32_t len = lens ? lens[0] : stride;
        if (len > stride) len = stride;
        if (mbox_query(ctx, nullptr, ix, recs, len, (uint32_t *)out)) return AIX_OK;
    }
    return run_record_batches(ctx, recs, stride, lens, q, out, ob,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *l, uint64_t nq, void *o) {
                                  return launch_tf13(ctx, ix, st, r, stride, l, nq, mode, o);
                              });
}

}  // extern "C"
