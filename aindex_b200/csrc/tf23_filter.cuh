// tf23_filter.cuh -- get_tf_values on a canonical-only index through the front filter (Index23Dev::bloom): the kernels of
// miss-dominated batches.  Reference semantics: AindexWrapper::get_tf_values python_wrapper.cpp:653-664 over
// get_tf_value_23mer :610-627; the filter only removes lookups whose answer is known to be 0.
#pragma once
#include "tf23_ring.cuh"

namespace aix {

// ---- front filter (Index23Dev::bloom) ---------------------------------------------------------------------------
// every stored (canonical) k-mer sets its four bits
__global__ void __launch_bounds__(256) bloom_build_kernel(const uint4 *__restrict__ recs, uint64_t n, unsigned long long *__restrict__ bloom,
                                                        uint32_t n_words) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = recs[i];
    uint32_t word, mlo, mhi;
    bloom_slot(((uint64_t)r.y << 32) | r.x, n_words, word, mlo, mhi);
    atomicOr(bloom + word, ((unsigned long long)mhi << 32) | mlo);
}

// get_tf_values on a canonical-only index for batches in which most queries are absent: the ring of tf23_stream_kernel,
// but a query is first tested against the front filter.  Rejected queries are answered 0 on the spot; the others (stored
// k-mers, ~4 % false positives, strings with a non-ACGT byte) are queued per warp -- a 16-bit slot number in shared
// memory -- and go through query23 in batches of 32, all lanes busy, with the rest of the queue drained when the
// warp has seen its last tile.  Same answers as tf23_stream_kernel<AIX_Q_TF, true> (tests/test_gpu_parity.py).
// qstats: {queries seen, queries that passed the filter}, reported by one CTA in 16 (a rate is all the host needs).
//
// Three kernels, one of them the default (tf_query.cu: launch23_mode; numbers in profiles/r02_filter_sweep.txt):
//   tf23_filter3_kernel   the filter word is LOADED a tile ahead, loop unrolled by two          99 - 103 G q/s
//   tf23_filter_kernel    the word is prefetched into L1 a tile ahead                           91 G q/s
//   tf23_filter2_kernel   two queries per lane and iteration (fewest instructions per query)    84 - 88 G q/s

// what the three kernels share: the per-warp queue of slot numbers and the lookup of queued queries
template <uint32_t kTileQueries>
struct FilterQueue {
    uint16_t *wq;         // this warp's slots in shared memory
    uint32_t n, passed;   // queued now / ever
    uint64_t i0;          // first query of the warp: slot s = query i0 + (s / kTileQueries) * kStWarps * kTileQueries + s % kTileQueries
    uint64_t last_query;  // of the batch (its bytes are read with care)

    __device__ __forceinline__ uint64_t query_of(uint32_t s) const {
        return i0 + (uint64_t)(s / kTileQueries) * (kStWarps * kTileQueries) + (s % kTileQueries);
    }
    // the lanes whose query passed append its slot number; b = ballot of `pass`
    __device__ __forceinline__ void push(bool pass, uint32_t b, uint32_t slot_no, unsigned lane) {
        if (pass) wq[n + __popc(b & ((1u << lane) - 1u))] = (uint16_t)slot_no;
        n += __popc(b);
        passed += __popc(b);
    }
    // the last `cnt` (<= 32) queued slots through the full lookup, one per lane (the order of the lookups does not matter:
    // taking them from the end means nothing is ever moved inside the queue)
    __device__ __forceinline__ void drain(uint32_t cnt, const Index23Dev &ix, const MphfDev &m, const uint8_t *recs, uint32_t *out,
                                          unsigned lane) {
        if (lane < cnt) {
            const uint64_t i = query_of(wq[n - cnt + lane]);
            uint64_t r0, r1, r2;
            load_query23(recs, i, last_query, r0, r1, r2);
            query23<AIX_Q_TF, true>(ix, m, r0, r1, r2, 23u, recs + i * 23, i, out);
        }
        __syncwarp();
        n -= cnt;
    }
};

// encode + validate one query of a tile and find its filter word (the first half of every kernel's iteration)
__device__ __forceinline__ void filter_front(const Index23Dev &ix, const uint32_t (&x)[7], uint32_t off, uint32_t &word, uint32_t &g,
                                             bool &all_acgt) {
    uint64_t r0, r1, r2, u, r;
    words23(x, off, r0, r1, r2);
    encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
    bloom_word(u <= r ? u : r, ix.bloom_words, word, g);
}
// does the query go on to the lookup?  (a string with a non-ACGT byte always does: the reference hashes its raw bytes)
__device__ __forceinline__ bool filter_pass(uint2 w, uint32_t g, bool all_acgt) {
    uint32_t mlo, mhi;
    bloom_masks(g, mlo, mhi);
    return !all_acgt || ((~w.x & mlo) | (~w.y & mhi)) == 0u;  // every bit set
}

// ---- the default: word loaded a tile ahead --------------------------------------------------------------------
// ncu of tf23_filter_kernel (below): 39 % of all stall samples sit on the first use of the filter word although it was
// prefetched into L1 an iteration earlier -- the L1 hit rate is 8 %: 32 warps x 32 lanes have 1024 prefetched lines in
// flight per SM, as many as L1 has lines, and most are gone when their load arrives (the prefetch still turns the load's
// DRAM latency into an L2 hit).  A register cannot be evicted: the loop is unrolled by two with the words of even and odd
// tiles in their own registers (A / B), so that no value is copied across the loop edge (a copy would wait for the load):
//     front(0, A);  { front(it, B); back(it - 1, A); front(it + 1, A); back(it, B); } ...  back(last)
// front(t) = wait for tile t, encode + validate its queries, issue the load of the filter word;
// back(t)  = test the word, answer 0 or queue the query; the queue (96 slots) is drained once per two tiles.
// kWindow: the filter is covered by an access-policy window of the launch (persisting L2 lines, tf_query.cu) -- the word is
// then read with a plain load, whose L2 policy is the window's, instead of the evict_last hint.
template <int kMinBlocks, int kTiles, bool kWindow = false>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_filter3_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                               uint64_t n_tiles, uint32_t *__restrict__ out,
                                                                               unsigned long long *__restrict__ qstats) {
    __shared__ __align__(128) uint8_t slots[kStWarps][kStStages][kStSlot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    __shared__ uint16_t queue[kStWarps][96];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WarpRing<kStTileBytes, kStSlot> ring;
    ring.init(&slots[wid][0][0], &bars[wid][0], lane);
    uint64_t tile0;
    const uint32_t my_tiles = warp_tiles<kTiles>(n_tiles, wid, tile0);
    if (my_tiles == 0) return;
    ring.start(recs + tile0 * kStTileBytes, my_tiles, lane);
    FilterQueue<32> q = {queue[wid], 0u, 0u, tile0 * 32u, n_tiles * 32u - 1};
    auto front = [&](uint32_t it, uint2 &w, uint32_t &g, bool &all_acgt) {
        const uint32_t tile = ring.acquire(it, lane);
        uint32_t x[7], word;
        lds_words7(tile, lane * 23u, x);
        __syncwarp();
        filter_front(ix, x, lane * 23u, word, g, all_acgt);
        w = kWindow ? __ldg(ix.bloom + word) : ld_evict_last_u32x2(ix.bloom + word);
        ring.advance();
    };
    auto back = [&](uint32_t it, const uint2 &w, uint32_t g, bool all_acgt) {
        const bool pass = filter_pass(w, g, all_acgt);
        if (!pass) __stcs(out + q.i0 + (uint64_t)it * (kStWarps * 32u) + lane, 0u);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
        if (b) q.push(pass, b, it * 32u + lane, lane);
    };
    uint2 wA, wB;
    uint32_t gA, gB;
    bool okA, okB;
    front(0u, wA, gA, okA);
    uint32_t it = 1;
    for (; it + 1 < my_tiles; it += 2) {
        front(it, wB, gB, okB);
        back(it - 1u, wA, gA, okA);
        front(it + 1u, wA, gA, okA);
        back(it, wB, gB, okB);
        __syncwarp();
        while (q.n >= 32u) q.drain(32u, ix, m, recs, out, lane);
    }
    if (it < my_tiles) {
        front(it, wB, gB, okB);
        back(it - 1u, wA, gA, okA);
        back(it, wB, gB, okB);
    } else {
        back(it - 1u, wA, gA, okA);
    }
    __syncwarp();
    while (q.n) q.drain(q.n < 32u ? q.n : 32u, ix, m, recs, out, lane);
    if ((blockIdx.x & 15u) == 0u && lane == 0) {
        atomicAdd(qstats, (unsigned long long)my_tiles * 32ull);
        atomicAdd(qstats + 1, (unsigned long long)q.passed);
    }
}

// ---- AIX_FILTER_KERNEL=1: word prefetched a tile ahead -----------------------------------------------------------
// Software pipeline, one tile deep: iteration `it` encodes tile `it` and PREFETCHES its filter word into L1, then finishes
// tile `it - 1`, whose word was prefetched an iteration ago.  (Before it: 40 % of all stall samples on the first use of a
// word loaded in the same iteration.)
template <int kMinBlocks, int kTiles = kStTilesPerWarp>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_filter_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                              uint64_t n_tiles, uint32_t *__restrict__ out,
                                                                              unsigned long long *__restrict__ qstats) {
    __shared__ __align__(128) uint8_t slots[kStWarps][kStStages][kStSlot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    __shared__ uint16_t queue[kStWarps][64];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WarpRing<kStTileBytes, kStSlot> ring;
    ring.init(&slots[wid][0][0], &bars[wid][0], lane);
    uint64_t tile0;
    const uint32_t my_tiles = warp_tiles<kTiles>(n_tiles, wid, tile0);
    if (my_tiles == 0) return;
    ring.start(recs + tile0 * kStTileBytes, my_tiles, lane);
    FilterQueue<32> q = {queue[wid], 0u, 0u, tile0 * 32u, n_tiles * 32u - 1};
    uint32_t word_prev = 0, g_prev = 0;
    bool acgt_prev = true;
    for (uint32_t it = 0; it <= my_tiles; ++it) {
        uint32_t word = 0, g = 0;
        bool all_acgt = true;
        if (it < my_tiles) {
            const uint32_t tile = ring.acquire(it, lane);
            uint32_t x[7];
            lds_words7(tile, lane * 23u, x);
            __syncwarp();
            filter_front(ix, x, lane * 23u, word, g, all_acgt);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ix.bloom + word));
            ring.advance();
        }
        if (it > 0) {  // finish tile it - 1
            const bool pass = filter_pass(ld_evict_last_u32x2(ix.bloom + word_prev), g_prev, acgt_prev);
            if (!pass) __stcs(out + q.i0 + (uint64_t)(it - 1u) * (kStWarps * 32u) + lane, 0u);
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
            if (b) {
                q.push(pass, b, (it - 1u) * 32u + lane, lane);
                __syncwarp();
                if (q.n >= 32u) q.drain(32u, ix, m, recs, out, lane);
            }
        }
        word_prev = word;
        g_prev = g;
        acgt_prev = all_acgt;
    }
    if (q.n) q.drain(q.n, ix, m, recs, out, lane);
    if ((blockIdx.x & 15u) == 0u && lane == 0) {
        atomicAdd(qstats, (unsigned long long)my_tiles * 32ull);
        atomicAdd(qstats + 1, (unsigned long long)q.passed);
    }
}

// ---- AIX_FILTER_KERNEL=2: two queries per lane and iteration ---------------------------------------------------------
// A tile is 64 queries (1472 B = 92 * 16, one bulk copy), lane l owns queries 2l and 2l + 1 of it.  What a tile costs apart
// from its queries -- the refill of the ring by lane 0 (a divergent branch every lane waits for), the wait on the slot's
// barrier, loop control, the queue bookkeeping -- is paid once per two queries, the two encodes are independent
// instruction streams, and a rejected pair is answered with one 8-byte store (`out` must be 8-byte aligned).  17 % fewer
// instructions per query than one query per lane -- and slower: 72 registers (3 resident CTAs), or 60 with ptxas held to 4.
constexpr uint32_t kF2TileBytes = 64u * 23u;
constexpr int kF2Slot = 1488;  // the seventh word of lane 31's second query ends at byte 1476
template <int kMinBlocks, int kTiles>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_filter2_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                               uint64_t n_tiles, uint32_t *__restrict__ out,
                                                                               unsigned long long *__restrict__ qstats) {
    __shared__ __align__(128) uint8_t slots[kStWarps][kStStages][kF2Slot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    __shared__ uint16_t queue[kStWarps][96];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    WarpRing<kF2TileBytes, kF2Slot> ring;
    ring.init(&slots[wid][0][0], &bars[wid][0], lane);
    uint64_t tile0;
    const uint32_t my_tiles = warp_tiles<kTiles>(n_tiles, wid, tile0);
    if (my_tiles == 0) return;
    ring.start(recs + tile0 * kF2TileBytes, my_tiles, lane);
    FilterQueue<64> q = {queue[wid], 0u, 0u, tile0 * 64u, n_tiles * 64u - 1};
    uint32_t word_prev[2] = {0, 0}, g_prev[2] = {0, 0};
    bool acgt_prev[2] = {true, true};
    for (uint32_t it = 0; it <= my_tiles; ++it) {
        uint32_t word[2] = {0, 0}, g[2] = {0, 0};
        bool all_acgt[2] = {true, true};
        if (it < my_tiles) {
            const uint32_t tile = ring.acquire(it, lane);
            uint32_t x[2][7];
#pragma unroll
            for (int j = 0; j < 2; ++j) lds_words7(tile, lane * 46u + 23u * j, x[j]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                filter_front(ix, x[j], lane * 46u + 23u * j, word[j], g[j], all_acgt[j]);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ix.bloom + word[j]));
            }
            ring.advance();
        }
        if (it > 0) {  // finish tile it - 1
            bool pass[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) pass[j] = filter_pass(ld_evict_last_u32x2(ix.bloom + word_prev[j]), g_prev[j], acgt_prev[j]);
            uint32_t *o = out + q.i0 + (uint64_t)(it - 1u) * (kStWarps * 64u) + 2u * lane;
            if (!pass[0] && !pass[1]) __stcs(reinterpret_cast<uint2 *>(o), make_uint2(0u, 0u));
            else {
                if (!pass[0]) __stcs(o, 0u);
                if (!pass[1]) __stcs(o + 1, 0u);
            }
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, pass[0]), b1 = __ballot_sync(0xFFFFFFFFu, pass[1]);
            if (b0 | b1) {
                const uint32_t s = (it - 1u) * 64u + 2u * lane;
                q.push(pass[0], b0, s, lane);
                q.push(pass[1], b1, s + 1u, lane);
                __syncwarp();
                while (q.n >= 32u) q.drain(32u, ix, m, recs, out, lane);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            word_prev[j] = word[j];
            g_prev[j] = g[j];
            acgt_prev[j] = all_acgt[j];
        }
    }
    if (q.n) q.drain(q.n, ix, m, recs, out, lane);
    if ((blockIdx.x & 15u) == 0u && lane == 0) {
        atomicAdd(qstats, (unsigned long long)my_tiles * 64ull);
        atomicAdd(qstats + 1, (unsigned long long)q.passed);
    }
}

}  // namespace aix
