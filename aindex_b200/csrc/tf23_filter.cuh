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
// k-mers, ~3 % false positives, strings with a non-ACGT byte) are queued per warp -- a 16-bit slot number in shared
// memory -- and go through query23 in batches of 32, all lanes busy, with the tail of the queue drained when the
// warp has seen its last tile.  Same answers as tf23_stream_kernel<AIX_Q_TF, true> (tests/test_gpu_parity.py).
// qstats: {queries seen, queries that passed the filter}, reported by one CTA in 16 (a rate is all the host needs).
template <int kMinBlocks, int kTiles = kStTilesPerWarp>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_filter_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                              uint64_t n_tiles, uint32_t *__restrict__ out,
                                                                              unsigned long long *__restrict__ qstats) {
    __shared__ __align__(128) uint8_t ring[kStWarps][kStStages][kStSlot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    __shared__ uint16_t queue[kStWarps][64];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t ring0 = smem_addr(&ring[wid][0][0]), bar0 = smem_addr(&bars[wid][0]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStStages; ++s) mbar_init(&bars[wid][s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    constexpr int kTilesPerCta = kStWarps * kTiles;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kTilesPerCta + wid;  // this warp's first tile
    if (tile0 >= n_tiles) return;
    const uint64_t left = n_tiles - tile0;
    const uint32_t my_tiles = left >= (uint64_t)kTilesPerCta ? (uint32_t)kTiles : (uint32_t)((left + kStWarps - 1) / kStWarps);
    const uint64_t policy = l2_policy_evict_first();
    constexpr uint32_t kStride = kStWarps * kStTileBytes;
    const uint8_t *src = recs + tile0 * kStTileBytes;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStStages - 1; ++s) {
            if ((uint32_t)s < my_tiles) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(kStTileBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(ring0 + (uint32_t)kStSlot * s), "l"(src + (uint64_t)kStride * s), "r"(kStTileBytes), "r"(bar0 + 8u * s), "l"(policy) : "memory");
            }
        }
    }
    src += (uint64_t)kStride * (kStStages - 1);
    const uint64_t i0 = tile0 * 32u;  // first query of this warp; slot s of the warp = query i0 + (s >> 5) * 256 + (s & 31)
    const uint64_t last_query = n_tiles * 32u - 1;
    uint16_t *wq = queue[wid];
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qn = 0, n_passed = 0;
    // queued slots [0, cnt) through the full lookup, one per lane
    auto drain = [&](uint32_t cnt) {
        if (lane < cnt) {
            const uint32_t s = wq[lane];
            const uint64_t i = i0 + (uint64_t)(s >> 5) * (kStWarps * 32u) + (s & 31u);
            const uint8_t *p = recs + i * 23;
            uint64_t r0, r1, r2;
            if (i != last_query) load_window23(p, r0, r1, r2);
            else {  // the last query of the batch: never read past the buffer
                r0 = r1 = r2 = 0;
#pragma unroll 1
                for (int j = 0; j < 23; ++j) {
                    const uint64_t b = p[j];
                    if (j < 8) r0 |= b << (8 * j);
                    else if (j < 16) r1 |= b << (8 * (j - 8));
                    else r2 |= b << (8 * (j - 16));
                }
            }
            query23<AIX_Q_TF, true>(ix, m, r0, r1, r2, 23u, p, i, out);
        }
        __syncwarp();
    };
    // Software pipeline, one tile deep: iteration `it` encodes tile `it` and PREFETCHES its filter word into L1, then finishes
    // tile `it - 1`, whose word was prefetched an iteration ago -- the word's latency (L2, or HBM for the half of the filter
    // that is not resident) is covered by a tile's worth of encode instead of stalling the warp at the test (ncu of the
    // unpipelined loop: 40 % of all stall samples sat on the first use of the word).  A prefetch rather than an early load:
    // a loaded value carried over the loop edge is copied into the "previous" registers at the top of the next
    // iteration, and that copy waits for the load.
    uint32_t slot = 0, phase = 0;
    uint32_t word_prev = 0, g_prev = 0;
    bool acgt_prev = true;
    for (uint32_t it = 0; it <= my_tiles; ++it) {
        uint32_t word = 0, g = 0;
        bool all_acgt = true;
        if (it < my_tiles) {
            if (lane == 0 && it + (kStStages - 1) < my_tiles) {
                const uint32_t sn = slot == 0 ? kStStages - 1 : slot - 1;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * sn), "r"(kStTileBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(ring0 + (uint32_t)kStSlot * sn), "l"(src), "r"(kStTileBytes), "r"(bar0 + 8u * sn), "l"(policy) : "memory");
            }
            src += kStride;
            {
                uint32_t done;
                do {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                 : "=r"(done) : "r"(bar0 + 8u * slot), "r"(phase) : "memory");
                } while (!done);
            }
            const uint32_t base = lane * 23u;
            const uint32_t a = ring0 + (uint32_t)kStSlot * slot + (base & ~3u), sh = (base & 3u) * 8u;
            uint32_t x0, x1, x2, x3, x4, x5, x6;
            asm volatile("ld.shared.u32 %0, [%7];\nld.shared.u32 %1, [%7+4];\nld.shared.u32 %2, [%7+8];\nld.shared.u32 %3, [%7+12];\n"
                         "ld.shared.u32 %4, [%7+16];\nld.shared.u32 %5, [%7+20];\nld.shared.u32 %6, [%7+24];"
                         : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4), "=r"(x5), "=r"(x6) : "r"(a) : "memory");
            __syncwarp();
            const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
                           y3 = __funnelshift_r(x3, x4, sh), y4 = __funnelshift_r(x4, x5, sh), y5 = __funnelshift_r(x5, x6, sh);
            const uint64_t r0 = ((uint64_t)y1 << 32) | y0, r1 = ((uint64_t)y3 << 32) | y2,
                           r2 = (((uint64_t)y5 << 32) | y4) & 0x00FFFFFFFFFFFFFFULL;
            uint64_t u, r;
            encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
            bloom_word(u <= r ? u : r, ix.bloom_words, word, g);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ix.bloom + word));
            if (++slot == kStStages) { slot = 0; phase ^= 1u; }
        }
        if (it > 0) {  // finish tile it - 1
            const uint2 w_prev = ld_evict_last_u32x2(ix.bloom + word_prev);
            uint32_t mlo, mhi;
            bloom_masks(g_prev, mlo, mhi);
            const bool pass = !acgt_prev || ((w_prev.x & mlo) == mlo && (w_prev.y & mhi) == mhi);
            if (!pass) __stcs(out + i0 + (uint64_t)(it - 1u) * (kStWarps * 32u) + lane, 0u);
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
            if (b) {
                if (pass) wq[qn + __popc(b & lt)] = (uint16_t)((it - 1u) * 32u + lane);
                qn += __popc(b);
                n_passed += __popc(b);
                __syncwarp();
                if (qn >= 32u) {
                    drain(32u);
                    const uint16_t v = wq[32u + lane];
                    __syncwarp();
                    wq[lane] = v;
                    qn -= 32u;
                    __syncwarp();
                }
            }
        }
        word_prev = word;
        g_prev = g;
        acgt_prev = all_acgt;
    }
    if (qn) drain(qn);
    if ((blockIdx.x & 15u) == 0u && lane == 0) {
        atomicAdd(qstats, (unsigned long long)my_tiles * 32ull);
        atomicAdd(qstats + 1, (unsigned long long)n_passed);
    }
}


// The same kernel with TWO queries per lane and iteration: a tile is 64 queries (1472 B = 92 * 16, one bulk copy), lane l
// owns queries 2l and 2l + 1 of it.  What a tile costs apart from its queries -- the refill of the ring by lane 0 (a
// divergent branch every lane waits for), the wait on the slot's barrier, loop control, the queue bookkeeping -- is paid
// once per two queries, the two encodes are independent instruction streams, and a rejected pair is answered with one
// 8-byte store.  Slot numbers are still 16 bits: it * 64 + 2 * lane + j.  The queue is drained from its END (the order of
// the lookups does not matter), so nothing is ever moved inside it.  `out` must be 8-byte aligned.
constexpr uint32_t kF2TileBytes = 64u * 23u;
constexpr int kF2Slot = 1488;  // the seventh word of lane 31's second query ends at byte 1476
template <int kMinBlocks, int kTiles>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_filter2_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                               uint64_t n_tiles, uint32_t *__restrict__ out,
                                                                               unsigned long long *__restrict__ qstats) {
    __shared__ __align__(128) uint8_t ring[kStWarps][kStStages][kF2Slot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    __shared__ uint16_t queue[kStWarps][96];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t ring0 = smem_addr(&ring[wid][0][0]), bar0 = smem_addr(&bars[wid][0]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStStages; ++s) mbar_init(&bars[wid][s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    constexpr int kTilesPerCta = kStWarps * kTiles;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kTilesPerCta + wid;  // this warp's first tile
    if (tile0 >= n_tiles) return;
    const uint64_t left = n_tiles - tile0;
    const uint32_t my_tiles = left >= (uint64_t)kTilesPerCta ? (uint32_t)kTiles : (uint32_t)((left + kStWarps - 1) / kStWarps);
    const uint64_t policy = l2_policy_evict_first();
    constexpr uint32_t kStride = kStWarps * kF2TileBytes;
    const uint8_t *src = recs + tile0 * kF2TileBytes;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStStages - 1; ++s) {
            if ((uint32_t)s < my_tiles) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(kF2TileBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(ring0 + (uint32_t)kF2Slot * s), "l"(src + (uint64_t)kStride * s), "r"(kF2TileBytes), "r"(bar0 + 8u * s), "l"(policy) : "memory");
            }
        }
    }
    src += (uint64_t)kStride * (kStStages - 1);
    const uint64_t i0 = tile0 * 64u;  // first query of this warp; slot s of the warp = query i0 + (s >> 6) * 512 + (s & 63)
    const uint64_t last_query = n_tiles * 64u - 1;
    uint16_t *wq = queue[wid];
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qn = 0, n_passed = 0;
    // the last `cnt` queued slots through the full lookup, one per lane
    auto drain = [&](uint32_t cnt) {
        if (lane < cnt) {
            const uint32_t s = wq[qn - cnt + lane];
            const uint64_t i = i0 + (uint64_t)(s >> 6) * (kStWarps * 64u) + (s & 63u);
            const uint8_t *p = recs + i * 23;
            uint64_t r0, r1, r2;
            if (i != last_query) load_window23(p, r0, r1, r2);
            else {  // the last query of the batch: never read past the buffer
                r0 = r1 = r2 = 0;
#pragma unroll 1
                for (int j = 0; j < 23; ++j) {
                    const uint64_t b = p[j];
                    if (j < 8) r0 |= b << (8 * j);
                    else if (j < 16) r1 |= b << (8 * (j - 8));
                    else r2 |= b << (8 * (j - 16));
                }
            }
            query23<AIX_Q_TF, true>(ix, m, r0, r1, r2, 23u, p, i, out);
        }
        __syncwarp();
        qn -= cnt;
    };
    // one-tile software pipeline as in tf23_filter_kernel: iteration `it` encodes the two queries of tile `it` and prefetches
    // their filter words, then tests the two queries of tile `it - 1`
    uint32_t slot = 0, phase = 0;
    uint32_t word_prev[2] = {0, 0}, g_prev[2] = {0, 0};
    bool acgt_prev[2] = {true, true};
    for (uint32_t it = 0; it <= my_tiles; ++it) {
        uint32_t word[2] = {0, 0}, g[2] = {0, 0};
        bool all_acgt[2] = {true, true};
        if (it < my_tiles) {
            if (lane == 0 && it + (kStStages - 1) < my_tiles) {
                const uint32_t sn = slot == 0 ? kStStages - 1 : slot - 1;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * sn), "r"(kF2TileBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(ring0 + (uint32_t)kF2Slot * sn), "l"(src), "r"(kF2TileBytes), "r"(bar0 + 8u * sn), "l"(policy) : "memory");
            }
            src += kStride;
            {
                uint32_t done;
                do {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                 : "=r"(done) : "r"(bar0 + 8u * slot), "r"(phase) : "memory");
                } while (!done);
            }
            uint32_t x[2][7];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t base = lane * 46u + 23u * j;
                const uint32_t a = ring0 + (uint32_t)kF2Slot * slot + (base & ~3u);
                asm volatile("ld.shared.u32 %0, [%7];\nld.shared.u32 %1, [%7+4];\nld.shared.u32 %2, [%7+8];\nld.shared.u32 %3, [%7+12];\n"
                             "ld.shared.u32 %4, [%7+16];\nld.shared.u32 %5, [%7+20];\nld.shared.u32 %6, [%7+24];"
                             : "=r"(x[j][0]), "=r"(x[j][1]), "=r"(x[j][2]), "=r"(x[j][3]), "=r"(x[j][4]), "=r"(x[j][5]), "=r"(x[j][6])
                             : "r"(a) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t sh = ((lane * 46u + 23u * j) & 3u) * 8u;
                const uint32_t y0 = __funnelshift_r(x[j][0], x[j][1], sh), y1 = __funnelshift_r(x[j][1], x[j][2], sh),
                               y2 = __funnelshift_r(x[j][2], x[j][3], sh), y3 = __funnelshift_r(x[j][3], x[j][4], sh),
                               y4 = __funnelshift_r(x[j][4], x[j][5], sh), y5 = __funnelshift_r(x[j][5], x[j][6], sh);
                const uint64_t r0 = ((uint64_t)y1 << 32) | y0, r1 = ((uint64_t)y3 << 32) | y2,
                               r2 = (((uint64_t)y5 << 32) | y4) & 0x00FFFFFFFFFFFFFFULL;
                uint64_t u, r;
                encode_validate23_rc(r0, r1, r2, all_acgt[j], u, r);
                bloom_word(u <= r ? u : r, ix.bloom_words, word[j], g[j]);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ix.bloom + word[j]));
            }
            if (++slot == kStStages) { slot = 0; phase ^= 1u; }
        }
        if (it > 0) {  // finish tile it - 1
            bool pass[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint2 w_prev = ld_evict_last_u32x2(ix.bloom + word_prev[j]);
                uint32_t mlo, mhi;
                bloom_masks(g_prev[j], mlo, mhi);
                pass[j] = !acgt_prev[j] || ((~w_prev.x & mlo) | (~w_prev.y & mhi)) == 0u;  // every bit set
            }
            uint32_t *o = out + i0 + (uint64_t)(it - 1u) * (kStWarps * 64u) + 2u * lane;
            if (!pass[0] && !pass[1]) __stcs(reinterpret_cast<uint2 *>(o), make_uint2(0u, 0u));
            else {
                if (!pass[0]) __stcs(o, 0u);
                if (!pass[1]) __stcs(o + 1, 0u);
            }
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, pass[0]), b1 = __ballot_sync(0xFFFFFFFFu, pass[1]);
            if (b0 | b1) {
                const uint32_t n0 = __popc(b0), s = (it - 1u) * 64u + 2u * lane;
                if (pass[0]) wq[qn + __popc(b0 & lt)] = (uint16_t)s;
                if (pass[1]) wq[qn + n0 + __popc(b1 & lt)] = (uint16_t)(s + 1u);
                qn += n0 + __popc(b1);
                n_passed += n0 + __popc(b1);
                __syncwarp();
                while (qn >= 32u) drain(32u);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            word_prev[j] = word[j];
            g_prev[j] = g[j];
            acgt_prev[j] = all_acgt[j];
        }
    }
    if (qn) drain(qn);
    if ((blockIdx.x & 15u) == 0u && lane == 0) {
        atomicAdd(qstats, (unsigned long long)my_tiles * 64ull);
        atomicAdd(qstats + 1, (unsigned long long)n_passed);
    }
}


// The filter kernel with the filter word LOADED a tile ahead instead of prefetched.  ncu of tf23_filter_kernel: 39 % of
// all stall samples sit on the first use of the word although it was prefetched into L1 an iteration earlier -- the L1 hit
// rate is 8 %: 32 warps x 32 lanes have 1024 prefetched lines in flight per SM, as many as L1 has lines, and most are gone
// when their load arrives (the prefetch still turns the load's DRAM latency into an L2 hit).  A register cannot be evicted:
// the loop is unrolled by two with the words of even and odd tiles in their own registers (A / B), so that no value is
// copied across the loop edge (a copy would wait for the load):
//     front(0, A);  { front(it, B); back(it - 1, A); front(it + 1, A); back(it, B); } ...  back(last)
// front(t) = wait for tile t, encode + validate its queries, issue the load of the filter word;
// back(t)  = test the word, answer 0 or queue the query; the queue (96 slots) is drained from its end, once per two tiles.
// kWindow: the filter is covered by an access-policy window of the launch (persisting L2 lines, tf_query.cu) -- the word is
// then read with a plain load, whose L2 policy is the window's, instead of the evict_last hint.
template <int kMinBlocks, int kTiles, bool kWindow = false>
__global__ void __launch_bounds__(kStWarps * 32, kMinBlocks) tf23_filter3_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ recs,
                                                                               uint64_t n_tiles, uint32_t *__restrict__ out,
                                                                               unsigned long long *__restrict__ qstats) {
    __shared__ __align__(128) uint8_t ring[kStWarps][kStStages][kStSlot];
    __shared__ __align__(8) uint64_t bars[kStWarps][kStStages];
    __shared__ uint16_t queue[kStWarps][96];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t ring0 = smem_addr(&ring[wid][0][0]), bar0 = smem_addr(&bars[wid][0]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStStages; ++s) mbar_init(&bars[wid][s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    constexpr int kTilesPerCta = kStWarps * kTiles;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kTilesPerCta + wid;  // this warp's first tile
    if (tile0 >= n_tiles) return;
    const uint64_t left = n_tiles - tile0;
    const uint32_t my_tiles = left >= (uint64_t)kTilesPerCta ? (uint32_t)kTiles : (uint32_t)((left + kStWarps - 1) / kStWarps);
    const uint64_t policy = l2_policy_evict_first();
    constexpr uint32_t kStride = kStWarps * kStTileBytes;
    const uint8_t *src = recs + tile0 * kStTileBytes;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStStages - 1; ++s) {
            if ((uint32_t)s < my_tiles) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(kStTileBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(ring0 + (uint32_t)kStSlot * s), "l"(src + (uint64_t)kStride * s), "r"(kStTileBytes), "r"(bar0 + 8u * s), "l"(policy) : "memory");
            }
        }
    }
    src += (uint64_t)kStride * (kStStages - 1);
    const uint64_t i0 = tile0 * 32u;  // first query of this warp; slot s of the warp = query i0 + (s >> 5) * 256 + (s & 31)
    const uint64_t last_query = n_tiles * 32u - 1;
    uint16_t *wq = queue[wid];
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qn = 0, n_passed = 0;
    // the last `cnt` queued slots through the full lookup, one per lane
    auto drain = [&](uint32_t cnt) {
        if (lane < cnt) {
            const uint32_t s = wq[qn - cnt + lane];
            const uint64_t i = i0 + (uint64_t)(s >> 5) * (kStWarps * 32u) + (s & 31u);
            const uint8_t *p = recs + i * 23;
            uint64_t r0, r1, r2;
            if (i != last_query) load_window23(p, r0, r1, r2);
            else {  // the last query of the batch: never read past the buffer
                r0 = r1 = r2 = 0;
#pragma unroll 1
                for (int j = 0; j < 23; ++j) {
                    const uint64_t b = p[j];
                    if (j < 8) r0 |= b << (8 * j);
                    else if (j < 16) r1 |= b << (8 * (j - 8));
                    else r2 |= b << (8 * (j - 16));
                }
            }
            query23<AIX_Q_TF, true>(ix, m, r0, r1, r2, 23u, p, i, out);
        }
        __syncwarp();
        qn -= cnt;
    };
    uint32_t slot = 0, phase = 0;
    auto front = [&](uint32_t it, uint2 &w, uint32_t &g, bool &all_acgt) {
        if (lane == 0 && it + (kStStages - 1) < my_tiles) {
            const uint32_t sn = slot == 0 ? kStStages - 1 : slot - 1;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * sn), "r"(kStTileBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(ring0 + (uint32_t)kStSlot * sn), "l"(src), "r"(kStTileBytes), "r"(bar0 + 8u * sn), "l"(policy) : "memory");
        }
        src += kStride;
        {
            uint32_t done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar0 + 8u * slot), "r"(phase) : "memory");
            } while (!done);
        }
        const uint32_t base = lane * 23u;
        const uint32_t a = ring0 + (uint32_t)kStSlot * slot + (base & ~3u), sh = (base & 3u) * 8u;
        uint32_t x0, x1, x2, x3, x4, x5, x6;
        asm volatile("ld.shared.u32 %0, [%7];\nld.shared.u32 %1, [%7+4];\nld.shared.u32 %2, [%7+8];\nld.shared.u32 %3, [%7+12];\n"
                     "ld.shared.u32 %4, [%7+16];\nld.shared.u32 %5, [%7+20];\nld.shared.u32 %6, [%7+24];"
                     : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4), "=r"(x5), "=r"(x6) : "r"(a) : "memory");
        __syncwarp();
        const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
                       y3 = __funnelshift_r(x3, x4, sh), y4 = __funnelshift_r(x4, x5, sh), y5 = __funnelshift_r(x5, x6, sh);
        const uint64_t r0 = ((uint64_t)y1 << 32) | y0, r1 = ((uint64_t)y3 << 32) | y2,
                       r2 = (((uint64_t)y5 << 32) | y4) & 0x00FFFFFFFFFFFFFFULL;
        uint64_t u, r;
        encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
        uint32_t word;
        bloom_word(u <= r ? u : r, ix.bloom_words, word, g);
        w = kWindow ? __ldg(ix.bloom + word) : ld_evict_last_u32x2(ix.bloom + word);
        if (++slot == kStStages) { slot = 0; phase ^= 1u; }
    };
    auto back = [&](uint32_t it, const uint2 &w, uint32_t g, bool all_acgt) {
        uint32_t mlo, mhi;
        bloom_masks(g, mlo, mhi);
        const bool pass = !all_acgt || ((~w.x & mlo) | (~w.y & mhi)) == 0u;  // every bit set
        if (!pass) __stcs(out + i0 + (uint64_t)it * (kStWarps * 32u) + lane, 0u);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
        if (b) {
            if (pass) wq[qn + __popc(b & lt)] = (uint16_t)(it * 32u + lane);
            qn += __popc(b);
            n_passed += __popc(b);
        }
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
        while (qn >= 32u) drain(32u);
    }
    if (it < my_tiles) {
        front(it, wB, gB, okB);
        back(it - 1u, wA, gA, okA);
        back(it, wB, gB, okB);
    } else {
        back(it - 1u, wA, gA, okA);
    }
    __syncwarp();
    while (qn) drain(qn < 32u ? qn : 32u);
    if ((blockIdx.x & 15u) == 0u && lane == 0) {
        atomicAdd(qstats, (unsigned long long)my_tiles * 32ull);
        atomicAdd(qstats + 1, (unsigned long long)n_passed);
    }
}

}  // namespace aix
