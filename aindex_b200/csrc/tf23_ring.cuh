// tf23_ring.cuh -- the per-warp TMA ring of the streaming 23-mer query kernels (tf_query.cu, tf23_filter.cuh):
// constants and the mbarrier / bulk-copy helpers.
#pragma once
#include "aix_internal.cuh"
#include "query23.cuh"

namespace aix {

// ---- K3 streaming form: persistent CTAs, one TMA ring per warp ------------------------------------
// The fixed kernel above pays three dependent round trips per query (query bytes from HBM, MPHF
// records, fingerprint / index record) and a CTA barrier between the first two.  Here every warp
// owns a 3-slot ring of 32-query tiles (736 B) in shared memory that lane 0 fills with
// cp.async.bulk (TMA, UBLKCP in SASS) two tiles ahead, completion signalled on one mbarrier per
// slot: the query bytes are already on chip when a warp starts a tile and there is no CTA-wide
// barrier.
constexpr int kStWarps = 8;
#ifndef AIX_ST_STAGES
#define AIX_ST_STAGES 3
#endif
#ifndef AIX_ST_TILES
#define AIX_ST_TILES 16
#endif
constexpr int kStStages = AIX_ST_STAGES;  // ring depth and tiles per warp are compile-time knobs (profiles/r01_tf23_sweep.txt)
constexpr uint32_t kStTileBytes = 32u * 23u;  // 736 = 46 * 16: legal bulk-copy size, slots stay 16-byte aligned
constexpr int kStSlot = 768;                  // the seventh word of lane 31 ends at byte 740

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy) : "memory");
}

// One warp's ring: kStStages slots of kSlot bytes in shared memory, each filled by lane 0 with ONE bulk copy of kTileBytes
// (a multiple of 16), kStStages - 1 tiles ahead of the tile the warp works on; one mbarrier per slot.  The warp's tiles are
// kStWarps tiles apart in the batch (warp w of a CTA takes tiles w, w + 8, ...).
template <uint32_t kTileBytes, uint32_t kSlot>
struct WarpRing {
    static constexpr uint32_t kStride = kStWarps * kTileBytes;  // bytes between a warp's consecutive tiles
    uint32_t ring0, bar0;  // shared-memory addresses of this warp's slots and barriers
    const uint8_t *src;    // next tile to fetch
    uint64_t policy;
    uint32_t slot, phase, my_tiles;

    // before the warp knows whether it has work: barriers initialised by lane 0, visible to the async proxy
    __device__ __forceinline__ void init(const void *slots, uint64_t *bars, unsigned lane) {
        ring0 = smem_addr(slots);
        bar0 = smem_addr(bars);
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStStages; ++s) mbar_init(bars + s, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    __device__ __forceinline__ void fill(uint32_t s, const uint8_t *from) const {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(kTileBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(ring0 + kSlot * s), "l"(from), "r"(kTileBytes), "r"(bar0 + 8u * s), "l"(policy) : "memory");
    }
    // the first kStStages - 1 tiles of the warp's `tiles` tiles, the first of them at `first`
    __device__ __forceinline__ void start(const uint8_t *first, uint32_t tiles, unsigned lane) {
        my_tiles = tiles;
        policy = l2_policy_evict_first();
        slot = phase = 0;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStStages - 1; ++s)
                if ((uint32_t)s < my_tiles) fill((uint32_t)s, first + (uint64_t)kStride * s);
        }
        src = first + (uint64_t)kStride * (kStStages - 1);
    }
    // tile `it`: refill the slot every lane finished reading in the previous tile (the __syncwarp after its loads), then
    // wait for this tile's bytes; returns the shared-memory address of the tile
    __device__ __forceinline__ uint32_t acquire(uint32_t it, unsigned lane) {
        if (lane == 0 && it + (kStStages - 1) < my_tiles) fill(slot == 0 ? kStStages - 1 : slot - 1, src);
        src += kStride;
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar0 + 8u * slot), "r"(phase) : "memory");
        } while (!done);
        return ring0 + kSlot * slot;
    }
    __device__ __forceinline__ void advance() {
        if (++slot == kStStages) { slot = 0; phase ^= 1u; }
    }
};

// the tiles of warp `wid` of this CTA when a CTA owns kStWarps * kTiles consecutive tiles: its first tile and how many
template <int kTiles>
__device__ __forceinline__ uint32_t warp_tiles(uint64_t n_tiles, unsigned wid, uint64_t &tile0) {
    constexpr int kTilesPerCta = kStWarps * kTiles;
    tile0 = (uint64_t)blockIdx.x * kTilesPerCta + wid;
    if (tile0 >= n_tiles) return 0u;
    const uint64_t left = n_tiles - tile0;  // tiles from tile0 on
    return left >= (uint64_t)kTilesPerCta ? (uint32_t)kTiles : (uint32_t)((left + kStWarps - 1) / kStWarps);
}

// the 23 bytes at byte offset `off` of a tile in shared memory: seven aligned words ...
__device__ __forceinline__ void lds_words7(uint32_t tile, uint32_t off, uint32_t (&x)[7]) {
    asm volatile("ld.shared.u32 %0, [%7];\nld.shared.u32 %1, [%7+4];\nld.shared.u32 %2, [%7+8];\nld.shared.u32 %3, [%7+12];\n"
                 "ld.shared.u32 %4, [%7+16];\nld.shared.u32 %5, [%7+20];\nld.shared.u32 %6, [%7+24];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]) : "r"(tile + (off & ~3u)) : "memory");
}
// ... funnel-shifted into three little-endian words (zero padded past byte 22)
__device__ __forceinline__ void words23(const uint32_t (&x)[7], uint32_t off, uint64_t &r0, uint64_t &r1, uint64_t &r2) {
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t y0 = __funnelshift_r(x[0], x[1], sh), y1 = __funnelshift_r(x[1], x[2], sh), y2 = __funnelshift_r(x[2], x[3], sh),
                   y3 = __funnelshift_r(x[3], x[4], sh), y4 = __funnelshift_r(x[4], x[5], sh), y5 = __funnelshift_r(x[5], x[6], sh);
    r0 = ((uint64_t)y1 << 32) | y0;
    r1 = ((uint64_t)y3 << 32) | y2;
    r2 = (((uint64_t)y5 << 32) | y4) & 0x00FFFFFFFFFFFFFFULL;
}

// query i of a batch of fixed 23-byte records straight from global memory; the LAST query of the buffer is read byte by byte
// (load_window23 may touch the word behind the window)
__device__ __forceinline__ void load_query23(const uint8_t *recs, uint64_t i, uint64_t last_query, uint64_t &r0, uint64_t &r1, uint64_t &r2) {
    const uint8_t *p = recs + i * 23;
    if (i != last_query) {
        load_window23(p, r0, r1, r2);
        return;
    }
    r0 = r1 = r2 = 0;
#pragma unroll 1
    for (int j = 0; j < 23; ++j) {
        const uint64_t b = p[j];
        if (j < 8) r0 |= b << (8 * j);
        else if (j < 16) r1 |= b << (8 * (j - 8));
        else r2 |= b << (8 * (j - 16));
    }
}

// tiles per warp and CTA: a CTA owns kStWarps * kStTilesPerWarp consecutive tiles (4096 queries), warp w takes
// tiles w, w + 8, ...  Small enough that the hardware scheduler evens out SM speed differences (one wave of
// resident CTAs per launch left a quarter of the warp slots idle at the tail), long enough that the two
// exposed loads of the ring prologue are amortised.
constexpr int kStTilesPerWarp = AIX_ST_TILES;
constexpr int kStTilesPerCta = kStWarps * kStTilesPerWarp;

}  // namespace aix
