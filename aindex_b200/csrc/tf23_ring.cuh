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

// tiles per warp and CTA: a CTA owns kStWarps * kStTilesPerWarp consecutive tiles (4096 queries), warp w takes
// tiles w, w + 8, ...  Small enough that the hardware scheduler evens out SM speed differences (one wave of
// resident CTAs per launch left a quarter of the warp slots idle at the tail), long enough that the two
// exposed loads of the ring prologue are amortised.
constexpr int kStTilesPerWarp = AIX_ST_TILES;
constexpr int kStTilesPerCta = kStWarps * kStTilesPerWarp;

}  // namespace aix
