// atomic_roofline.cu -- measured denominators for the two random-access bound kernels:
//   (1) RED.ADD.U32 to uniform-random addresses of a table of 16..256 MiB (13-mer counting:
//       256 MiB = the 4^13 histogram; smaller tables = the slices of the multi-pass mode)
//   (2) random 16-byte gathers from a table of 0.8 GB (23-mer lookup: one {checker,tf} record
//       per probe) -- the "HBM random-access roofline" of BASELINE.json's north star.
// Keys / indices are precomputed and streamed (4 B each), so the kernels do nothing but the
// random access.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_roofline atomic_roofline.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void gen_keys(uint32_t *k, uint64_t n, uint32_t mask, uint64_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    k[i] = (uint32_t)x & mask;
}

__global__ void red_kernel(const uint4 *__restrict__ keys, uint64_t n4, uint32_t *__restrict__ table) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    uint4 k = __ldcs(keys + i);
    atomicAdd(table + k.x, 1u);
    atomicAdd(table + k.y, 1u);
    atomicAdd(table + k.z, 1u);
    atomicAdd(table + k.w, 1u);
}

__global__ void gather16_kernel(const uint4 *__restrict__ idx, uint64_t n4, const uint4 *__restrict__ table, uint32_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    uint4 k = __ldcs(idx + i);
    uint4 a = __ldg(table + k.x), b = __ldg(table + k.y), c = __ldg(table + k.z), d = __ldg(table + k.w);
    out[i] = a.x ^ b.y ^ c.z ^ d.w;
}

// (3) the shape of one 23-mer lookup with everything but the memory requests removed: 23 streamed bytes in and 4 bytes
//     out per thread, kRec scattered 16-byte loads from an L2-resident record table (the MPHF records, evict_last) and
//     kByte scattered 1-byte loads from a second L2-resident table (the fingerprint tier).  Indices come from a 3-step
//     integer mix of the streamed bytes.  Its rate is the L1TEX / L2 request ceiling of the lookup kernel:
//     tier layout = (3, 1), fused layout = (3, 0).
template <int kRec, int kByte>
__global__ void __launch_bounds__(256) lookup_shape_kernel(const uint8_t *__restrict__ recs, uint64_t q, const uint4 *__restrict__ table,
                                                         uint32_t rec_mask, const uint8_t *__restrict__ bytes, uint32_t byte_mask,
                                                         uint32_t *__restrict__ out) {
    __shared__ __align__(16) uint32_t tile[(256 * 23 + 16) / 4 + 4];
    const uint64_t q0 = (uint64_t)blockIdx.x * 256;
    const uint4 *src = reinterpret_cast<const uint4 *>(recs + q0 * 23);
    for (int v = threadIdx.x; v < 368; v += 256) reinterpret_cast<uint4 *>(tile)[v] = __ldcs(src + v);  // 256 * 23 = 368 * 16
    __syncthreads();
    const uint64_t i = q0 + threadIdx.x;
    if (i >= q) return;
    const uint32_t w = (threadIdx.x * 23u) >> 2;
    uint32_t x = tile[w] ^ (tile[w + 2] * 0x9E3779B9u) ^ (tile[w + 4] * 0x85EBCA6Bu);
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    uint32_t acc = 0;
#pragma unroll
    for (int r = 0; r < kRec; ++r) {
        x = x * 0x2C1B3C6Du + 0x297A2D39u;
        uint4 v;
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(table + ((x >> 7) & rec_mask)), "l"(p));
        acc ^= v.x + v.w;
    }
#pragma unroll
    for (int r = 0; r < kByte; ++r) {
        x = x * 0x2C1B3C6Du + 0x297A2D39u;
        uint32_t f;
        asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(f) : "l"(bytes + ((x >> 5) & byte_mask)), "l"(p));
        acc += f;
    }
    __stcs(out + i, acc);
}

// the same requests behind the input path of tf23_stream_kernel (csrc/tf_query.cu): every warp owns a 3-slot ring of
// 32-query tiles (736 B) filled by cp.async.bulk two tiles ahead, one mbarrier per slot, no CTA barrier
template <int kRec, int kByte>
__global__ void __launch_bounds__(256) lookup_shape_stream_kernel(const uint8_t *__restrict__ recs, uint64_t n_tiles, const uint4 *__restrict__ table,
                                                                uint32_t rec_mask, const uint8_t *__restrict__ bytes, uint32_t byte_mask,
                                                                uint32_t *__restrict__ out) {
    constexpr int kWarps = 8, kStages = 3, kSlot = 768, kTilesPerWarp = 16;
    constexpr uint32_t kTileBytes = 736;
    __shared__ __align__(128) uint8_t ring[kWarps][kStages][kSlot];
    __shared__ __align__(8) uint64_t bars[kWarps][kStages];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(&ring[wid][0][0]), bar0 = (uint32_t)__cvta_generic_to_shared(&bars[wid][0]);
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint64_t tile0 = (uint64_t)blockIdx.x * (kWarps * kTilesPerWarp) + wid;
    if (tile0 >= n_tiles) return;
    const uint64_t left = n_tiles - tile0;
    const uint32_t my_tiles = left >= (uint64_t)(kWarps * kTilesPerWarp) ? (uint32_t)kTilesPerWarp : (uint32_t)((left + kWarps - 1) / kWarps);
    uint64_t pol_in, pol_keep;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_in));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    constexpr uint32_t kStride = kWarps * kTileBytes;
    const uint8_t *src = recs + tile0 * kTileBytes;
    auto fill = [&](uint32_t s, const uint8_t *from) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(kTileBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(ring0 + (uint32_t)kSlot * s), "l"(from), "r"(kTileBytes), "r"(bar0 + 8u * s), "l"(pol_in) : "memory");
    };
    if (lane == 0)
        for (int s = 0; s < kStages - 1; ++s)
            if ((uint32_t)s < my_tiles) fill(s, src + (uint64_t)kStride * s);
    src += (uint64_t)kStride * (kStages - 1);
    uint64_t i = tile0 * 32u + lane;
    uint32_t slot = 0, phase = 0;
    for (uint32_t it = 0; it < my_tiles; ++it) {
        if (lane == 0 && it + (kStages - 1) < my_tiles) fill(slot == 0 ? kStages - 1 : slot - 1, src);
        src += kStride;
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar0 + 8u * slot), "r"(phase) : "memory");
        } while (!done);
        const uint32_t a = ring0 + (uint32_t)kSlot * slot + ((lane * 23u) & ~3u);
        uint32_t x0, x2, x4;
        asm volatile("ld.shared.u32 %0, [%3];\nld.shared.u32 %1, [%3+8];\nld.shared.u32 %2, [%3+16];" : "=r"(x0), "=r"(x2), "=r"(x4) : "r"(a) : "memory");
        __syncwarp();
        uint32_t x = x0 ^ (x2 * 0x9E3779B9u) ^ (x4 * 0x85EBCA6Bu), acc = 0;
#pragma unroll
        for (int r = 0; r < kRec; ++r) {
            x = x * 0x2C1B3C6Du + 0x297A2D39u;
            uint4 v;
            asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "l"(table + ((x >> 7) & rec_mask)), "l"(pol_keep));
            acc ^= v.x + v.w;
        }
#pragma unroll
        for (int r = 0; r < kByte; ++r) {
            x = x * 0x2C1B3C6Du + 0x297A2D39u;
            uint32_t f;
            asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(f) : "l"(bytes + ((x >> 5) & byte_mask)), "l"(pol_keep));
            acc += f;
        }
        __stcs(out + i, acc);
        i += (uint64_t)kWarps * 32u;
        if (++slot == kStages) { slot = 0; phase ^= 1u; }
    }
}

template <int kRec, int kByte>
static void run_lookup_shape_stream(const uint8_t *recs, uint64_t q, const uint4 *table, uint32_t rec_mask, const uint8_t *bytes,
                                    uint32_t byte_mask, uint32_t *out, cudaEvent_t a, cudaEvent_t b) {
    float best = 1e30f;
    const uint64_t n_tiles = q / 32, grid = (n_tiles + 127) / 128;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        lookup_shape_stream_kernel<kRec, kByte><<<(unsigned)grid, 256>>>(recs, n_tiles, table, rec_mask, bytes, byte_mask, out);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 1 && ms < best) best = ms;
    }
    printf("lookup_shape_stream recs16=%d bytes1=%d record_table_MiB=%u byte_table_MiB=%u ms=%.3f Gqueries/s=%.2f\n", kRec, kByte,
           (unsigned)(((uint64_t)rec_mask + 1) * 16 >> 20), (unsigned)(((uint64_t)byte_mask + 1) >> 20), best, q / (best * 1e-3) / 1e9);
}

template <int kRec, int kByte>
static void run_lookup_shape(const uint8_t *recs, uint64_t q, const uint4 *table, uint32_t rec_mask, const uint8_t *bytes,
                             uint32_t byte_mask, uint32_t *out, cudaEvent_t a, cudaEvent_t b) {
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        lookup_shape_kernel<kRec, kByte><<<(unsigned)(q / 256), 256>>>(recs, q, table, rec_mask, bytes, byte_mask, out);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 1 && ms < best) best = ms;
    }
    printf("lookup_shape recs16=%d bytes1=%d record_table_MiB=%u byte_table_MiB=%u ms=%.3f Gqueries/s=%.2f\n", kRec, kByte,
           (unsigned)(((uint64_t)rec_mask + 1) * 16 >> 20), (unsigned)(((uint64_t)byte_mask + 1) >> 20), best, q / (best * 1e-3) / 1e9);
}

int main(int argc, char **argv) {
    if (argc > 1) {  // optional: L2 fetch granularity in bytes (32 / 64 / 128)
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])));
    }
    size_t gran = 0;
    CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
    printf("cudaLimitMaxL2FetchGranularity=%zu\n", gran);
    const uint64_t n = 1ull << 29;  // 512 Mi keys = 2 GiB of keys per launch
    uint32_t *keys, *table, *out;
    CK(cudaMalloc(&keys, n * 4));
    CK(cudaMalloc(&table, 1ull << 30));
    CK(cudaMalloc(&out, n));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int log2_entries = 22; log2_entries <= 26; ++log2_entries) {
        uint32_t mask = (1u << log2_entries) - 1;
        gen_keys<<<(unsigned)(n / 256), 256>>>(keys, n, mask, 12345);
        CK(cudaMemset(table, 0, 1ull << 28));
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(a);
            red_kernel<<<(unsigned)(n / 4 / 256), 256>>>((const uint4 *)keys, n / 4, table);
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it && ms < best) best = ms;
        }
        printf("red_u32 table_MiB=%u ms=%.3f Gatomics/s=%.2f\n", (1u << log2_entries) * 4 >> 20, best, n / (best * 1e-3) / 1e9);
    }
    for (int log2_entries = 20; log2_entries <= 26; log2_entries += 2) {  // 16-byte records
        uint32_t mask = (1u << log2_entries) - 1;
        gen_keys<<<(unsigned)(n / 256), 256>>>(keys, n, mask, 777);
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(a);
            gather16_kernel<<<(unsigned)(n / 4 / 256), 256>>>((const uint4 *)keys, n / 4, (const uint4 *)table, out);
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it && ms < best) best = ms;
        }
        printf("gather16 table_MiB=%u ms=%.3f Ggathers/s=%.2f\n", (1u << log2_entries) * 16 >> 20, best, n / (best * 1e-3) / 1e9);
    }
    {   // the lookup's request shape: 100 Mi queries of 23 bytes, record table 16 / 32 / 64 MiB, byte table 32 / 64 MiB
        const uint64_t q = 100ull << 20;
        uint8_t *recs, *bytes;
        CK(cudaMalloc(&recs, q * 23 + 64));
        CK(cudaMalloc(&bytes, 64ull << 20));
        gen_keys<<<(unsigned)((q * 23 / 4 + 255) / 256), 256>>>((uint32_t *)recs, q * 23 / 4, 0xFFFFFFFFu, 99);
        CK(cudaMemset(bytes, 1, 64ull << 20));
        CK(cudaDeviceSynchronize());
        run_lookup_shape<3, 1>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, (32u << 20) - 1, out, a, b);  // 16 MiB + 32 MiB
        run_lookup_shape<3, 1>(recs, q, (const uint4 *)table, (1u << 21) - 1, bytes, (64u << 20) - 1, out, a, b);  // 32 MiB + 64 MiB (> L2 share)
        run_lookup_shape<3, 0>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, 0, out, a, b);                // 16 MiB
        run_lookup_shape<3, 0>(recs, q, (const uint4 *)table, (1u << 22) - 1, bytes, 0, out, a, b);                // 64 MiB (the fused C2 records: 61.5 MB)
        run_lookup_shape<4, 0>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, 0, out, a, b);
        run_lookup_shape<2, 0>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, 0, out, a, b);
        // the same with the TMA-ring input path of the product kernel: the request ceiling of tf23_stream_kernel
        run_lookup_shape_stream<3, 1>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, (32u << 20) - 1, out, a, b);
        run_lookup_shape_stream<3, 1>(recs, q, (const uint4 *)table, (1u << 21) - 1, bytes, (64u << 20) - 1, out, a, b);
        run_lookup_shape_stream<3, 0>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, 0, out, a, b);
        run_lookup_shape_stream<3, 0>(recs, q, (const uint4 *)table, (1u << 22) - 1, bytes, 0, out, a, b);
        run_lookup_shape_stream<2, 0>(recs, q, (const uint4 *)table, (1u << 20) - 1, bytes, 0, out, a, b);
    }
    return 0;
}
