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
    return 0;
}
