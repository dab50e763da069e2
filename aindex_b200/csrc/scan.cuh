// scan.cuh -- device-wide exclusive prefix sum (three kernels: tile reduce, tile scan, down-sweep)
// over a functor, and the decoupled look-back primitives shared by the positions emit pass
// (positions.cu) and the radix sort (radix_sort.cu).
//
// Reference loop replaced by exclusive_scan: AIndexCompressed ctor src/hash.hpp:373-378
// (indices[i] = indices[i-1] + tf[i-1]) and AIndex13 ctor src/compute_aindex13.cpp:58-64.
#pragma once
#include "aix_internal.cuh"

namespace aix {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ unsigned long long block_scan_u64(unsigned long long v, unsigned long long *sm /*>=33*/,
                                                             unsigned long long &total) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) sm[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned long long t = lane < nw ? sm[lane] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o);
            if (lane >= (unsigned)o) t += y;
        }
        sm[lane] = t;
    }
    __syncthreads();
    total = sm[nw - 1];
    unsigned long long off = wid ? sm[wid - 1] : 0ull;
    __syncthreads();
    return off + x - v;
}

template <typename F>
__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(F f, uint64_t n, unsigned long long *__restrict__ tile_sum) {
    __shared__ unsigned long long sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
        if (base + j < n) s += f(base + j);
    unsigned long long total;
    block_scan_u64(s, sm, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// exclusive scan of the tile sums in place; tile_sum[n_tiles] receives the grand total
static __global__ void scan_tiles_kernel(unsigned long long *__restrict__ tile_sum, uint64_t n_tiles) {
    __shared__ unsigned long long sm[33];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t i0 = 0; i0 < n_tiles; i0 += blockDim.x) {
        uint64_t i = i0 + threadIdx.x;
        unsigned long long v = i < n_tiles ? tile_sum[i] : 0ull, total;
        unsigned long long ex = block_scan_u64(v, sm, total);
        if (i < n_tiles) tile_sum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sum[n_tiles] = carry;
}

template <typename F>
__global__ void __launch_bounds__(kScanBlock) scan_down_kernel(F f, uint64_t n, const unsigned long long *__restrict__ tile_off,
                                                             unsigned long long *__restrict__ out /* n+1 */) {
    __shared__ unsigned long long sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    unsigned long long v[kScanItems], s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        v[j] = base + j < n ? f(base + j) : 0ull;
        s += v[j];
    }
    unsigned long long total;
    unsigned long long run = tile_off[blockIdx.x] + block_scan_u64(s, sm, total);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
        if (base + j + 1 == n) out[n] = run;
    }
}

static inline uint64_t scan_tiles(uint64_t n) { return (n + kScanTile - 1) / kScanTile; }
// bytes of tile scratch exclusive_scan needs for n items
static inline size_t scan_scratch_bytes(uint64_t n) { return (scan_tiles(n) + 2) * 8; }

// out[i] = sum of f(j), j < i, for i in [0, n]  (out has n + 1 entries)
template <typename F>
static int exclusive_scan(aix_ctx *ctx, cudaStream_t st, F f, uint64_t n, unsigned long long *out, void *tile_scratch) {
    uint64_t tiles = scan_tiles(n);
    if (n == 0) {
        AIX_CUDA(ctx, cudaMemsetAsync(out, 0, 8, st));
        return AIX_OK;
    }
    scan_reduce_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(f, n, (unsigned long long *)tile_scratch);
    AIX_LAUNCH_CHECK(ctx);
    scan_tiles_kernel<<<1, 1024, 0, st>>>((unsigned long long *)tile_scratch, tiles);
    AIX_LAUNCH_CHECK(ctx);
    scan_down_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(f, n, (const unsigned long long *)tile_scratch, out);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

// ---- decoupled look-back -----------------------------------------------------------------
// One 64-bit status word per (tile, counter): bits 63:62 = state, bits 61:0 = value.  A word is
// written whole, so no fence is needed between value and state; readers use relaxed GPU-scope
// loads (never served from L1).  Tiles take their id from an atomic counter, so every
// predecessor of a running tile is running or finished: the spin always terminates.
constexpr unsigned long long kLbAggregate = 1ull << 62;  // value = this tile's own count
constexpr unsigned long long kLbInclusive = 2ull << 62;  // value = count of this tile and all before it
constexpr unsigned long long kLbValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive prefix of `mine` over all tiles before `tile`; status words of this counter are
// `stride` apart (status[t * stride] belongs to tile t).  Publishes this tile's words.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *status, uint64_t stride, uint64_t tile,
                                                                 unsigned long long mine) {
    if (tile == 0) {
        st_relaxed_u64(status, kLbInclusive | mine);
        return 0;
    }
    st_relaxed_u64(status + tile * stride, kLbAggregate | mine);
    unsigned long long excl = 0;
    for (uint64_t p = tile; p-- > 0;) {
        unsigned long long s;
        do {
            s = ld_relaxed_u64(status + p * stride);
        } while ((s >> 62) == 0);
        excl += s & kLbValueMask;
        if ((s >> 62) == 2) break;
    }
    st_relaxed_u64(status + tile * stride, kLbInclusive | (excl + mine));
    return excl;
}

}  // namespace aix
