// radix_sort.cu -- hand-written LSD radix sort of 64-bit keys (one scatter kernel per digit with
// decoupled look-back, "onesweep" style) and run-length encoding of a sorted key array.
//
// Users: the positions-index build (positions.cu: packed (bucket << pos_bits | position) keys, sorted on
// the bucket bits only -- the sort is stable and the keys are emitted in position order, so every bucket
// comes out ascending = the order of the 1-thread reference worker, src/hash.cpp:1006-1051) and the
// canonical 23-mer table (mphf_build.cu: sort + run-length = the `sort | uniq -c` of the reference's
// jellyfish / kmer_counter stage, scripts/compute_aindex.py:140-182).
//
// Per pass every key is read once and written once (16 B); the digit histograms of all passes come from
// one extra read of the keys.  A tile is 8192 keys: ranked inside each warp with match_any (stable),
// staged in shared memory in digit order, then written out as runs of ~32 keys (256 B) per digit.
#include "radix_sort.cuh"
#include "scan.cuh"

namespace aix {

constexpr int kRsItems = 16;
constexpr int kRsMaxBits = 8;
constexpr int kRsMaxRadix = 1 << kRsMaxBits;
constexpr int kRsMaxPasses = 8;
constexpr int kRsMaxRanges = 16;
// stage[tile] (aliased by the per-warp histograms) + total[256] + dstart[256] + gbase[256]
constexpr size_t rs_smem(int threads) { return (size_t)threads * kRsItems * 8 + kRsMaxRadix * 4 * 2 + kRsMaxRadix * 8; }

struct RsPlan {
    int begin_bit, n_pass, bits;  // digit p covers [begin_bit + p*bits, min(end_bit, begin_bit + (p+1)*bits))
    int end_bit;
    __host__ __device__ int shift(int p) const { return begin_bit + p * bits; }
    __host__ __device__ int width(int p) const {
        int w = end_bit - shift(p);
        return w < bits ? w : bits;
    }
};

// the digit of a key: a bit field (the sort passes) or the index of the range that holds the key (stable partition
// of packed keys by owner, positions.cu: range r = [bound[r], bound[r + 1]))
struct BitsDigit {
    int shift;
    uint32_t mask;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const { return (uint32_t)(k >> shift) & mask; }
};
struct RangeDigit {
    unsigned long long bound[kRsMaxRanges];  // bound[0] = 0; entries from n_ranges on are unused
    int n_ranges;
    __device__ __forceinline__ uint32_t operator()(uint64_t k) const {
        uint32_t o = 0;
#pragma unroll
        for (int r = 1; r < kRsMaxRanges; ++r)
            if (r < n_ranges && k >= bound[r]) o = r;
        return o;
    }
};

// single-digit histogram for any digit functor (the range partition)
template <typename Digit>
__global__ void __launch_bounds__(512) rs_hist1_kernel(const uint64_t *__restrict__ keys, uint64_t n, Digit dg,
                                                     unsigned long long *__restrict__ hist) {
    __shared__ uint32_t sh[kRsMaxRadix];
    for (int i = threadIdx.x; i < kRsMaxRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + blockDim.x - 1) / blockDim.x * blockDim.x;
    const uint64_t lo = (uint64_t)blockIdx.x * per;
    const uint64_t hi = lo + per < n ? lo + per : n;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&sh[dg(keys[i])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kRsMaxRadix; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// hist[p][d] += number of keys whose digit p is d, for every pass in one read of the keys
__global__ void __launch_bounds__(512) rs_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, RsPlan plan,
                                                    unsigned long long *__restrict__ hist) {
    __shared__ uint32_t sh[kRsMaxPasses * kRsMaxRadix];
    for (int i = threadIdx.x; i < plan.n_pass * kRsMaxRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + blockDim.x - 1) / blockDim.x * blockDim.x;
    const uint64_t lo = (uint64_t)blockIdx.x * per;
    const uint64_t hi = lo + per < n ? lo + per : n;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint64_t k = keys[i];
        for (int p = 0; p < plan.n_pass; ++p) {
            const uint32_t d = (uint32_t)(k >> plan.shift(p)) & ((1u << plan.width(p)) - 1u);
            atomicAdd(&sh[p * kRsMaxRadix + d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.n_pass * kRsMaxRadix; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// base[p][d] = number of keys whose digit p is below d  (one CTA of 256 threads per pass)
__global__ void __launch_bounds__(kRsMaxRadix) rs_base_kernel(const unsigned long long *__restrict__ hist,
                                                            unsigned long long *__restrict__ base) {
    __shared__ unsigned long long sm[33];
    unsigned long long total;
    const unsigned long long v = hist[blockIdx.x * kRsMaxRadix + threadIdx.x];
    base[blockIdx.x * kRsMaxRadix + threadIdx.x] = block_scan_u64(v, sm, total);
}

// The lanes of the warp whose digit equals this lane's: kBits ballots.  MATCH.ANY gives the same mask in one instruction,
// but it runs on the SM's address-divergence unit at ~2 cycles per distinct value in the warp: with 128-256 digit values
// almost every lane is distinct, and ncu showed that pipe at 85 % with DRAM at 28 % (profiles/r02_c5sort_ncu.txt).
// Ballots run on the ALU / vote path at full rate.
template <int kBits>
__device__ __forceinline__ uint32_t warp_peers(uint32_t d, bool valid) {
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
    for (int b = 0; b < kBits; ++b) {
        const bool p = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, p);
        peers &= p ? m : ~m;
    }
    return peers;
}

// One digit pass.  Stable: equal digits keep their input order.  256 threads x 16 keys = 4096-key tiles, 4 CTAs / SM
// (512-thread tiles were slower: 55.5 vs 48.5 ms per 6.4 G-key pass).  kBits = width of the digit (the number of ballots
// per key); kBits = 0 selects the match_any ranking with a run-time width (kept for A/B runs, AIX_RS_RANK=match).
template <int kRsThreads, int kBits, int kMinBlocks, typename Digit>
__global__ void __launch_bounds__(kRsThreads, kMinBlocks)
rs_pass_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n, Digit dg, int bits,
               const unsigned long long *__restrict__ digit_base, unsigned long long *__restrict__ status,
               unsigned int *__restrict__ tile_counter) {
    constexpr int kRsTile = kRsThreads * kRsItems, kRsWarps = kRsThreads / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *stage = (uint64_t *)smem;                                        // [kRsTile]
    uint32_t *whist = (uint32_t *)smem;                                        // [kRsWarps][radix], dead before stage is written
    uint32_t *s_total = (uint32_t *)(smem + (size_t)kRsTile * 8);              // [256] keys of this tile per digit
    uint32_t *s_dstart = s_total + kRsMaxRadix;                                // [256] first stage slot of the digit
    unsigned long long *s_gbase = (unsigned long long *)(s_dstart + kRsMaxRadix);  // [256] out index of stage slot 0 of the digit
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_wsum[kRsMaxRadix / 32];

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t radix = 1u << bits;  // kBits >= bits (ballots above the width see zero bits)
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (uint32_t i = tid; i < kRsWarps * radix; i += kRsThreads) whist[i] = 0;
    __syncthreads();
    const uint64_t tile = s_tile;
    const uint64_t tile_base = tile * kRsTile;
    const uint32_t tile_n = n - tile_base < (uint64_t)kRsTile ? (uint32_t)(n - tile_base) : (uint32_t)kRsTile;

    // keys, warp-striped: item j of lane l of warp w = tile_base + w*32*ITEMS + j*32 + l
    uint64_t key[kRsItems];
    uint32_t pos2[kRsItems / 2];  // two 16-bit slots per word: rank in the warp's digit group, later the stage slot
    const uint32_t wbase = warp * (32 * kRsItems) + lane;
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t idx = wbase + j * 32;
        key[j] = idx < tile_n ? (uint64_t)__ldcs((const unsigned long long *)(in + tile_base + idx)) : ~0ull;
    }
    // rank inside the warp: lanes with the same digit form a group.  Every lane of a group reads the digit's running
    // count of this warp (one broadcast LDS), the group's lowest lane then adds the group size.
    uint32_t *my_hist = whist + warp * radix;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const bool valid = wbase + j * 32 < tile_n;
        const uint32_t d = dg(key[j]);
        uint32_t peers;
        if (kBits) peers = warp_peers<kBits>(d, valid);
        else peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 0xFFFFFFFFu);
        const uint32_t base = valid ? my_hist[d] : 0u;
        __syncwarp();
        if (valid && (peers & lt) == 0u) my_hist[d] = base + __popc(peers);
        __syncwarp();
        const uint32_t r = (base + __popc(peers & lt)) & 0xFFFFu;
        if (j & 1) pos2[j >> 1] |= r << 16;
        else pos2[j >> 1] = r;
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, tile total; publish the aggregate at once
    unsigned long long *my_status = status + tile * radix;
    uint32_t total = 0;
    if (tid < radix) {
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = whist[w * radix + tid];
            whist[w * radix + tid] = total;
            total += c;
        }
        s_total[tid] = total;
        st_relaxed_u64(my_status + tid, (tile == 0 ? kLbInclusive : kLbAggregate) | total);
    }
    // exclusive scan of the totals over the digits -> first stage slot of every digit
    if (tid < kRsMaxRadix) {
        uint32_t x = tid < radix ? total : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) s_wsum[warp] = x;
        s_dstart[tid] = x - (tid < radix ? total : 0u);  // exclusive inside the warp
    }
    __syncthreads();
    if (tid < kRsMaxRadix) {
        uint32_t add = 0;
        for (unsigned w = 0; w < warp; ++w) add += s_wsum[w];
        s_dstart[tid] += add;
    }
    __syncthreads();
    // stage slot of every key
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t d = dg(key[j]) & (radix - 1u);
        const uint32_t add = my_hist[d] + s_dstart[d];
        pos2[j >> 1] += (j & 1) ? add << 16 : add;  // slots stay below 4096: no carry out of either half
    }
    __syncthreads();  // whist is dead from here on: stage may overwrite it
#pragma unroll
    for (int j = 0; j < kRsItems; ++j)
        if (wbase + j * 32 < tile_n) stage[(j & 1) ? pos2[j >> 1] >> 16 : pos2[j >> 1] & 0xFFFFu] = key[j];
    // look back over the earlier tiles, one thread per digit
    if (tid < radix) {
        unsigned long long excl = 0;
        if (tile > 0) {
            // four predecessors are read at once (independent loads), then consumed in order
            uint64_t p = tile;
            bool done = false;
            while (!done) {
                unsigned long long s4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) s4[u] = p > (uint64_t)u ? ld_relaxed_u64(status + (p - 1 - u) * radix + tid) : 0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (done || p == 0) break;
                    unsigned long long s = s4[u];
                    while ((s >> 62) == 0) s = ld_relaxed_u64(status + (p - 1) * radix + tid);
                    excl += s & kLbValueMask;
                    --p;
                    if ((s >> 62) == 2 || p == 0) done = true;
                }
            }
            st_relaxed_u64(my_status + tid, kLbInclusive | (excl + total));
        }
        s_gbase[tid] = digit_base[tid] + excl - s_dstart[tid];
    }
    __syncthreads();
    for (uint32_t i = tid; i < tile_n; i += kRsThreads) {
        const uint64_t k = stage[i];
        const uint32_t d = dg(k);
        out[s_gbase[d] + i] = k;
    }
}

// ---- run-length encoding of a sorted array ---------------------------------------------------
struct HeadFlag {
    const uint64_t *k;
    __device__ unsigned long long operator()(uint64_t i) const { return (i == 0 || k[i] != k[i - 1]) ? 1ull : 0ull; }
};

// run r starts at the r-th head: uniq[r] = key, starts[r] = its index
__global__ void __launch_bounds__(kScanBlock) rle_heads_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                             const unsigned long long *__restrict__ tile_off,
                                                             uint64_t *__restrict__ uniq, unsigned long long *__restrict__ starts) {
    __shared__ unsigned long long sm[33];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t k[kScanItems + 1];
    k[0] = (base > 0 && base - 1 < n) ? keys[base - 1] : 0;
    unsigned long long s = 0;
    uint32_t heads = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        k[j + 1] = base + j < n ? keys[base + j] : 0;
        const bool h = base + j < n && (base + j == 0 || k[j + 1] != k[j]);
        heads |= (h ? 1u : 0u) << j;
        s += h ? 1 : 0;
    }
    unsigned long long total;
    unsigned long long r = tile_off[blockIdx.x] + block_scan_u64(s, sm, total);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if ((heads >> j) & 1u) {
            uniq[r] = k[j + 1];
            starts[r] = base + j;
            ++r;
        }
    }
}

__global__ void rle_counts_kernel(const unsigned long long *__restrict__ starts, uint64_t n_runs, uint64_t n,
                                  uint32_t *__restrict__ counts) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const unsigned long long e = r + 1 < n_runs ? starts[r + 1] : n;
    const unsigned long long c = e - starts[r];
    counts[r] = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c;  // the tf arrays are u32 (src/hash.hpp:97)
}

// ranking variant of the pass kernel: ballots (default) or match_any (AIX_RS_RANK=match, for A/B runs)
static bool rs_rank_by_match() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("AIX_RS_RANK");
        v = (e && !strcmp(e, "match")) ? 1 : 0;
    }
    return v == 1;
}

template <int kBits, int kMinBlocks, typename Digit>
static cudaError_t rs_launch_one(unsigned tiles, cudaStream_t st, const uint64_t *in, uint64_t *out, uint64_t n, const Digit &dg, int bits,
                                 const unsigned long long *base, unsigned long long *status, unsigned int *counter) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(rs_pass_kernel<256, kBits, kMinBlocks, Digit>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)rs_smem(256));
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    rs_pass_kernel<256, kBits, kMinBlocks, Digit><<<tiles, 256, rs_smem(256), st>>>(in, out, n, dg, bits, base, status, counter);
    return cudaGetLastError();
}

// the sort passes: digit width 1..8, ballot ranking; AIX_RS_RANK=match / AIX_RS_MINBLOCKS=3 select the A/B variants
static cudaError_t rs_launch_pass(unsigned tiles, cudaStream_t st, const uint64_t *in, uint64_t *out, uint64_t n, const BitsDigit &dg, int bits,
                                  const unsigned long long *base, unsigned long long *status, unsigned int *counter) {
    static int three = -1;
    if (three < 0) {
        const char *e = getenv("AIX_RS_MINBLOCKS");
        three = (e && atoi(e) == 3) ? 1 : 0;
    }
    if (rs_rank_by_match()) return rs_launch_one<0, 4>(tiles, st, in, out, n, dg, bits, base, status, counter);
#define AIX_RS_CASE(B)                                                                                          \
    case B:                                                                                                     \
        return three ? rs_launch_one<B, 3>(tiles, st, in, out, n, dg, bits, base, status, counter)              \
                     : rs_launch_one<B, 4>(tiles, st, in, out, n, dg, bits, base, status, counter);
    switch (bits) {
        AIX_RS_CASE(1) AIX_RS_CASE(2) AIX_RS_CASE(3) AIX_RS_CASE(4) AIX_RS_CASE(5) AIX_RS_CASE(6) AIX_RS_CASE(7)
        default: return three ? rs_launch_one<8, 3>(tiles, st, in, out, n, dg, bits, base, status, counter)
                              : rs_launch_one<8, 4>(tiles, st, in, out, n, dg, bits, base, status, counter);
    }
#undef AIX_RS_CASE
}

// the range partition: at most 16 ranges = 4 ballots (rounds above the digit's width see zero bits and change nothing)
static cudaError_t rs_launch_pass(unsigned tiles, cudaStream_t st, const uint64_t *in, uint64_t *out, uint64_t n, const RangeDigit &dg, int bits,
                                  const unsigned long long *base, unsigned long long *status, unsigned int *counter) {
    return rs_launch_one<4, 4>(tiles, st, in, out, n, dg, bits, base, status, counter);
}

int radix_sort_u64(aix_ctx *ctx, cudaStream_t st, uint64_t *keys, uint64_t *alt, uint64_t n, int begin_bit, int end_bit,
                   uint64_t **sorted) {
    *sorted = keys;
    if (begin_bit < 0 || end_bit > 64 || (n && (!keys || !alt))) return ctx->fail(AIX_ERR_ARG, "radix sort: bad arguments");
    if (n < 2 || end_bit <= begin_bit) return AIX_OK;
    RsPlan plan;
    plan.begin_bit = begin_bit;
    plan.end_bit = end_bit;
    const int total_bits = end_bit - begin_bit;
    plan.n_pass = (total_bits + kRsMaxBits - 1) / kRsMaxBits;
    plan.bits = (total_bits + plan.n_pass - 1) / plan.n_pass;
    const int tile_keys = 256 * kRsItems;
    const uint64_t tiles = (n + tile_keys - 1) / tile_keys;
    if (tiles >= (1ull << 31)) return ctx->fail(AIX_ERR_ARG, "radix sort: too many keys");
    AixTrace trace(st, "radix sort");
    // scratch: hist[8][256] | base[8][256] | counters[8] (+pad) | status[tiles][radix]
    const size_t front = (size_t)kRsMaxPasses * kRsMaxRadix * 8 * 2 + 64;
    const size_t status_bytes = (size_t)tiles * ((size_t)1 << plan.bits) * 8;
    unsigned char *scratch = nullptr;
    cudaError_t e = aix_pool_alloc(ctx, &scratch, front + status_bytes, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_NOMEM, "radix sort scratch (%zu bytes): %s", front + status_bytes, cudaGetErrorString(e));
    }
    unsigned long long *hist = (unsigned long long *)scratch;
    unsigned long long *base = hist + kRsMaxPasses * kRsMaxRadix;
    unsigned int *counters = (unsigned int *)(base + kRsMaxPasses * kRsMaxRadix);
    unsigned long long *status = (unsigned long long *)(scratch + front);
    auto fail = [&](cudaError_t err, const char *what) {
        cudaGetLastError();
        aix_pool_free(ctx, scratch, st);
        return ctx->fail(AIX_ERR_CUDA, "radix sort %s: %s", what, cudaGetErrorString(err));
    };
    if ((e = cudaMemsetAsync(scratch, 0, front, st)) != cudaSuccess) return fail(e, "memset");
    unsigned hgrid = (unsigned)((n + 512ull * 64 - 1) / (512ull * 64));
    const unsigned hmax = (unsigned)ctx->sm_count * 8u;
    if (hgrid > hmax) hgrid = hmax;
    if (hgrid < 1) hgrid = 1;
    trace.mark("scratch allocation");
    rs_hist_kernel<<<hgrid, 512, 0, st>>>(keys, n, plan, hist);
    rs_base_kernel<<<plan.n_pass, kRsMaxRadix, 0, st>>>(hist, base);
    ctx->launches += 2;
    trace.mark("digit histograms of all passes");
    uint64_t *src = keys, *dst = alt;
    for (int p = 0; p < plan.n_pass; ++p) {
        if ((e = cudaMemsetAsync(status, 0, (size_t)tiles * ((size_t)1 << plan.width(p)) * 8, st)) != cudaSuccess) return fail(e, "memset");
        if ((e = rs_launch_pass((unsigned)tiles, st, src, dst, n, BitsDigit{plan.shift(p), (1u << plan.width(p)) - 1u}, plan.width(p),
                                base + p * kRsMaxRadix, status, counters + p)) != cudaSuccess)
            return fail(e, "pass launch");
        ctx->launches++;
        trace.mark("digit pass");
        uint64_t *t = src; src = dst; dst = t;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "launch");
    aix_pool_free(ctx, scratch, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return ctx->fail(AIX_ERR_CUDA, "radix sort run: %s", cudaGetErrorString(e));
    trace.mark("free scratch");
    *sorted = src;
    return AIX_OK;
}

// Stable partition of n keys into n_ranges key ranges (bound[r] = first key of range r, bound[0] = 0): out = the keys
// grouped by range in range order, input order kept inside a range; counts[r] = keys of range r (host array).
int partition_by_range(aix_ctx *ctx, cudaStream_t st, const uint64_t *keys, uint64_t *out, uint64_t n, const uint64_t *bound,
                       int n_ranges, uint64_t *counts) {
    if (n_ranges < 1 || n_ranges > kRsMaxRanges) return ctx->fail(AIX_ERR_ARG, "partition: 1..%d ranges", kRsMaxRanges);
    for (int r = 0; r < n_ranges; ++r) counts[r] = 0;
    if (n == 0) return AIX_OK;
    RangeDigit dg;
    for (int r = 0; r < kRsMaxRanges; ++r) dg.bound[r] = r < n_ranges ? bound[r] : ~0ull;
    dg.bound[0] = 0;
    dg.n_ranges = n_ranges;
    int bits = 1;
    while ((1 << bits) < n_ranges) ++bits;
    const uint64_t tiles = (n + 256 * kRsItems - 1) / (256 * kRsItems);
    if (tiles >= (1ull << 31)) return ctx->fail(AIX_ERR_ARG, "partition: too many keys");
    const size_t front = (size_t)kRsMaxRadix * 8 * 2 + 64;
    const size_t status_bytes = (size_t)tiles * ((size_t)1 << bits) * 8;
    unsigned char *scratch = nullptr;
    cudaError_t e = aix_pool_alloc(ctx, &scratch, front + status_bytes, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_NOMEM, "partition scratch: %s", cudaGetErrorString(e));
    }
    unsigned long long *hist = (unsigned long long *)scratch, *base = hist + kRsMaxRadix;
    unsigned int *counter = (unsigned int *)(base + kRsMaxRadix);
    unsigned long long *status = (unsigned long long *)(scratch + front);
    e = cudaMemsetAsync(scratch, 0, front + status_bytes, st);
    unsigned hgrid = (unsigned)((n + 512ull * 64 - 1) / (512ull * 64));
    const unsigned hmax = (unsigned)ctx->sm_count * 8u;
    if (hgrid > hmax) hgrid = hmax;
    if (hgrid < 1) hgrid = 1;
    rs_hist1_kernel<RangeDigit><<<hgrid, 512, 0, st>>>(keys, n, dg, hist);
    rs_base_kernel<<<1, kRsMaxRadix, 0, st>>>(hist, base);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = rs_launch_pass((unsigned)tiles, st, keys, out, n, dg, bits, base, status, counter);
    ctx->launches += 3;
    unsigned long long h[kRsMaxRanges];
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, hist, sizeof h, cudaMemcpyDeviceToHost, st);
    aix_pool_free(ctx, scratch, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_CUDA, "partition: %s", cudaGetErrorString(e));
    }
    for (int r = 0; r < n_ranges; ++r) counts[r] = h[r];
    return AIX_OK;
}

int rle_u64(aix_ctx *ctx, cudaStream_t st, const uint64_t *sorted, uint64_t n, uint64_t *uniq, uint32_t *counts,
            uint64_t *n_runs_out) {
    *n_runs_out = 0;
    if (n == 0) return AIX_OK;
    const uint64_t tiles = scan_tiles(n);
    unsigned long long *tile_off = nullptr, *starts = nullptr;
    auto fail = [&](cudaError_t err, const char *what) {
        cudaGetLastError();
        aix_pool_free(ctx, tile_off, st);
        aix_pool_free(ctx, starts, st);
        return ctx->fail(err == cudaErrorMemoryAllocation ? AIX_ERR_NOMEM : AIX_ERR_CUDA, "run-length %s: %s", what, cudaGetErrorString(err));
    };
    cudaError_t e = aix_pool_alloc(ctx, &tile_off, scan_scratch_bytes(n), st);
    if (e != cudaSuccess) return fail(e, "scratch");
    scan_reduce_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(HeadFlag{sorted}, n, tile_off);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tile_off, tiles);
    ctx->launches += 2;
    unsigned long long n_runs = 0;
    if ((e = cudaMemcpyAsync(&n_runs, tile_off + tiles, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e, "copy");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e, "count");
    if ((e = aix_pool_alloc(ctx, &starts, (n_runs + 1) * 8, st)) != cudaSuccess) return fail(e, "starts");
    rle_heads_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(sorted, n, tile_off, uniq, starts);
    rle_counts_kernel<<<aix_grid(n_runs, 256), 256, 0, st>>>(starts, n_runs, n, counts);
    ctx->launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "launch");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e, "run");
    aix_pool_free(ctx, tile_off, st);
    aix_pool_free(ctx, starts, st);
    *n_runs_out = n_runs;
    return AIX_OK;
}

}  // namespace aix

using namespace aix;

extern "C" {

int aix_sort_u64_dev(aix_ctx *ctx, uint64_t *keys_dev, uint64_t *alt_dev, uint64_t n, int begin_bit, int end_bit,
                     int *result_in_alt) {
    if (!ctx || !result_in_alt) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *sorted = keys_dev;
    AIX_TRY(radix_sort_u64(ctx, ctx->stream, keys_dev, alt_dev, n, begin_bit, end_bit, &sorted));
    *result_in_alt = sorted == alt_dev && sorted != keys_dev ? 1 : 0;
    return AIX_OK;
}

int aix_partition_u64_dev(aix_ctx *ctx, const uint64_t *keys_dev, uint64_t *out_dev, uint64_t n, const uint64_t *bounds,
                          int n_ranges, uint64_t *counts_out) {
    if (!ctx || !bounds || !counts_out || (n && (!keys_dev || !out_dev))) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return partition_by_range(ctx, ctx->stream, keys_dev, out_dev, n, bounds, n_ranges, counts_out);
}

int aix_rle_u64_dev(aix_ctx *ctx, const uint64_t *sorted_dev, uint64_t n, uint64_t *uniq_dev, uint32_t *counts_dev,
                    uint64_t *n_runs) {
    if (!ctx || !n_runs || (n && (!sorted_dev || !uniq_dev || !counts_dev))) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return rle_u64(ctx, ctx->stream, sorted_dev, n, uniq_dev, counts_dev, n_runs);
}

}  // extern "C"
