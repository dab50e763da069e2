// positions.cu -- K6/K7: positions index (CSR of 1-based read-file offsets per k-mer) build
// and batched query.
//
// Reference build: AIndexCompressed ctor src/hash.hpp:365-399 (exclusive prefix sum of tf),
// fill_index_from_reads :407-444, lu_compressed_worker src/hash.cpp:960-1060; 13-mer form
// src/compute_aindex13.cpp:36-86 (prefix sum), :125-239 (worker).  Parity target = the
// 1-thread reference: every bucket holds its occurrences in ascending order, only the first
// tf[h] are kept, the tail stays 0 when there are fewer.
// Reference query: get_positions_23mer python_wrapper.cpp:800-822, get_positions_13mer
// :1070-1101.
//
// GPU build = device prefix sum -> count pass (occurrences per bucket) -> scatter with one
// atomic cursor per bucket -> per-bucket ascending sort (atomic order is not deterministic)
// -> clip to tf.  When no bucket has more occurrences than tf (the normal case: tf was
// counted on the same reads) the scatter goes straight into the final layout.
#include <cub/device/device_radix_sort.cuh>

#include "aix_internal.cuh"
#include "query23.cuh"

namespace aix {

constexpr uint64_t kNoBucket = ~0ULL;
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

// ---- exclusive prefix sum of per-bucket sizes (u64 out, n+1 entries) -------------------
struct TfFromRecs {
    const uint4 *recs;
    __device__ uint64_t operator()(uint64_t i) const { return recs[i].z; }
};
struct TfFromU64 {
    const uint64_t *v;
    __device__ uint64_t operator()(uint64_t i) const { return v[i]; }
};
struct TfFromU32 {
    const uint32_t *v;
    __device__ uint64_t operator()(uint64_t i) const { return v[i]; }
};

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

__global__ void scan_tiles_kernel(unsigned long long *__restrict__ tile_sum, uint64_t n_tiles) {
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

template <typename F>
static int exclusive_scan(aix_ctx *ctx, cudaStream_t st, F f, uint64_t n, unsigned long long *out, void *tile_scratch) {
    uint64_t tiles = (n + kScanTile - 1) / kScanTile;
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

// ---- bucket of the window starting at byte i (or kNoBucket) ---------------------------
// k = 23: hash.cpp:1006-1051.  Skip windows containing '\n', '~', 'N'; canonical form by
// numeric compare of the strict 2-bit values; the forward form is hashed from the RAW bytes
// (so a lower-case / IUPAC byte makes the verify fail), the reverse form from the decoded
// string; checker verify.
__device__ __forceinline__ uint64_t bucket23(const Index23Dev &ix, const MphfDev &m, const uint8_t *p) {
    uint64_t r0, r1, r2;
    load_window23(p, r0, r1, r2);
    bool all_acgt;
    uint64_t u = encode_validate23(r0, r1, r2, all_acgt), r = revcomp23(u);
    if (all_acgt) {
        Hit h = lookup_packed23<true>(ix, m, u, r, true, r0, r1, r2);  // one probe of min(u, r)
        return h.strand ? h.h : kNoBucket;
    }
#pragma unroll
    for (int j = 0; j < 23; ++j) {
        uint64_t w = j < 8 ? r0 : (j < 16 ? r1 : r2);
        uint32_t ch = (uint32_t)(w >> (8 * (j & 7))) & 0xFFu;
        if (ch == '\n' || ch == '~' || ch == 'N') return kNoBucket;
    }
    uint64_t us = encode23_strict(r0, r1, r2), rs = revcomp23(us), a, b, c, h;
    uint32_t tf;
    if (us <= rs) {
        jenkins_short(m.seed, r0, r1, r2, 23u, a, b, c);
        h = mphf_eval(m, a, b, c);
        return probe23(ix, h, us, tf) ? h : kNoBucket;
    }
    h = mphf_lookup23(m, us);
    return probe23(ix, h, rs, tf) ? h : kNoBucket;
}

// k = 13: compute_aindex13.cpp:186-216.  Upper-case ACGT windows only, forward strand.
__device__ __forceinline__ uint64_t bucket13(const MphfDev &m, const uint8_t *p) {
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 13; ++j) {
        uint32_t ch = __ldg(p + j);
        if (!is_acgt_upper(ch)) return kNoBucket;
        v = (v << 2) | base_code_strict(ch);
    }
    uint64_t h = mphf_lookup13(m, revcomp13(v));
    return h < AIX_TOTAL_13MERS ? h : kNoBucket;
}

// phase 1: tmp[off[h] + cursor[h]++] = i + 1 (every occurrence, general path)
// phase 2: the optimistic single pass -- positions[indices[h] + cursor[h]++] = i + 1 while the slot is
//          below tf[h]; cursor[h] ends as the true occurrence count, so a bucket with more occurrences
//          than tf is detected afterwards (classify) and only then the general path runs
template <int K, int kPhase, typename F>
__global__ void __launch_bounds__(256) positions_scan_kernel(Index23Dev ix, MphfDev m, F tf, const uint8_t *__restrict__ reads,
                                                           uint64_t start, uint64_t n_win_end /* len-k+1 */,
                                                           uint32_t *__restrict__ cursor,
                                                           const unsigned long long *__restrict__ off,
                                                           unsigned long long *__restrict__ dst) {
    uint64_t i = start + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_win_end) return;
    uint64_t h = K == 23 ? bucket23(ix, m, reads + i) : bucket13(m, reads + i);
    if (h == kNoBucket) return;
    uint32_t slot = atomicAdd(cursor + h, 1u);
    if (kPhase == 1 || (uint64_t)slot < tf(h)) dst[off[h] + slot] = i + 1;
}

// any bucket with more occurrences than tf?  also classifies buckets by size for the sort
template <typename F>
__global__ void classify_kernel(F tf, const uint32_t *__restrict__ occ, uint64_t n, int *__restrict__ over,
                                uint32_t *__restrict__ medium_list, uint32_t *__restrict__ large_list,
                                unsigned int *__restrict__ counts /* [0] medium, [1] large */, uint32_t small_max,
                                uint32_t medium_max) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    uint32_t c = occ[h];
    if ((uint64_t)c > tf(h)) *over = 1;
    if (c > medium_max) large_list[atomicAdd(counts + 1, 1u)] = (uint32_t)h;
    else if (c > small_max) medium_list[atomicAdd(counts + 0, 1u)] = (uint32_t)h;
}

// small buckets: one thread each, insertion sort (atomic arrival order is nearly sorted)
__global__ void sort_small_kernel(unsigned long long *__restrict__ data, const unsigned long long *__restrict__ off,
                                  const uint32_t *__restrict__ occ, uint64_t n, uint32_t small_max) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    uint32_t c = occ[h];
    if (c < 2 || c > small_max) return;
    unsigned long long *a = data + off[h];
    for (uint32_t i = 1; i < c; ++i) {
        unsigned long long x = a[i];
        uint32_t j = i;
        while (j > 0 && a[j - 1] > x) {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = x;
    }
}

// medium buckets: one CTA each, bitonic sort in shared memory
constexpr uint32_t kSmallMax = 48;
constexpr uint32_t kMediumMax = 4096;
__global__ void __launch_bounds__(512) sort_medium_kernel(unsigned long long *__restrict__ data,
                                                        const unsigned long long *__restrict__ off,
                                                        const uint32_t *__restrict__ occ,
                                                        const uint32_t *__restrict__ medium_list) {
    __shared__ unsigned long long s[kMediumMax];
    const uint32_t h = medium_list[blockIdx.x];
    const uint32_t c = occ[h];
    unsigned long long *a = data + off[h];
    uint32_t p2 = 1;
    while (p2 < c) p2 <<= 1;
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) s[i] = i < c ? a[i] : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= p2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
                uint32_t l = i ^ j;
                if (l > i) {
                    unsigned long long x = s[i], y = s[l];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { s[i] = y; s[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) a[i] = s[i];
}

// general path: copy the first min(occ, tf) sorted occurrences into the final layout
template <typename F>
__global__ void clip_copy_kernel(F tf, const unsigned long long *__restrict__ tmp, const unsigned long long *__restrict__ tmp_off,
                                 const uint32_t *__restrict__ occ, const unsigned long long *__restrict__ indices, uint64_t n,
                                 unsigned long long *__restrict__ positions) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    uint64_t keep = occ[h];
    uint64_t cap = tf(h);
    if (keep > cap) keep = cap;
    const unsigned long long *src = tmp + tmp_off[h];
    unsigned long long *dst = positions + indices[h];
    for (uint64_t i = 0; i < keep; ++i) dst[i] = src[i];
}

// ---- K7 query ----------------------------------------------------------------------------
// counts pass / fill pass of get_positions: bucket slice, zeros skipped, values - 1
template <int K>
__global__ void positions_query_kernel(Index23Dev ix, MphfDev m, const unsigned long long *__restrict__ indices,
                                       uint64_t n_indices, const unsigned long long *__restrict__ positions,
                                       uint64_t n_positions, const uint8_t *__restrict__ recs, uint32_t stride,
                                       const uint8_t *__restrict__ lens, uint64_t q, unsigned long long *__restrict__ counts,
                                       const unsigned long long *__restrict__ out_off, unsigned long long *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < q;  // inactive lanes of the last warp still take part in the warp-wide slice scans
    uint32_t len = active ? (lens ? lens[i] : stride) : 0u;
    if (len > stride) len = stride;
    const uint8_t *p = recs + (active ? i : 0) * stride;
    uint64_t h = kNoBucket;
    if (active && len == (uint32_t)K) {
        if (K == 23) {
            // get_pfid (hash.hpp:150-170): single probe of the lexicographically smaller string
            uint64_t w[3] = {0, 0, 0};
            for (int j = 0; j < 23; ++j) w[j >> 3] |= (uint64_t)__ldg(p + j) << (8 * (j & 7));
            uint64_t us = encode23_strict(w[0], w[1], w[2]), rs = revcomp23(us), v0, v1, v2, a, b, c;
            ascii_words23_from_rc(us, v0, v1, v2);  // ASCII of rs
            uint32_t tf;
            if (cmp_words23(w[0], w[1], w[2], v0, v1, v2) <= 0) {
                jenkins_short(m.seed, w[0], w[1], w[2], 23u, a, b, c);
                uint64_t hh = mphf_eval(m, a, b, c);
                if (probe23(ix, hh, us, tf)) h = hh;
            } else {
                jenkins_short(m.seed, v0, v1, v2, 23u, a, b, c);
                uint64_t hh = mphf_eval(m, a, b, c);
                if (probe23(ix, hh, rs, tf)) h = hh;
            }
        } else {
            h = bucket13(m, p);  // python_wrapper.cpp:1076-1087: upper-case ACGT only
        }
    }
    // bucket slices are scanned by the whole warp, one query (lane) after the other: coalesced 256-byte reads and
    // writes, zeros squeezed out in order with a ballot (python_wrapper.cpp:816-820: skip 0, store pos - 1)
    unsigned long long b = 0, e = 0;
    if (h != kNoBucket && h + 1 < n_indices) {
        b = indices[h];
        e = indices[h + 1];
        if (e > n_positions) e = n_positions;  // python_wrapper.cpp:1092
        if (e < b) e = b;
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long w = (out && e > b) ? out_off[i] : 0;
    unsigned long long cnt = 0;
    unsigned todo = __ballot_sync(0xFFFFFFFFu, e > b);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned long long sb = __shfl_sync(0xFFFFFFFFu, b, src), se = __shfl_sync(0xFFFFFFFFu, e, src);
        const unsigned long long sw = __shfl_sync(0xFFFFFFFFu, w, src);
        unsigned long long done = 0;
        for (unsigned long long j0 = sb; j0 < se; j0 += 32) {
            const unsigned long long j = j0 + lane;
            const unsigned long long v = j < se ? __ldg(positions + j) : 0ull;
            const unsigned nz = __ballot_sync(0xFFFFFFFFu, v != 0);
            if (out && v) out[sw + done + __popc(nz & ((1u << lane) - 1u))] = v - 1;
            done += __popc(nz);
        }
        if ((int)lane == src) cnt = done;
    }
    if (counts && active) counts[i] = cnt;
}

}  // namespace aix

using namespace aix;

// worker prologue, hash.cpp:973-988 / compute_aindex13.cpp:133-147: advance the start past
// leading windows that contain '\n', '~' or '?'
static uint64_t first_start(const uint8_t *c, uint64_t len, uint64_t k) {
    uint64_t start = 0;
    while (start + k <= len) {
        bool found = false;
        for (uint64_t i = start; i < start + k; ++i) {
            if (c[i] == '\n' || c[i] == '~' || c[i] == '?') {
                start = i + 1;
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    return start;
}

// Device part of the build.  reads_dev must be readable for 8 bytes past len (aligned word
// loads of the last windows); `start` = first_start of the image.  On success *indices_dev
// (u64[n+1]) and *positions_dev (u64[total], at least one element) are owned by the caller.
template <int K, typename F>
static int build_core(aix_ctx *ctx, Index23Dev id, MphfDev md, F tf, uint64_t n, const uint8_t *reads_dev, uint64_t len,
                      uint64_t start, unsigned long long **indices_dev, unsigned long long **positions_dev,
                      uint64_t *total_out) {
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    unsigned long long *indices = nullptr, *positions = nullptr, *tmp = nullptr, *tmp_off = nullptr, *tiles = nullptr;
    uint32_t *occ = nullptr, *cursor = nullptr, *medium = nullptr, *large = nullptr;
    int *over = nullptr;
    unsigned int *cls = nullptr;
    auto cleanup_tmp = [&]() {
        cudaFree(tmp); cudaFree(tmp_off); cudaFree(tiles);
        cudaFree(occ); cudaFree(cursor); cudaFree(medium); cudaFree(large); cudaFree(over); cudaFree(cls);
    };
    auto cleanup = [&]() {
        cleanup_tmp();
        cudaFree(indices); cudaFree(positions);
    };
#define PB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                         \
            cleanup();                                                                                  \
            return ctx->fail(AIX_ERR_CUDA, "positions build %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
        }                                                                                               \
    } while (0)
    AixTrace trace(st, "positions build");
    const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile + 1;
    PB_CUDA(cudaMalloc(&indices, (n + 1) * 8));
    PB_CUDA(cudaMalloc(&tiles, n_tiles * 8));
    int rc = exclusive_scan(ctx, st, tf, n, indices, tiles);
    if (rc != AIX_OK) { cleanup(); return rc; }
    unsigned long long total = 0;
    PB_CUDA(cudaMemcpyAsync(&total, indices + n, 8, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    PB_CUDA(cudaMalloc(&positions, (total ? total : 1) * 8));
    PB_CUDA(cudaMemsetAsync(positions, 0, total * 8, st));
    if (len >= (uint64_t)K && n && total) {
        const uint64_t n_win_end = len - K + 1;
        if (start < n_win_end) {
            PB_CUDA(cudaMalloc(&occ, n * 4));
            PB_CUDA(cudaMalloc(&medium, n * 4));
            PB_CUDA(cudaMalloc(&large, n * 4));
            PB_CUDA(cudaMalloc(&over, sizeof(int)));
            PB_CUDA(cudaMalloc(&cls, 2 * sizeof(unsigned int)));
            PB_CUDA(cudaMemsetAsync(occ, 0, n * 4, st));
            PB_CUDA(cudaMemsetAsync(over, 0, sizeof(int), st));
            PB_CUDA(cudaMemsetAsync(cls, 0, 2 * sizeof(unsigned int), st));
            trace.mark("prefix sum + allocations");
            // grids are limited to 2^31-1 CTAs: scan the image in launches of <= 2^38 windows
            const uint64_t kLaunchWin = 1ull << 38;
            // optimistic single pass straight into the final layout (occ doubles as the cursor)
            for (uint64_t w0 = start; w0 < n_win_end; w0 += kLaunchWin) {
                const uint64_t w1 = n_win_end - w0 < kLaunchWin ? n_win_end : w0 + kLaunchWin;
                positions_scan_kernel<K, 2><<<aix_grid(w1 - w0, 256), 256, 0, st>>>(id, md, tf, reads_dev, w0, w1, occ, indices, positions);
                ctx->launches++;
            }
            trace.mark("scatter pass (lookup + cursor atomic + 8-byte store per window)");
            classify_kernel<<<aix_grid(n, 256), 256, 0, st>>>(tf, occ, n, over, medium, large, cls, kSmallMax, kMediumMax);
            ctx->launches++;
            int h_over = 0;
            unsigned int h_cls[2] = {0, 0};
            PB_CUDA(cudaMemcpyAsync(&h_over, over, sizeof(int), cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaMemcpyAsync(h_cls, cls, sizeof h_cls, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
            trace.mark("classify");
            unsigned long long *data = positions;
            const unsigned long long *off = indices;
            if (h_over) {  // some bucket overflows its tf: scatter everything aside, sort, then clip
                PB_CUDA(cudaMemsetAsync(positions, 0, total * 8, st));
                PB_CUDA(cudaMalloc(&tmp_off, (n + 1) * 8));
                rc = exclusive_scan(ctx, st, TfFromU32{occ}, n, tmp_off, tiles);
                if (rc != AIX_OK) { cleanup(); return rc; }
                unsigned long long tmp_total = 0;
                PB_CUDA(cudaMemcpyAsync(&tmp_total, tmp_off + n, 8, cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaStreamSynchronize(st));
                PB_CUDA(cudaMalloc(&tmp, (tmp_total ? tmp_total : 1) * 8));
                PB_CUDA(cudaMalloc(&cursor, n * 4));
                PB_CUDA(cudaMemsetAsync(cursor, 0, n * 4, st));
                data = tmp;
                off = tmp_off;
                for (uint64_t w0 = start; w0 < n_win_end; w0 += kLaunchWin) {
                    const uint64_t w1 = n_win_end - w0 < kLaunchWin ? n_win_end : w0 + kLaunchWin;
                    positions_scan_kernel<K, 1><<<aix_grid(w1 - w0, 256), 256, 0, st>>>(id, md, tf, reads_dev, w0, w1, cursor, off, data);
                    ctx->launches++;
                }
                trace.mark("general path: second scatter");
            }
            sort_small_kernel<<<aix_grid(n, 256), 256, 0, st>>>(data, off, occ, n, kSmallMax);
            ctx->launches++;
            if (h_cls[0]) {
                sort_medium_kernel<<<h_cls[0], 512, 0, st>>>(data, off, occ, medium);
                ctx->launches++;
            }
            if (h_cls[1]) {  // very large buckets (repeats): one device radix sort each
                std::vector<uint32_t> lg(h_cls[1]);
                PB_CUDA(cudaMemcpyAsync(lg.data(), large, (size_t)h_cls[1] * 4, cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaStreamSynchronize(st));
                for (uint32_t h : lg) {
                    unsigned long long o2[2];
                    uint32_t c = 0;
                    PB_CUDA(cudaMemcpy(o2, off + h, 8, cudaMemcpyDeviceToHost));
                    PB_CUDA(cudaMemcpy(&c, occ + h, 4, cudaMemcpyDeviceToHost));
                    unsigned long long *alt = nullptr;
                    void *ws = nullptr;
                    size_t ws_bytes = 0;
                    PB_CUDA(cudaMalloc(&alt, (size_t)c * 8));
                    cub::DeviceRadixSort::SortKeys(nullptr, ws_bytes, data + o2[0], alt, (int)c, 0, 64, st);
                    cudaError_t e2 = cudaMalloc(&ws, ws_bytes ? ws_bytes : 1);
                    if (e2 == cudaSuccess) e2 = cub::DeviceRadixSort::SortKeys(ws, ws_bytes, data + o2[0], alt, (int)c, 0, 64, st);
                    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(data + o2[0], alt, (size_t)c * 8, cudaMemcpyDeviceToDevice, st);
                    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
                    cudaFree(alt);
                    cudaFree(ws);
                    ctx->launches += 6;
                    PB_CUDA(e2);
                }
            }
            if (h_over) {
                clip_copy_kernel<<<aix_grid(n, 256), 256, 0, st>>>(tf, tmp, tmp_off, occ, indices, n, positions);
                ctx->launches++;
            }
            trace.mark("per-bucket sort (+ clip)");
        }
    }
    PB_CUDA(cudaStreamSynchronize(st));
#undef PB_CUDA
    cleanup_tmp();
    trace.mark("free scratch");
    *indices_dev = indices;
    *positions_dev = positions;
    *total_out = total;
    return AIX_OK;
}

// host image -> device build -> host arrays
template <int K, typename F>
static int build_impl(aix_ctx *ctx, Index23Dev id, MphfDev md, F tf, uint64_t n, const uint8_t *reads, uint64_t len,
                      uint64_t *indices_out, uint64_t *positions_out) {
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t *reads_dev = nullptr;
    cudaError_t e = cudaMalloc(&reads_dev, len + 64);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_NOMEM, "positions build: reads image: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(reads_dev, reads, len, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(reads_dev + len, '\n', 64, ctx->stream);
    unsigned long long *indices = nullptr, *positions = nullptr;
    uint64_t total = 0;
    int rc = build_core<K>(ctx, id, md, tf, n, reads_dev, len, first_start(reads, len, K), &indices, &positions, &total);
    cudaFree(reads_dev);
    if (rc != AIX_OK) return rc;
    e = cudaMemcpyAsync(indices_out, indices, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && total) e = cudaMemcpyAsync(positions_out, positions, total * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(indices);
    cudaFree(positions);
    if (e != cudaSuccess) return ctx->fail(AIX_ERR_CUDA, "positions build: download: %s", cudaGetErrorString(e));
    return AIX_OK;
}

// first_start of an image that lives in HBM: the prologue only looks at a prefix, fetched in pieces
static int first_start_dev(aix_ctx *ctx, const uint8_t *reads_dev, uint64_t len, uint64_t k, uint64_t *start_out) {
    const uint64_t kPiece = 1ull << 20;
    std::vector<uint8_t> buf;
    uint64_t base = 0;  // windows before `base` are known to be skipped
    while (true) {
        const uint64_t n = len - base < kPiece ? len - base : kPiece;
        buf.resize(n);
        AIX_CUDA(ctx, cudaMemcpy(buf.data(), reads_dev + base, n, cudaMemcpyDeviceToHost));
        const uint64_t s = first_start(buf.data(), n, k);
        if (s + k <= n || base + n == len) {  // settled inside the piece, or the image ended
            *start_out = base + s;
            return AIX_OK;
        }
        base += s > 0 ? s : 1;  // unreachable s == 0 (then s + k <= n unless n < k = image end)
    }
}

template <int K, typename F>
static int build_dev_impl(aix_ctx *ctx, Index23Dev id, MphfDev md, F tf, uint64_t n, const uint8_t *reads_dev, uint64_t len,
                          aix_positions **out) {
    *out = nullptr;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t start = 0;
    if (len >= (uint64_t)K) AIX_TRY(first_start_dev(ctx, reads_dev, len, K, &start));
    unsigned long long *indices = nullptr, *positions = nullptr;
    uint64_t total = 0;
    AIX_TRY((build_core<K>(ctx, id, md, tf, n, reads_dev, len, start, &indices, &positions, &total)));
    aix_positions *p = new aix_positions();
    p->n_indices = n + 1;
    p->n_positions = total;
    p->indices_dev = (uint64_t *)indices;
    p->positions_dev = (uint64_t *)positions;
    *out = p;
    return AIX_OK;
}

__global__ void sum_tf_kernel(const uint4 *__restrict__ recs, uint64_t n, unsigned long long *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = i < n ? recs[i].z : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
__global__ void sum_u64_kernel(const uint64_t *__restrict__ a, uint64_t n, unsigned long long *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = i < n ? a[i] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

extern "C" {

int aix_positions_total23(aix_ctx *ctx, const aix_index23 *ix, uint64_t *total) {
    if (!ctx || !ix || !total) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *acc;
    AIX_TRY(ctx->reserve(SCR_TMP1, 8, &acc));
    AIX_CUDA(ctx, cudaMemsetAsync(acc, 0, 8, ctx->stream));
    if (ix->n) {
        sum_tf_kernel<<<aix_grid(ix->n, 256), 256, 0, ctx->stream>>>(ix->recs_dev, ix->n, (unsigned long long *)acc);
        AIX_LAUNCH_CHECK(ctx);
    }
    AIX_CUDA(ctx, cudaMemcpyAsync(total, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

int aix_positions_total13(aix_ctx *ctx, const aix_index13 *ix, uint64_t *total) {
    if (!ctx || !ix || !total) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *acc;
    AIX_TRY(ctx->reserve(SCR_TMP1, 8, &acc));
    AIX_CUDA(ctx, cudaMemsetAsync(acc, 0, 8, ctx->stream));
    sum_u64_kernel<<<aix_grid(AIX_TOTAL_13MERS, 256), 256, 0, ctx->stream>>>(ix->tf_mphf_dev, AIX_TOTAL_13MERS, (unsigned long long *)acc);
    AIX_LAUNCH_CHECK(ctx);
    AIX_CUDA(ctx, cudaMemcpyAsync(total, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

int aix_positions_build23(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *reads, uint64_t len, uint64_t *indices_out,
                          uint64_t *positions_out) {
    if (!ctx || !ix || !indices_out || (len && !reads)) return AIX_ERR_ARG;
    return build_impl<23>(ctx, ix->dev(), ix->mphf->dev(), TfFromRecs{ix->recs_dev}, ix->n, reads, len, indices_out, positions_out);
}

int aix_positions_build13(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *reads, uint64_t len, uint64_t *indices_out,
                          uint64_t *positions_out) {
    if (!ctx || !ix || !indices_out || (len && !reads)) return AIX_ERR_ARG;
    Index23Dev none = {};
    return build_impl<13>(ctx, none, ix->mphf->dev(), TfFromU64{ix->tf_mphf_dev}, AIX_TOTAL_13MERS, reads, len, indices_out, positions_out);
}

int aix_positions_build23_dev(aix_ctx *ctx, const aix_index23 *ix, const uint8_t *reads_dev, uint64_t len,
                              aix_positions **out) {
    if (!ctx || !ix || !out || (len && !reads_dev)) return AIX_ERR_ARG;
    return build_dev_impl<23>(ctx, ix->dev(), ix->mphf->dev(), TfFromRecs{ix->recs_dev}, ix->n, reads_dev, len, out);
}

int aix_positions_build13_dev(aix_ctx *ctx, const aix_index13 *ix, const uint8_t *reads_dev, uint64_t len,
                              aix_positions **out) {
    if (!ctx || !ix || !out || (len && !reads_dev)) return AIX_ERR_ARG;
    Index23Dev none = {};
    return build_dev_impl<13>(ctx, none, ix->mphf->dev(), TfFromU64{ix->tf_mphf_dev}, AIX_TOTAL_13MERS, reads_dev, len, out);
}

int aix_positions_info(const aix_positions *p, uint64_t info[2]) {
    if (!p || !info) return AIX_ERR_ARG;
    info[0] = p->n_indices;
    info[1] = p->n_positions;
    return AIX_OK;
}

int aix_positions_arrays_dev(const aix_positions *p, const uint64_t **indices_dev, const uint64_t **positions_dev) {
    if (!p) return AIX_ERR_ARG;
    if (indices_dev) *indices_dev = p->indices_dev;
    if (positions_dev) *positions_dev = p->positions_dev;
    return AIX_OK;
}

int aix_positions_download(aix_ctx *ctx, const aix_positions *p, uint64_t *indices_out, uint64_t *positions_out) {
    if (!ctx || !p) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (indices_out && p->n_indices)
        AIX_CUDA(ctx, cudaMemcpyAsync(indices_out, p->indices_dev, p->n_indices * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (positions_out && p->n_positions)
        AIX_CUDA(ctx, cudaMemcpyAsync(positions_out, p->positions_dev, p->n_positions * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

int aix_positions_upload(aix_ctx *ctx, const uint64_t *indices, uint64_t n_indices, const uint64_t *positions,
                         uint64_t n_positions, aix_positions **out) {
    if (!ctx || !out || (n_indices && !indices) || (n_positions && !positions)) return AIX_ERR_ARG;
    *out = nullptr;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    aix_positions *p = new aix_positions();
    p->n_indices = n_indices; p->n_positions = n_positions;
    cudaError_t e = cudaMalloc(&p->indices_dev, (n_indices ? n_indices : 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&p->positions_dev, (n_positions ? n_positions : 1) * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->indices_dev, indices, n_indices * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->positions_dev, positions, n_positions * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        aix_positions_destroy(ctx, p);
        return ctx->fail(AIX_ERR_NOMEM, "positions upload: %s", cudaGetErrorString(e));
    }
    *out = p;
    return AIX_OK;
}

void aix_positions_destroy(aix_ctx *ctx, aix_positions *p) {
    if (!p) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (p->indices_dev) cudaFree(p->indices_dev);
    if (p->positions_dev) cudaFree(p->positions_dev);
    delete p;
}

int aix_positions_query_dev(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13, const aix_positions *p,
                            const uint8_t *recs_dev, uint32_t stride, const uint8_t *lens_dev, uint64_t q, int k,
                            uint64_t *counts_dev, const uint64_t *offs_dev, uint64_t *pos_out_dev) {
    if (!ctx || !p) return AIX_ERR_ARG;
    if (k == 23 && !ix23) return ctx->fail(AIX_ERR_STATE, "23-mer index not loaded");
    if (k == 13 && !ix13) return ctx->fail(AIX_ERR_STATE, "13-mer index not loaded");
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (q == 0) return AIX_OK;
    if (!recs_dev || !stride || (!counts_dev && !pos_out_dev) || (pos_out_dev && !offs_dev)) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Index23Dev id = {};
    MphfDev md;
    if (k == 23) {
        id = ix23->dev();
        md = ix23->mphf->dev();
        positions_query_kernel<23><<<aix_grid(q, 128), 128, 0, st>>>(
            id, md, (const unsigned long long *)p->indices_dev, p->n_indices, (const unsigned long long *)p->positions_dev,
            p->n_positions, recs_dev, stride, lens_dev, q, (unsigned long long *)counts_dev,
            (const unsigned long long *)offs_dev, (unsigned long long *)pos_out_dev);
    } else {
        md = ix13->mphf->dev();
        positions_query_kernel<13><<<aix_grid(q, 128), 128, 0, st>>>(
            id, md, (const unsigned long long *)p->indices_dev, p->n_indices, (const unsigned long long *)p->positions_dev,
            p->n_positions, recs_dev, stride, lens_dev, q, (unsigned long long *)counts_dev,
            (const unsigned long long *)offs_dev, (unsigned long long *)pos_out_dev);
    }
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_positions_query(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13, const aix_positions *p,
                        const uint8_t *recs, uint32_t stride, const uint8_t *lens, uint64_t q, int k, uint64_t *counts_out,
                        const uint64_t *offs, uint64_t *pos_out) {
    if (!ctx || !p) return AIX_ERR_ARG;
    if (k == 23 && !ix23) return ctx->fail(AIX_ERR_STATE, "23-mer index not loaded");
    if (k == 13 && !ix13) return ctx->fail(AIX_ERR_STATE, "13-mer index not loaded");
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (q == 0) return AIX_OK;
    if (!recs || !stride || (!counts_out && !pos_out) || (pos_out && !offs)) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    void *r_dev, *l_dev = nullptr, *c_dev = nullptr, *o_dev = nullptr, *out_dev = nullptr;
    AIX_TRY(ctx->reserve(SCR_IN0, q * stride + 64, &r_dev));
    AIX_CUDA(ctx, cudaMemcpyAsync(r_dev, recs, q * stride, cudaMemcpyHostToDevice, st));
    if (lens) {
        AIX_TRY(ctx->reserve(SCR_LEN0, q, &l_dev));
        AIX_CUDA(ctx, cudaMemcpyAsync(l_dev, lens, q, cudaMemcpyHostToDevice, st));
    }
    if (counts_out) AIX_TRY(ctx->reserve(SCR_OUT0, q * 8, &c_dev));
    uint64_t total = 0;
    if (pos_out) {
        total = offs[q];
        AIX_TRY(ctx->reserve(SCR_OUT1, (q + 1) * 8, &o_dev));
        AIX_CUDA(ctx, cudaMemcpyAsync(o_dev, offs, (q + 1) * 8, cudaMemcpyHostToDevice, st));
        AIX_TRY(ctx->reserve(SCR_TMP0, (total ? total : 1) * 8, &out_dev));
    }
    AIX_TRY(aix_positions_query_dev(ctx, ix23, ix13, p, (const uint8_t *)r_dev, stride, (const uint8_t *)l_dev, q, k,
                                    (uint64_t *)c_dev, (const uint64_t *)o_dev, (uint64_t *)out_dev));
    if (counts_out) AIX_CUDA(ctx, cudaMemcpyAsync(counts_out, c_dev, q * 8, cudaMemcpyDeviceToHost, st));
    if (pos_out && total) AIX_CUDA(ctx, cudaMemcpyAsync(pos_out, out_dev, total * 8, cudaMemcpyDeviceToHost, st));
    AIX_CUDA(ctx, cudaStreamSynchronize(st));
    return AIX_OK;
}

}  // extern "C"
