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
// GPU build = device prefix sum of tf -> ONE streaming lookup pass that emits a packed key
// (bucket << pos_bits | position) per occurrence, compacted in position order (decoupled
// look-back) -> hand-written stable LSD radix sort on the bucket bits (radix_sort.cu).  CSR
// order is bucket order and the stable sort keeps positions ascending inside a bucket, so when
// every bucket holds exactly tf occurrences (the normal case: tf was counted on the same
// reads) the sorted low words ARE positions[]: no cursor atomics, no indices[h] gather, no
// scattered 8-byte stores, no per-bucket sort.  Otherwise the sorted keys are clipped into the
// tf-sized layout (first tf kept, zero tail).
#include "aix_internal.cuh"
#include "query23.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace aix {

constexpr uint64_t kNoBucket = ~0ULL;

// per-bucket sizes for the prefix sum (tf of the index = the size of every bucket, hash.hpp:373-378)
struct TfFromRecs {
    const uint4 *recs;
    __device__ uint64_t operator()(uint64_t i) const { return recs[i].z; }
};
struct TfFromU64 {
    const uint64_t *v;
    __device__ uint64_t operator()(uint64_t i) const { return v[i]; }
};

// ---- bucket of the window starting at byte i (or kNoBucket) ---------------------------
// k = 23: hash.cpp:1006-1051.  Skip windows containing '\n', '~', 'N'; canonical form by
// numeric compare of the strict 2-bit values; the forward form is hashed from the RAW bytes
// (so a lower-case / IUPAC byte makes the verify fail), the reverse form from the decoded
// string; checker verify.
__device__ __forceinline__ uint64_t bucket23(const Index23Dev &ix, const MphfDev &m, const uint8_t *p) {
    uint64_t r0, r1, r2;
    load_window23(p, r0, r1, r2);
    bool all_acgt;
    uint64_t u, r;
    encode_validate23_rc(r0, r1, r2, all_acgt, u, r);
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
        return probe23(ix, h, us, tf, false) ? h : kNoBucket;  // raw bytes were hashed
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

// ---- emit pass: one key per occurrence, in position order ------------------------------------
constexpr int kEmitThreads = 256;
constexpr int kEmitItems = 8;
constexpr int kEmitTile = kEmitThreads * kEmitItems;  // windows per CTA

// keys[r] = (bucket << pos_bits) | (pos_base + i + 1) for the r-th window (in order of i) that has a bucket (pos_base =
// offset of this image inside the whole reads file when the file is sharded over GPUs, else 0).  The tile's
// keys are staged in shared memory, counted with ballots in window order, and the tile's first output slot comes
// from a decoupled look-back over the earlier tiles (scan.cuh).  Keys beyond `cap` are counted but not stored;
// the last tile stores the grand total in *n_valid.
template <int K>
__global__ void __launch_bounds__(kEmitThreads) positions_emit_kernel(Index23Dev ix, MphfDev m, const uint8_t *__restrict__ reads,
                                                                    uint64_t start, uint64_t n_win_end, uint64_t pos_base, int pos_bits,
                                                                    uint64_t *__restrict__ keys, uint64_t cap,
                                                                    unsigned long long *__restrict__ status,
                                                                    unsigned int *__restrict__ tile_counter,
                                                                    unsigned long long *__restrict__ n_valid) {
    __shared__ uint64_t s_key[kEmitTile];
    __shared__ uint32_t s_cnt[kEmitItems * (kEmitThreads / 32)];
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_excl;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint64_t tile = s_tile;
    const uint64_t base = start + tile * kEmitTile;
#pragma unroll 1
    for (int j = 0; j < kEmitItems; ++j) {
        const uint64_t i = base + (uint64_t)j * kEmitThreads + tid;
        uint64_t h = kNoBucket;
        if (i < n_win_end) h = K == 23 ? bucket23(ix, m, reads + i) : bucket13(m, reads + i);
        s_key[j * kEmitThreads + tid] = h == kNoBucket ? ~0ULL : ((h << pos_bits) | (pos_base + i + 1));
    }
    // every thread reads back its own keys only: no barrier needed before the ballots
    uint32_t rank[kEmitItems];
    uint32_t have = 0;
#pragma unroll
    for (int j = 0; j < kEmitItems; ++j) {
        const bool v = s_key[j * kEmitThreads + tid] != ~0ULL;
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, v);
        rank[j] = __popc(b & ((1u << lane) - 1u));
        have |= (v ? 1u : 0u) << j;
        if (lane == 0) s_cnt[j * (kEmitThreads / 32) + warp] = __popc(b);
    }
    __syncthreads();
    if (warp == 0) {  // 64 (item, warp) counts in window order -> exclusive prefix; look back for the tile's first slot
        const uint32_t v0 = s_cnt[2 * lane], v1 = s_cnt[2 * lane + 1];
        uint32_t x = v0 + v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
        s_cnt[2 * lane] = x - v0 - v1;
        s_cnt[2 * lane + 1] = x - v1;
        if (lane == 0) {
            const unsigned long long excl = lookback_exclusive(status, 1, tile, total);
            s_excl = excl;
            if (tile + 1 == gridDim.x) *n_valid = excl + total;
        }
    }
    __syncthreads();
    const unsigned long long excl = s_excl;
#pragma unroll
    for (int j = 0; j < kEmitItems; ++j) {
        if ((have >> j) & 1u) {
            const unsigned long long r = excl + s_cnt[j * (kEmitThreads / 32) + warp] + rank[j];
            if (r < cap) keys[r] = s_key[j * kEmitThreads + tid];
        }
    }
}

// normal case, one pass over the sorted keys: positions[j] = low word of key j, written to the spare buffer (the
// keys stay intact), and the test that makes this valid -- key j lies inside the bucket that owns slot slot_base + j,
// for every j; with n_valid == number of slots that means every bucket holds exactly its tf occurrences.  (slot_base
// = first slot of this GPU's bucket range when the index is built by several GPUs, else 0.)  Two keys per thread.
__global__ void __launch_bounds__(256) positions_finalize_kernel(const uint64_t *__restrict__ sorted, uint64_t n, uint64_t slot_base,
                                                               int pos_bits, const unsigned long long *__restrict__ indices,
                                                               uint64_t n_buckets, unsigned long long *__restrict__ positions,
                                                               int *__restrict__ bad) {
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (j >= n) return;
    const uint64_t mask = (1ULL << pos_bits) - 1, s0 = slot_base + j;
    bool ok = true;
    if (j + 1 < n) {
        const ulonglong2 k = __ldcs((const ulonglong2 *)(sorted + j));  // 16-byte aligned: j is even
        const uint64_t h0 = k.x >> pos_bits, h1 = k.y >> pos_bits;
        ok = h0 < n_buckets && h1 < n_buckets && s0 >= indices[h0] && s0 < indices[h0 + 1] && s0 + 1 >= indices[h1] &&
             s0 + 1 < indices[h1 + 1];
        __stcs((ulonglong2 *)(positions + j), make_ulonglong2(k.x & mask, k.y & mask));
    } else {
        const uint64_t k = sorted[j], h = k >> pos_bits;
        ok = h < n_buckets && s0 >= indices[h] && s0 < indices[h + 1];
        positions[j] = k & mask;
    }
    if (!ok) *bad = 1;
}

// general case: first sorted slot of every bucket that occurs (run_start is indexed by bucket - h_base) ...
__global__ void positions_run_start_kernel(const uint64_t *__restrict__ sorted, uint64_t n, int pos_bits, uint64_t h_base,
                                           unsigned long long *__restrict__ run_start) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t h = sorted[j] >> pos_bits;
    if (j == 0 || (sorted[j - 1] >> pos_bits) != h) run_start[h - h_base] = j;
}

// ... then the first tf[h] occurrences of every bucket go to positions[indices[h] - slot_base ...]; the rest of the
// bucket stays 0 (hash.cpp:1037-1041: slot >= tf is dropped; unfilled slots keep the zero of the allocation)
template <typename F>
__global__ void positions_clip_kernel(F tf, const uint64_t *__restrict__ sorted, uint64_t n, int pos_bits, uint64_t h_base,
                                      uint64_t slot_base, const unsigned long long *__restrict__ run_start,
                                      const unsigned long long *__restrict__ indices,
                                      unsigned long long *__restrict__ positions) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = sorted[j];
    const uint64_t h = key >> pos_bits;
    const uint64_t r = j - run_start[h - h_base];
    if (r < tf(h)) positions[indices[h] - slot_base + r] = key & ((1ULL << pos_bits) - 1);
}

// smallest bucket h with indices[h] >= target[r] (the owner boundaries of the multi-GPU build): one thread per target
__global__ void positions_split_kernel(const unsigned long long *__restrict__ indices, uint64_t n_buckets,
                                       const unsigned long long *__restrict__ target, int n_targets,
                                       unsigned long long *__restrict__ out) {
    const int r = threadIdx.x;
    if (r >= n_targets) return;
    uint64_t lo = 0, hi = n_buckets;  // indices has n_buckets + 1 entries, non-decreasing
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (indices[mid] >= target[r]) hi = mid;
        else lo = mid + 1;
    }
    out[r] = lo;
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
                if (probe23(ix, hh, us, tf, false)) h = hh;  // raw bytes were hashed
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

static int bit_length(uint64_t x) {
    int b = 0;
    while (x) { ++b; x >>= 1; }
    return b;
}

#define PB_FAIL(ctx, e__) (ctx)->fail((e__) == cudaErrorMemoryAllocation ? AIX_ERR_NOMEM : AIX_ERR_CUDA, "positions build %s:%d: %s", \
                                      __FILE__, __LINE__, cudaGetErrorString(e__))

// One key per occurrence of the windows [start, n_win_end) of an image in HBM, in position order, into *keys_out (pool
// memory, `cap` entries; *cap_io = 0 asks for the exact size: sum(tf) first, then the true count if that was too small).
template <int K>
static int emit_keys(aix_ctx *ctx, cudaStream_t st, Index23Dev id, MphfDev md, const uint8_t *reads_dev, uint64_t start, uint64_t n_win_end,
                     uint64_t pos_base, int pos_bits, uint64_t first_cap, uint64_t **keys_out, uint64_t *cap_out, uint64_t *n_valid_out) {
    *keys_out = nullptr;
    *n_valid_out = 0;
    const uint64_t n_win = n_win_end - start;
    const uint64_t e_tiles = (n_win + kEmitTile - 1) / kEmitTile;
    if (e_tiles >= (1ull << 31)) return ctx->fail(AIX_ERR_ARG, "positions build: reads image too large");
    // emit scratch: [0] n_valid, [1] tile counter, [8 ...] one look-back word per tile
    unsigned long long *scratch = nullptr;
    uint64_t *keys = nullptr;
    cudaError_t e = aix_pool_alloc(ctx, &scratch, (8 + e_tiles) * 8, st);
    if (e != cudaSuccess) { cudaGetLastError(); return PB_FAIL(ctx, e); }
    unsigned long long n_valid = 0;
    uint64_t cap = first_cap ? first_cap : 1;
    for (int attempt = 0; attempt < 2; ++attempt) {
        e = aix_pool_alloc(ctx, &keys, cap * 8, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(scratch, 0, (8 + e_tiles) * 8, st);
        if (e == cudaSuccess) {
            AixTraceSpan span(st);
            positions_emit_kernel<K><<<(unsigned)e_tiles, kEmitThreads, 0, st>>>(id, md, reads_dev, start, n_win_end, pos_base, pos_bits, keys, cap,
                                                                                scratch + 8, (unsigned int *)(scratch + 1), scratch);
            ctx->launches++;
            e = cudaGetLastError();
            span.done("positions_emit_kernel");
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&n_valid, scratch, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaGetLastError();
            aix_pool_free(ctx, keys, st); aix_pool_free(ctx, scratch, st);
            return PB_FAIL(ctx, e);
        }
        if (n_valid <= cap) break;
        // more occurrences in the reads than the index's tf sums to (index counted on other reads): all of them must be
        // ordered before the first tf of every bucket can be picked -- emit again with room for all
        aix_pool_free(ctx, keys, st);
        keys = nullptr;
        cap = n_valid;
    }
    aix_pool_free(ctx, scratch, st);
    *keys_out = keys;
    *cap_out = cap;
    *n_valid_out = n_valid;
    return AIX_OK;
}

// Sorted keys -> the slots [slot_base, slot_base + slot_count) of positions[] (the buckets [h_base, h_base + h_count)).
// `sorted` and `spare` (same capacity, may be null) are pool buffers that this function takes over: one of them becomes
// *positions_out in the normal case.
template <typename F>
static int finish_sorted(aix_ctx *ctx, cudaStream_t st, F tf, uint64_t *sorted, uint64_t *spare, uint64_t spare_cap, uint64_t n_valid,
                         uint64_t slot_base, uint64_t slot_count, uint64_t h_base, uint64_t h_count,
                         const unsigned long long *indices, uint64_t n_buckets, int pos_bits, unsigned long long **positions_out) {
    *positions_out = nullptr;
    int *bad_dev = nullptr;
    unsigned long long *positions = nullptr, *run_start = nullptr;
    auto done = [&](cudaError_t e) {
        cudaGetLastError();
        aix_pool_free(ctx, sorted, st); aix_pool_free(ctx, spare, st); aix_pool_free(ctx, bad_dev, st);
        aix_pool_free(ctx, positions, st); aix_pool_free(ctx, run_start, st);
        return PB_FAIL(ctx, e);
    };
    cudaError_t e;
    int bad = 1;
    if (n_valid == slot_count && spare && spare_cap >= slot_count && slot_count) {
        if ((e = aix_pool_alloc(ctx, &bad_dev, sizeof(int), st)) != cudaSuccess) return done(e);
        if ((e = cudaMemsetAsync(bad_dev, 0, sizeof(int), st)) != cudaSuccess) return done(e);
        positions_finalize_kernel<<<aix_grid((n_valid + 1) / 2, 256), 256, 0, st>>>(sorted, n_valid, slot_base, pos_bits, indices, n_buckets,
                                                                                 (unsigned long long *)spare, bad_dev);
        ctx->launches++;
        if ((e = cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(e);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(e);
    }
    if (!bad) {  // every bucket exactly full: the low words of the sorted keys are positions[]
        positions = (unsigned long long *)spare;
        spare = nullptr;
    } else {
        aix_pool_free(ctx, spare, st);
        spare = nullptr;
        if ((e = aix_pool_alloc(ctx, &positions, (slot_count ? slot_count : 1) * 8, st)) != cudaSuccess) return done(e);
        if ((e = cudaMemsetAsync(positions, 0, (slot_count ? slot_count : 1) * 8, st)) != cudaSuccess) return done(e);
        if (n_valid) {
            if ((e = aix_pool_alloc(ctx, &run_start, (h_count ? h_count : 1) * 8, st)) != cudaSuccess) return done(e);
            positions_run_start_kernel<<<aix_grid(n_valid, 256), 256, 0, st>>>(sorted, n_valid, pos_bits, h_base, run_start);
            positions_clip_kernel<<<aix_grid(n_valid, 256), 256, 0, st>>>(tf, sorted, n_valid, pos_bits, h_base, slot_base, run_start, indices, positions);
            ctx->launches += 2;
            if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
        }
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(e);
    }
    aix_pool_free(ctx, sorted, st); aix_pool_free(ctx, bad_dev, st); aix_pool_free(ctx, run_start, st);
    *positions_out = positions;
    return AIX_OK;
}

// Device part of the build.  reads_dev must be readable for 8 bytes past len (aligned word
// loads of the last windows); `start` = first_start of the image.  On success *indices_dev
// (u64[n+1]) and *positions_dev (u64[total], at least one element) are owned by the caller (pool memory).
template <int K, typename F>
static int build_core(aix_ctx *ctx, Index23Dev id, MphfDev md, F tf, uint64_t n, const uint8_t *reads_dev, uint64_t len,
                      uint64_t start, unsigned long long **indices_dev, unsigned long long **positions_dev,
                      uint64_t *total_out) {
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->l2_unpin();
    cudaStream_t st = ctx->stream;
    unsigned long long *indices = nullptr, *positions = nullptr, *tiles = nullptr;
    cudaError_t e = aix_pool_alloc(ctx, &indices, (n + 1) * 8, st);
    if (e == cudaSuccess) e = aix_pool_alloc(ctx, &tiles, scan_scratch_bytes(n), st);
    auto bail = [&](int rc) {
        aix_pool_free(ctx, indices, st); aix_pool_free(ctx, tiles, st);
        return rc;
    };
    if (e != cudaSuccess) { cudaGetLastError(); return bail(PB_FAIL(ctx, e)); }
    AixTrace trace(st, "positions build");
    int rc = exclusive_scan(ctx, st, tf, n, indices, tiles);
    if (rc != AIX_OK) return bail(rc);
    unsigned long long total = 0;
    e = cudaMemcpyAsync(&total, indices + n, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaGetLastError(); return bail(PB_FAIL(ctx, e)); }
    aix_pool_free(ctx, tiles, st);
    tiles = nullptr;
    trace.mark("prefix sum of tf");
    const uint64_t n_win_end = len >= (uint64_t)K ? len - K + 1 : 0;
    const bool any_window = n && total && start < n_win_end;
    if (!any_window) {
        e = aix_pool_alloc(ctx, &positions, (total ? total : 1) * 8, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(positions, 0, (total ? total : 1) * 8, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaGetLastError(); aix_pool_free(ctx, positions, st); return bail(PB_FAIL(ctx, e)); }
    } else {
        // key = bucket << pos_bits | position (1-based byte offset <= len)
        const int pos_bits = bit_length(len), h_bits = bit_length(n - 1) ? bit_length(n - 1) : 1;
        if (pos_bits + h_bits > 64)
            return bail(ctx->fail(AIX_ERR_ARG, "positions build: %d position bits + %d bucket bits do not fit a 64-bit key", pos_bits, h_bits));
        uint64_t *keys = nullptr, *alt = nullptr, cap = 0, n_valid = 0;
        // the normal case needs exactly `total` keys; more valid windows than that = general case (second attempt inside)
        rc = emit_keys<K>(ctx, st, id, md, reads_dev, start, n_win_end, 0, pos_bits, total, &keys, &cap, &n_valid);
        if (rc != AIX_OK) return bail(rc);
        trace.mark("emit pass (lookup + ordered compaction of one key per occurrence)");
        e = aix_pool_alloc(ctx, &alt, cap * 8, st);
        if (e != cudaSuccess) { cudaGetLastError(); aix_pool_free(ctx, keys, st); return bail(PB_FAIL(ctx, e)); }
        uint64_t *sorted = keys;
        rc = radix_sort_u64(ctx, st, keys, alt, n_valid, pos_bits, pos_bits + h_bits, &sorted);
        if (rc != AIX_OK) { aix_pool_free(ctx, keys, st); aix_pool_free(ctx, alt, st); return bail(rc); }
        trace.mark("radix sort on the bucket bits");
        uint64_t *spare = sorted == keys ? alt : keys;
        rc = finish_sorted(ctx, st, tf, sorted, spare, cap, n_valid, 0, total, 0, n, indices, n, pos_bits, &positions);
        if (rc != AIX_OK) return bail(rc);
        trace.mark("check + low words -> positions[] (or clip to tf per bucket)");
    }
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
    aix_pool_free(ctx, indices, ctx->stream);
    aix_pool_free(ctx, positions, ctx->stream);
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
    p->pooled = true;
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
    return build_impl<23>(ctx, ix->dev(), ix->mphf_dev(), TfFromRecs{ix->recs_dev}, ix->n, reads, len, indices_out, positions_out);
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
    return build_dev_impl<23>(ctx, ix->dev(), ix->mphf_dev(), TfFromRecs{ix->recs_dev}, ix->n, reads_dev, len, out);
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
    if (p->pooled && ctx) {
        aix_pool_free(ctx, p->indices_dev, ctx->stream);
        aix_pool_free(ctx, p->positions_dev, ctx->stream);
    } else {
        if (p->indices_dev) cudaFree(p->indices_dev);  // cudaFree also takes pool memory (it synchronises first)
        if (p->positions_dev) cudaFree(p->positions_dev);
    }
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
        md = ix23->mphf_dev();
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

// ---- positions build over several GPUs of one box ------------------------------------------------
// Reference: fill_index_from_reads splits the byte range of the reads file over worker threads (hash.hpp:407-444), all
// of them scattering into one positions array.  Here the workers are GPUs: GPU r looks up the windows that start in its
// byte range (index replicated) and emits packed keys in position order; the buckets are cut into one contiguous range
// per GPU with equal numbers of slots; every GPU partitions its keys by owner (one stable pass, radix_sort.cu) and
// copies each part straight into the owner's buffer over NVLink (part of GPU r before part of GPU r + 1: positions stay
// ascending); the owner sorts what it received on the bucket bits and owns that slice of positions[].
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0;
    uint64_t gen = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

double ms_since(double t0) { return (AixTrace::now() - t0) * 1e3; }

}  // namespace

extern "C" int aix_positions_build23_multi(aix_multi *mg, const aix_index23 *const *ix, const uint8_t *reads, uint64_t len,
                                           uint64_t *indices_out, uint64_t *positions_out, aix_multi_build_stats *stats) {
    if (!mg || !ix || !indices_out || (len && !reads) || mg->ctx.empty()) return AIX_ERR_ARG;
    const int n = (int)mg->ctx.size();
    if (n > 16) { mg->err = "at most 16 GPUs"; return AIX_ERR_ARG; }
    for (int r = 0; r < n; ++r)
        if (!ix[r] || ix[r]->n != ix[0]->n) { mg->err = "every GPU needs its own upload of the same index"; return AIX_ERR_ARG; }
    constexpr int K = 23;
    const uint64_t nb = ix[0]->n;
    const uint64_t n_win_end = len >= (uint64_t)K ? len - K + 1 : 0;
    const uint64_t first = first_start(reads, len, K);
    const int pos_bits = bit_length(len) ? bit_length(len) : 1, h_bits = bit_length(nb ? nb - 1 : 0) ? bit_length(nb ? nb - 1 : 0) : 1;
    if (pos_bits + h_bits > 64) { mg->err = "position bits + bucket bits do not fit a 64-bit key"; return AIX_ERR_ARG; }
    // byte ranges: windows that START in [c[r], c[r+1]) belong to GPU r (it also gets the K - 1 bytes that follow)
    std::vector<uint64_t> c(n + 1);
    for (int r = 0; r <= n; ++r) c[r] = first + (n_win_end > first ? (n_win_end - first) * (uint64_t)r / (uint64_t)n : 0);
    struct PerGpu {
        uint8_t *reads_dev = nullptr;
        unsigned long long *indices = nullptr, *positions = nullptr;
        uint64_t *keys = nullptr, *part = nullptr, *recv = nullptr, *alt = nullptr;
        uint64_t n_valid = 0, recv_total = 0;
    };
    std::vector<PerGpu> g(n);
    std::vector<uint64_t> counts((size_t)n * n, 0), bounds(n + 1, 0), slot_lo(n + 1, 0);
    unsigned long long total = 0;
    std::vector<int> rcs(n, AIX_OK);
    std::vector<double> t_emit(n, 0), t_exch(n, 0), t_sort(n, 0), t_up(n, 0), t_down(n, 0), t_alloc(n, 0);
    HostBarrier bar(n);
    auto any_failed = [&]() { for (int r = 0; r < n; ++r) if (rcs[r] != AIX_OK) return true; return false; };
    const double t_all = AixTrace::now();
    auto worker = [&](int r) {
        aix_ctx *ctx = mg->ctx[r];
        cudaStream_t st = ctx->stream;
        PerGpu &me = g[r];
        const char *phase = "start";
        auto fail_cuda = [&](cudaError_t e) {
            cudaGetLastError();
            size_t fr = 0, to = 0;
            cudaMemGetInfo(&fr, &to);
            rcs[r] = ctx->fail(e == cudaErrorMemoryAllocation ? AIX_ERR_NOMEM : AIX_ERR_CUDA,
                               "multi-GPU positions build, GPU %d of %d, phase %s: %s (%.1f of %.1f GB free)", r, n, phase,
                               cudaGetErrorString(e), fr / 1e9, to / 1e9);
        };
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) fail_cuda(e);
        const uint64_t lo = c[r], hi = c[r + 1];
        const uint64_t img_len = hi > lo ? (hi - lo) + K - 1 : 0;   // bytes [lo, hi + K - 1) <= len
        double t0 = AixTrace::now();
        // ---- phase 1: upload the shard, prefix sum of tf (every GPU keeps the offsets: its finalize pass needs them)
        phase = "upload + prefix sum";
        if (rcs[r] == AIX_OK) {
            e = aix_pool_alloc(ctx, &me.reads_dev, img_len + 64, st);
            if (e == cudaSuccess && img_len) e = cudaMemcpyAsync(me.reads_dev, reads + lo, img_len, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(me.reads_dev + img_len, '\n', 64, st);
            unsigned long long *tiles = nullptr;
            if (e == cudaSuccess) e = aix_pool_alloc(ctx, &me.indices, (nb + 1) * 8, st);
            if (e == cudaSuccess) e = aix_pool_alloc(ctx, &tiles, scan_scratch_bytes(nb), st);
            if (e != cudaSuccess) fail_cuda(e);
            else {
                int rc = exclusive_scan(ctx, st, TfFromRecs{ix[r]->recs_dev}, nb, me.indices, tiles);
                if (rc != AIX_OK) rcs[r] = rc;
            }
            aix_pool_free(ctx, tiles, st);
            if (rcs[r] == AIX_OK && r == 0) {
                // owner boundaries: equal numbers of slots per GPU
                e = cudaMemcpyAsync(&total, me.indices + nb, 8, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                unsigned long long *tgt = nullptr;
                std::vector<unsigned long long> h_t(2 * (n + 1));
                for (int o = 0; o <= n; ++o) h_t[o] = (unsigned long long)((unsigned __int128)total * (unsigned)o / (unsigned)n);
                if (e == cudaSuccess) e = aix_pool_alloc(ctx, &tgt, 2 * (n + 1) * 8, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(tgt, h_t.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) {
                    positions_split_kernel<<<1, 32, 0, st>>>(me.indices, nb, tgt, n + 1, tgt + (n + 1));
                    ctx->launches++;
                    e = cudaMemcpyAsync(h_t.data() + (n + 1), tgt + (n + 1), (n + 1) * 8, cudaMemcpyDeviceToHost, st);
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e == cudaSuccess) {
                    for (int o = 0; o <= n; ++o) bounds[o] = h_t[n + 1 + o];
                    bounds[0] = 0;
                    bounds[n] = nb;
                    for (int o = 1; o < n; ++o) if (bounds[o] < bounds[o - 1]) bounds[o] = bounds[o - 1];
                    for (int o = 0; o <= n && e == cudaSuccess; ++o)
                        e = cudaMemcpyAsync(&slot_lo[o], me.indices + bounds[o], 8, cudaMemcpyDeviceToHost, st);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                }
                aix_pool_free(ctx, tgt, st);
                if (e != cudaSuccess) fail_cuda(e);
            }
            if (rcs[r] == AIX_OK && (e = cudaStreamSynchronize(st)) != cudaSuccess) fail_cuda(e);
        }
        t_up[r] = ms_since(t0);
        bar.wait();
        // ---- phase 2: emit keys in position order, partition them by owner
        phase = "emit + partition";
        t0 = AixTrace::now();
        if (!any_failed() && hi > lo && nb && total) {
            uint64_t cap = 0;
            int rc = emit_keys<K>(ctx, st, ix[r]->dev(), ix[r]->mphf_dev(), me.reads_dev, 0, hi - lo, lo, pos_bits, hi - lo, &me.keys, &cap, &me.n_valid);
            if (rc == AIX_OK) {
                const double ta = AixTrace::now();
                e = aix_plain_alloc(ctx, &me.part, (me.n_valid ? me.n_valid : 1) * 8);  // read by the other GPUs' copy engines
                t_alloc[r] += ms_since(ta);
                if (e != cudaSuccess) fail_cuda(e);
                else {
                    std::vector<uint64_t> kb(n);
                    for (int o = 0; o < n; ++o) kb[o] = bounds[o] << pos_bits;
                    rc = partition_by_range(ctx, st, me.keys, me.part, me.n_valid, kb.data(), n, &counts[(size_t)r * n]);
                }
            }
            if (rc != AIX_OK && rcs[r] == AIX_OK) rcs[r] = rc;
            aix_pool_free(ctx, me.keys, st);
            me.keys = nullptr;
        }
        aix_pool_free(ctx, me.reads_dev, st);
        me.reads_dev = nullptr;
        t_emit[r] = ms_since(t0) - t_alloc[r];
        bar.wait();
        // ---- phase 3: every owner makes room for what it will receive
        phase = "receive buffers";
        t0 = AixTrace::now();
        if (!any_failed()) {
            for (int s2 = 0; s2 < n; ++s2) me.recv_total += counts[(size_t)s2 * n + r];
            if (me.recv_total > total + (n_win_end - first)) {  // cannot be: more keys than windows
                rcs[r] = ctx->fail(AIX_ERR_STATE, "multi-GPU positions build: GPU %d would receive %llu keys of %llu windows", r,
                                   (unsigned long long)me.recv_total, (unsigned long long)(n_win_end - first));
            } else {
                static char what[3][64];
                snprintf(what[0], 64, "receive buffer (%llu keys)", (unsigned long long)me.recv_total);
                phase = what[0];
                e = aix_plain_alloc(ctx, &me.recv, (me.recv_total ? me.recv_total : 1) * 8);  // written by the other GPUs
                if (e == cudaSuccess) { phase = "receive buffer: spare"; e = aix_plain_alloc(ctx, &me.alt, (me.recv_total ? me.recv_total : 1) * 8); }
                if (e == cudaSuccess) { phase = "receive buffer: sync"; e = cudaStreamSynchronize(st); }
                if (e != cudaSuccess) fail_cuda(e);
            }
        }
        t_alloc[r] += ms_since(t0);
        bar.wait();
        // ---- phase 4: the exchange -- part o of this GPU goes to owner o, behind the parts of the GPUs before this one.
        // Step i sends to owner (r + i) mod n: every step is a permutation, so no GPU receives from two peers at once
        // (all GPUs starting with owner 0 would share its 900 GB/s of NVLink ingress).
        phase = "exchange";
        t0 = AixTrace::now();
        if (!any_failed()) {
            std::vector<uint64_t> seg(n + 1, 0);
            for (int o = 0; o < n; ++o) seg[o + 1] = seg[o] + counts[(size_t)r * n + o];
            for (int i = 0; i < n && rcs[r] == AIX_OK; ++i) {
                const int o = (r + i) % n;
                const uint64_t cnt = counts[(size_t)r * n + o];
                uint64_t off = 0;
                for (int s2 = 0; s2 < r; ++s2) off += counts[(size_t)s2 * n + o];
                if (cnt) {
                    e = o == r ? cudaMemcpyAsync(g[o].recv + off, me.part + seg[o], cnt * 8, cudaMemcpyDeviceToDevice, st)
                               : cudaMemcpyPeerAsync(g[o].recv + off, mg->ctx[o]->device, me.part + seg[o], ctx->device, cnt * 8, st);
                    if (e != cudaSuccess) fail_cuda(e);
                }
            }
            if (rcs[r] == AIX_OK && (e = cudaStreamSynchronize(st)) != cudaSuccess) fail_cuda(e);
        }
        t_exch[r] = ms_since(t0);
        aix_pool_free(ctx, me.part, st);
        me.part = nullptr;
        bar.wait();
        // ---- phase 5: sort the received keys on the bucket bits, turn them into this GPU's slice of positions[]
        phase = "sort + finalize";
        t0 = AixTrace::now();
        if (!any_failed()) {
            uint64_t *sorted = me.recv;
            int rc = radix_sort_u64(ctx, st, me.recv, me.alt, me.recv_total, pos_bits, pos_bits + h_bits, &sorted);
            if (rc == AIX_OK) {
                uint64_t *spare = sorted == me.recv ? me.alt : me.recv;
                me.recv = me.alt = nullptr;  // finish_sorted takes both over
                rc = finish_sorted(ctx, st, TfFromRecs{ix[r]->recs_dev}, sorted, spare, me.recv_total ? me.recv_total : 1, me.recv_total,
                                   slot_lo[r], slot_lo[r + 1] - slot_lo[r], bounds[r], bounds[r + 1] - bounds[r], me.indices, nb, pos_bits,
                                   &me.positions);
            }
            if (rc != AIX_OK) rcs[r] = rc;
        }
        t_sort[r] = ms_since(t0);
        // ---- phase 6: download the slice (and the offsets, from GPU 0)
        phase = "download";
        t0 = AixTrace::now();
        if (rcs[r] == AIX_OK && !any_failed()) {
            const uint64_t cnt = slot_lo[r + 1] - slot_lo[r];
            e = cudaSuccess;
            if (cnt && positions_out) e = cudaMemcpyAsync(positions_out + slot_lo[r], me.positions, cnt * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && r == 0) e = cudaMemcpyAsync(indices_out, me.indices, (nb + 1) * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) fail_cuda(e);
        }
        t_down[r] = ms_since(t0);
        aix_pool_free(ctx, me.recv, st); aix_pool_free(ctx, me.alt, st);
        aix_pool_free(ctx, me.positions, st); aix_pool_free(ctx, me.indices, st);
        cudaStreamSynchronize(st);
        bar.wait();
    };
    {
        std::vector<std::thread> th;
        for (int r = 0; r < n; ++r) th.emplace_back(worker, r);
        for (auto &x : th) x.join();
    }
    for (int r = 0; r < n; ++r)
        if (rcs[r] != AIX_OK) { mg->err = aix_last_error(mg->ctx[r]); return rcs[r]; }
    if (stats) {
        auto mx = [&](const std::vector<double> &v) { double m = 0; for (double x : v) m = x > m ? x : m; return m; };
        stats->total_ms = ms_since(t_all);
        stats->upload_scan_ms = mx(t_up); stats->emit_partition_ms = mx(t_emit); stats->exchange_ms = mx(t_exch);
        stats->sort_finalize_ms = mx(t_sort); stats->download_ms = mx(t_down); stats->alloc_ms = mx(t_alloc);
        uint64_t moved = 0, all = 0;
        for (int r = 0; r < n; ++r)
            for (int o = 0; o < n; ++o) { all += counts[(size_t)r * n + o]; if (o != r) moved += counts[(size_t)r * n + o]; }
        stats->keys = all;
        stats->peer_bytes = moved * 8;
        stats->positions = total;
    }
    return AIX_OK;
}
