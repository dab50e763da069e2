// mphf_build.cu -- index construction on the GPU: canonical 23-mer table of a read set
// (sort + run-length), emphf-format MPHF construction by parallel hypergraph peeling, and
// the checker/tf fill.
//
// Reference: emphf mphf ctor src/emphf/mphf.hpp:22-67 (gamma 1.23, hash_domain, seed
// sequence std::mt19937_64(37), value assignment :52-64), hypergraph_sorter_seq.hpp:29-95
// (sequential peeling), ranked_bitpair_vector.hpp:17-31 (block ranks);
// index fill src/hash.cpp:671-723 (worker_for_fill_index), src/compute_index.cpp:53-67;
// canonical k-mer definition tests/analyze_kmers.py:25-33.
//
// The lookup side only needs *a* valid assignment, so the peeling order is free: the GPU
// peels all degree-1 vertices of a round in parallel (claim the edge with a CAS, then
// retract it from its three vertices), records the rounds, and assigns the 2-bit values
// round by round in reverse.  Both loops run inside one cooperative kernel each.
#include <cooperative_groups.h>
#include <random>

#include "aix_internal.cuh"
#include "radix_sort.cuh"

namespace cg = cooperative_groups;

namespace aix {

constexpr uint32_t kNoVertex = 0xFFFFFFFFu;
constexpr uint32_t kMaxRounds = 1u << 20;

struct PeelState {
    uint64_t n;              // edges (keys)
    uint64_t n_nodes;        // 3 * hash_domain
    uint64_t hash_domain;
    const uint32_t *edges;   // 3 per key
    unsigned long long *node;        // (sum of incident edge ids) << 24 | degree
    uint32_t *peel_vertex;   // per edge: vertex that peeled it
    uint32_t *peel_order;    // edges in peeling order
    uint32_t *frontier[2];
    unsigned long long *counters;    // [0] n_peeled, [1] fsize0, [2] fsize1, [3] n_rounds, [4] overflow
    uint32_t *round_end;     // peel_order boundary after each round
    unsigned long long *bv;  // bit-pair vector words
};

template <int K>
__global__ void edges_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, uint32_t *__restrict__ edges,
                             unsigned long long *__restrict__ node) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t a, b, c;
    if (K == 23) {
        uint64_t w0, w1, w2;
        ascii_words23_from_rc(revcomp23(kmers[i]), w0, w1, w2);
        jenkins_short(m.seed, w0, w1, w2, 23u, a, b, c);
    } else {
        uint64_t w0, w1;
        ascii_words13_from_rc(revcomp13((uint32_t)kmers[i]), w0, w1);
        jenkins_short(m.seed, w0, w1, 0, 13u, a, b, c);
    }
    const uint64_t d = m.hash_domain;
    uint32_t v[3] = {(uint32_t)fastmod(a, d, m.magic), (uint32_t)(d + fastmod(b, d, m.magic)),
                     (uint32_t)(2 * d + fastmod(c, d, m.magic))};
    const unsigned long long inc = ((unsigned long long)i << 24) | 1ull;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        edges[3 * i + j] = v[j];
        atomicAdd(node + v[j], inc);
    }
}

__global__ void peel_kernel(PeelState s) {
    cg::grid_group grid = cg::this_grid();
    const uint64_t tid = grid.thread_rank(), nth = grid.size();
    for (uint64_t v = tid; v < s.n_nodes; v += nth) {
        if ((s.node[v] & 0xFFFFFFull) == 1ull) {
            unsigned long long idx = atomicAdd(s.counters + 1, 1ull);
            s.frontier[0][idx] = (uint32_t)v;
        }
    }
    grid.sync();
    int cur = 0;
    uint32_t round = 0;
    uint64_t peeled_begin = 0;
    while (true) {
        const uint64_t fs = *(volatile unsigned long long *)(s.counters + 1 + cur);
        if (fs == 0 || round >= kMaxRounds) break;
        if (tid == 0) s.counters[1 + (cur ^ 1)] = 0;
        // phase A: every degree-1 vertex of the frontier claims its last edge
        for (uint64_t i = tid; i < fs; i += nth) {
            uint32_t v = s.frontier[cur][i];
            unsigned long long w = s.node[v];
            if ((w & 0xFFFFFFull) == 1ull) {
                uint32_t e = (uint32_t)(w >> 24);
                if (atomicCAS(s.peel_vertex + e, kNoVertex, v) == kNoVertex) {
                    unsigned long long pos = atomicAdd(s.counters + 0, 1ull);
                    s.peel_order[pos] = e;
                }
            }
        }
        grid.sync();
        const uint64_t peeled_end = *(volatile unsigned long long *)(s.counters + 0);
        if (tid == 0) s.round_end[round] = (uint32_t)peeled_end;
        // phase B: retract the claimed edges; vertices dropping to degree 1 form the next frontier
        for (uint64_t i = peeled_begin + tid; i < peeled_end; i += nth) {
            uint32_t e = s.peel_order[i];
            const unsigned long long dec = ((unsigned long long)e << 24) | 1ull;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint32_t u = s.edges[3 * (uint64_t)e + j];
                unsigned long long old = atomicAdd(s.node + u, 0ull - dec);
                if ((old & 0xFFFFFFull) == 2ull) {
                    unsigned long long idx = atomicAdd(s.counters + 1 + (cur ^ 1), 1ull);
                    s.frontier[cur ^ 1][idx] = u;
                }
            }
        }
        grid.sync();
        peeled_begin = peeled_end;
        cur ^= 1;
        ++round;
    }
    if (tid == 0) s.counters[3] = round;
}

// reverse peeling order, one round at a time (mphf.hpp:52-64)
__global__ void assign_kernel(PeelState s, uint32_t n_rounds) {
    cg::grid_group grid = cg::this_grid();
    const uint64_t tid = grid.thread_rank(), nth = grid.size();
    for (uint32_t r = n_rounds; r-- > 0;) {
        const uint64_t begin = r ? s.round_end[r - 1] : 0, end = s.round_end[r];
        for (uint64_t i = begin + tid; i < end; i += nth) {
            uint32_t e = s.peel_order[i];
            uint32_t v0 = s.peel_vertex[e];
            uint32_t target = (uint32_t)(v0 / s.hash_domain);  // orientation: which of the three nodes
            uint32_t assigned = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint32_t u = s.edges[3 * (uint64_t)e + j];
                if (u != v0) assigned += (uint32_t)((*(volatile unsigned long long *)(s.bv + (u >> 5)) >> ((u & 31) * 2)) & 3ull);
            }
            uint32_t val = (target + 9u - assigned) % 3u;
            if (val == 0) val = 3;  // assigned values must be non-zero to be ranked
            atomicOr(s.bv + (v0 >> 5), (unsigned long long)val << ((v0 & 31) * 2));
        }
        grid.sync();
    }
}

// block_cnt[b] = non-zero pairs of words [16b, 16b+16)
__global__ void block_count_kernel(const unsigned long long *__restrict__ bv, uint64_t n_words, uint64_t n_blocks,
                                   uint64_t *__restrict__ block_cnt) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    uint64_t c = 0;
    for (uint64_t w = b * 16; w < b * 16 + 16 && w < n_words; ++w) c += nonzero_pairs64(bv[w]);
    block_cnt[b] = c;
}

// exclusive scan in place, single CTA (n_blocks = n_nodes / 512 is small)
__global__ void block_scan_kernel(uint64_t *__restrict__ v, uint64_t n) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (uint64_t i0 = 0; i0 < n; i0 += blockDim.x) {
        uint64_t i = i0 + threadIdx.x;
        unsigned long long x0 = i < n ? v[i] : 0ull, x = x0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long t = lane < nw ? warp_tot[lane] : 0ull;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o);
                if (lane >= (unsigned)o) t += y;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        unsigned long long base = carry + (wid ? warp_tot[wid - 1] : 0ull);
        if (i < n) v[i] = base + x - x0;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[nw - 1];
        __syncthreads();
    }
}

// ---- index fill -------------------------------------------------------------------------
__global__ void index_fill_kernel(MphfDev m, const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ counts,
                                  uint64_t n, uint64_t *__restrict__ checker, uint32_t *__restrict__ tf,
                                  unsigned long long *__restrict__ bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = kmers[i];
    uint64_t h = mphf_lookup23(m, revcomp23(k));
    if (h >= n) {
        atomicAdd(bad, 1ull);
        return;
    }
    checker[h] = k;  // hash.cpp:718-719
    tf[h] = counts[i];
}

// ---- canonical 23-mers of a reads buffer -----------------------------------------------
// one thread per 16 window starts; invalid windows emit the all-ones sentinel (sorts last)
constexpr int kCanRoll = 16;
__global__ void canonical23_kernel(const uint8_t *__restrict__ bytes, uint64_t len, uint64_t *__restrict__ out) {
    const uint64_t n_win = len - 22;
    uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kCanRoll;
    if (i0 >= n_win) return;
    const uint64_t mask = (1ULL << 46) - 1;
    uint64_t f = 0, r = 0;
    uint32_t bad = 0;
    for (int j = 0; j < 22; ++j) {
        uint32_t ch = bytes[i0 + j];
        uint64_t c = base_code_strict(ch);
        f = (f << 2) | c;
        r = (r >> 2) | ((3 - c) << 44);
        bad = (bad << 1) | (is_acgt_upper(ch) ? 0u : 1u);
    }
    for (int t = 0; t < kCanRoll && i0 + t < n_win; ++t) {
        uint32_t ch = bytes[i0 + t + 22];
        uint64_t c = base_code_strict(ch);
        f = ((f << 2) | c) & mask;
        r = (r >> 2) | ((3 - c) << 44);
        bad = ((bad << 1) | (is_acgt_upper(ch) ? 0u : 1u)) & ((1u << 23) - 1);
        out[i0 + t] = bad ? ~0ULL : (f <= r ? f : r);
    }
}

// Inputs with 2^31 or more windows (or more than the HBM at hand can sort at once) are
// processed in passes over ranges of the canonical value: pass A counts the k-mers per
// top-12-bit bin, the host groups bins into ranges that fit, pass B emits the k-mers of one
// range compactly (order is irrelevant: they are sorted next).
constexpr int kCanBinsLog2 = 12;
constexpr int kCanBins = 1 << kCanBinsLog2;

// the 16 canonical k-mers of a thread (all-ones = invalid window or past the end)
__device__ __forceinline__ void canonical23_roll(const uint8_t *__restrict__ bytes, uint64_t n_win, uint64_t i0,
                                                 uint64_t (&km)[kCanRoll]) {
    const uint64_t mask = (1ULL << 46) - 1;
    uint64_t f = 0, r = 0;
    uint32_t bad = 0;
    for (int j = 0; j < 22; ++j) {
        uint32_t ch = bytes[i0 + j];
        uint64_t c = base_code_strict(ch);
        f = (f << 2) | c;
        r = (r >> 2) | ((3 - c) << 44);
        bad = (bad << 1) | (is_acgt_upper(ch) ? 0u : 1u);
    }
#pragma unroll
    for (int t = 0; t < kCanRoll; ++t) {
        km[t] = ~0ULL;
        if (i0 + t < n_win) {
            uint32_t ch = bytes[i0 + t + 22];
            uint64_t c = base_code_strict(ch);
            f = ((f << 2) | c) & mask;
            r = (r >> 2) | ((3 - c) << 44);
            bad = ((bad << 1) | (is_acgt_upper(ch) ? 0u : 1u)) & ((1u << 23) - 1);
            if (!bad) km[t] = f <= r ? f : r;
        }
    }
}

__global__ void __launch_bounds__(128) canonical23_bins_kernel(const uint8_t *__restrict__ bytes, uint64_t len,
                                                             unsigned long long *__restrict__ bins) {
    __shared__ uint32_t sbin[kCanBins];
    for (int i = threadIdx.x; i < kCanBins; i += blockDim.x) sbin[i] = 0;
    __syncthreads();
    const uint64_t n_win = len - 22;
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kCanRoll;
    if (i0 < n_win) {
        uint64_t km[kCanRoll];
        canonical23_roll(bytes, n_win, i0, km);
#pragma unroll
        for (int t = 0; t < kCanRoll; ++t)
            if (km[t] != ~0ULL) atomicAdd(&sbin[(uint32_t)(km[t] >> (46 - kCanBinsLog2))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kCanBins; i += blockDim.x)
        if (sbin[i]) atomicAdd(bins + i, (unsigned long long)sbin[i]);
}

// k-mers with lo <= value < hi, appended to out[*cursor ...] (one atomic per warp)
__global__ void __launch_bounds__(128) canonical23_range_kernel(const uint8_t *__restrict__ bytes, uint64_t len, uint64_t lo,
                                                              uint64_t hi, uint64_t *__restrict__ out,
                                                              unsigned long long *__restrict__ cursor) {
    const uint64_t n_win = len - 22;
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kCanRoll;
    uint64_t km[kCanRoll];
    uint32_t c = 0;
    if (i0 < n_win) {
        canonical23_roll(bytes, n_win, i0, km);
#pragma unroll
        for (int t = 0; t < kCanRoll; ++t) c += (km[t] >= lo && km[t] < hi) ? 1u : 0u;
    }
    const unsigned lane = threadIdx.x & 31u;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, x, 31);
    unsigned long long base = 0;
    if (lane == 31 && warp_total) base = atomicAdd(cursor, (unsigned long long)warp_total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    if (c) {
        uint64_t w = base + (x - c);
#pragma unroll
        for (int t = 0; t < kCanRoll; ++t)
            if (km[t] >= lo && km[t] < hi) out[w++] = km[t];
    }
}

}  // namespace aix

using namespace aix;

static int build_impl(aix_ctx *ctx, const uint64_t *kmers_dev, uint64_t n, int k, aix_mphf **out) {
    *out = nullptr;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->l2_unpin();
    // mphf.hpp:27: hash_domain = (ceil(n * gamma) + 2) / 3
    const uint64_t hash_domain = ((uint64_t)std::ceil((double)n * 1.23) + 2) / 3;
    const uint64_t n_nodes = 3 * hash_domain;
    if (n_nodes >= 0xFFFFFFFFull || n >= (1ull << 32))
        return ctx->fail(AIX_ERR_ARG, "mphf build: %llu keys exceed the 32-bit node range", (unsigned long long)n);
    aix_mphf *m = new aix_mphf();
    m->n = n; m->hash_domain = hash_domain; m->bv_size = n_nodes;
    m->n_words = (n_nodes + 31) / 32;
    m->n_blocks = (n_nodes + 511) / 512;
    m->words.assign(m->n_words, 0);
    m->block_ranks.assign(m->n_blocks, 0);
    if (n == 0) {  // empty key set: the reference writes a header-only hasher
        std::mt19937_64 rng(37);
        m->seed = rng();
        int rc = mphf_build_layout(ctx, m);
        if (rc != AIX_OK) { aix_mphf_destroy(ctx, m); return rc; }
        *out = m;
        return AIX_OK;
    }

    PeelState s = {};
    s.n = n; s.n_nodes = n_nodes; s.hash_domain = hash_domain;
    uint32_t *edges = nullptr;
    uint64_t *block_cnt = nullptr;
    auto cleanup = [&]() {
        cudaFree(edges); cudaFree(s.node); cudaFree(s.peel_vertex); cudaFree(s.peel_order);
        cudaFree(s.frontier[0]); cudaFree(s.frontier[1]); cudaFree(s.counters); cudaFree(s.round_end);
        cudaFree(s.bv); cudaFree(block_cnt);
    };
    cudaError_t e = cudaMalloc(&edges, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&s.node, n_nodes * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.peel_vertex, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.peel_order, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.frontier[0], n_nodes * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.frontier[1], n_nodes * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.counters, 8 * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.round_end, (size_t)kMaxRounds * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.bv, m->n_words * 8);
    if (e == cudaSuccess) e = cudaMalloc(&block_cnt, m->n_blocks * 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cleanup();
        delete m;
        return ctx->fail(AIX_ERR_NOMEM, "mphf build buffers: %s", cudaGetErrorString(e));
    }
    s.edges = edges;

    int blocks_per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, peel_kernel, 256, 0);
    int bps2 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps2, assign_kernel, 256, 0);
    if (bps2 < blocks_per_sm) blocks_per_sm = bps2;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (blocks_per_sm > 4) blocks_per_sm = 4;
    const unsigned coop_grid = (unsigned)(ctx->sm_count * blocks_per_sm);

    std::mt19937_64 rng(37);  // mphf.hpp:45
    int rc = AIX_ERR_BUILD;
    cudaStream_t st = ctx->stream;
    for (int trial = 0; trial < 64; ++trial) {
        m->seed = rng();  // BaseHasher::generate(rng), base_hash.hpp:30-34
        MphfDev md = m->dev();
        cudaMemsetAsync(s.node, 0, n_nodes * 8, st);
        cudaMemsetAsync(s.peel_vertex, 0xFF, n * 4, st);
        cudaMemsetAsync(s.counters, 0, 8 * 8, st);
        cudaMemsetAsync(s.bv, 0, m->n_words * 8, st);
        if (k == 23) edges_kernel<23><<<aix_grid(n, 256), 256, 0, st>>>(md, kmers_dev, n, edges, s.node);
        else edges_kernel<13><<<aix_grid(n, 256), 256, 0, st>>>(md, kmers_dev, n, edges, s.node);
        ctx->launches++;
        void *args[] = {&s};
        e = cudaLaunchCooperativeKernel((void *)peel_kernel, dim3(coop_grid), dim3(256), args, 0, st);
        ctx->launches++;
        unsigned long long counters[8];
        if (e == cudaSuccess) e = cudaMemcpyAsync(counters, s.counters, sizeof counters, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            rc = ctx->fail(AIX_ERR_CUDA, "mphf peel: %s", cudaGetErrorString(e));
            break;
        }
        if (counters[0] < n) continue;  // not peelable with this seed: next trial (mphf.hpp:47-51)
        uint32_t n_rounds = (uint32_t)counters[3];
        void *args2[] = {&s, &n_rounds};
        e = cudaLaunchCooperativeKernel((void *)assign_kernel, dim3(coop_grid), dim3(256), args2, 0, st);
        ctx->launches++;
        block_count_kernel<<<aix_grid(m->n_blocks, 256), 256, 0, st>>>(s.bv, m->n_words, m->n_blocks, block_cnt);
        block_scan_kernel<<<1, 1024, 0, st>>>(block_cnt, m->n_blocks);
        ctx->launches += 2;
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->words.data(), s.bv, m->n_words * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->block_ranks.data(), block_cnt, m->n_blocks * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            rc = ctx->fail(AIX_ERR_CUDA, "mphf assign: %s", cudaGetErrorString(e));
            break;
        }
        rc = AIX_OK;
        break;
    }
    cleanup();
    if (rc == AIX_ERR_BUILD) ctx->fail(AIX_ERR_BUILD, "hypergraph not peelable after 64 seeds (duplicate keys?)");
    if (rc == AIX_OK) rc = mphf_build_layout(ctx, m);
    if (rc != AIX_OK) {
        aix_mphf_destroy(ctx, m);
        return rc;
    }
    *out = m;
    return AIX_OK;
}

extern "C" {

int aix_mphf_build_dev(aix_ctx *ctx, const uint64_t *kmers_dev, uint64_t n, int k, aix_mphf **out) {
    if (!ctx || !out || (n && !kmers_dev)) return AIX_ERR_ARG;
    return build_impl(ctx, kmers_dev, n, k, out);
}

int aix_mphf_build(aix_ctx *ctx, const uint64_t *kmers, uint64_t n, int k, aix_mphf **out) {
    if (!ctx || !out || (n && !kmers)) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *d = nullptr;
    cudaError_t e = cudaMalloc(&d, (n ? n : 1) * 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_NOMEM, "mphf build keys: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(d, kmers, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    int rc = build_impl(ctx, d, n, k, out);
    cudaFree(d);
    return rc;
}

int aix_index23_fill_dev(aix_ctx *ctx, const aix_mphf *m, const uint64_t *kmers_dev, const uint32_t *counts_dev,
                         uint64_t n, uint64_t *checker_dev, uint32_t *tf_dev) {
    if (!ctx || !m) return AIX_ERR_ARG;
    if (n != m->n) return ctx->fail(AIX_ERR_ARG, "index fill: %llu k-mers for an MPHF of %llu keys",
                                    (unsigned long long)n, (unsigned long long)m->n);
    if (n == 0) return AIX_OK;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *bad = nullptr, hbad = 0;
    AIX_CUDA(ctx, cudaMalloc(&bad, 8));
    cudaMemsetAsync(bad, 0, 8, ctx->stream);
    cudaMemsetAsync(checker_dev, 0, n * 8, ctx->stream);  // hash.cpp:831-839
    cudaMemsetAsync(tf_dev, 0, n * 4, ctx->stream);
    index_fill_kernel<<<aix_grid(n, 256), 256, 0, ctx->stream>>>(m->dev(), kmers_dev, counts_dev, n, checker_dev, tf_dev, bad);
    ctx->launches++;
    cudaMemcpyAsync(&hbad, bad, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(bad);
    if (e != cudaSuccess) return ctx->fail(AIX_ERR_CUDA, "index fill: %s", cudaGetErrorString(e));
    if (hbad) return ctx->fail(AIX_ERR_ARG, "index fill: %llu k-mers are not keys of this MPHF", hbad);
    return AIX_OK;
}

int aix_index23_fill(aix_ctx *ctx, const aix_mphf *m, const uint64_t *kmers, const uint32_t *counts, uint64_t n,
                     uint64_t *checker_out, uint32_t *tf_out) {
    if (!ctx || !m || (n && (!kmers || !counts || !checker_out || !tf_out))) return AIX_ERR_ARG;
    if (n == 0) return AIX_OK;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *k_dev, *c_dev, *chk_dev, *tf_dev;
    AIX_TRY(ctx->reserve(SCR_IN0, n * 8, &k_dev));
    AIX_TRY(ctx->reserve(SCR_IN1, n * 4, &c_dev));
    AIX_TRY(ctx->reserve(SCR_OUT0, n * 8, &chk_dev));
    AIX_TRY(ctx->reserve(SCR_OUT1, n * 4, &tf_dev));
    AIX_CUDA(ctx, cudaMemcpyAsync(k_dev, kmers, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    AIX_CUDA(ctx, cudaMemcpyAsync(c_dev, counts, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    AIX_TRY(aix_index23_fill_dev(ctx, m, (const uint64_t *)k_dev, (const uint32_t *)c_dev, n, (uint64_t *)chk_dev, (uint32_t *)tf_dev));
    AIX_CUDA(ctx, cudaMemcpyAsync(checker_out, chk_dev, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaMemcpyAsync(tf_out, tf_dev, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

// sort + run-length of `n_keys` k-mers held in keys (keys_alt = spare of the same size, cnts = u32[n_keys]);
// the unique k-mers end up in *uniq_out (one of the two key buffers) with their counts in cnts.
// Hand-written LSD radix sort + run-length kernels (radix_sort.cu): the `sort | uniq -c` of the reference's
// counting stage (scripts/compute_aindex.py:140-182).
static int c23_sort_rle(aix_ctx *ctx, uint64_t *keys, uint64_t *keys_alt, uint32_t *cnts, uint64_t n_keys, int end_bit,
                        uint64_t **uniq_out, uint64_t *n_uniq) {
    *n_uniq = 0;
    *uniq_out = keys_alt;
    if (n_keys == 0) return AIX_OK;
    uint64_t *sorted = keys;
    AIX_TRY(aix::radix_sort_u64(ctx, ctx->stream, keys, keys_alt, n_keys, 0, end_bit, &sorted));
    uint64_t *spare = sorted == keys ? keys_alt : keys;  // reused for the unique keys
    AIX_TRY(aix::rle_u64(ctx, ctx->stream, sorted, n_keys, spare, cnts, n_uniq));
    *uniq_out = spare;
    return AIX_OK;
}

// per-pass key capacity: keys + spare + counts + sort scratch must fit the free HBM
static uint64_t c23_pass_capacity(uint64_t reserve_bytes) {
    uint64_t cap = 1ull << 36;
    if (const char *e = getenv("AIX_CANONICAL23_PASS_KEYS")) {  // test hook: force the multi-pass path
        uint64_t v = strtoull(e, nullptr, 10);
        if (v >= 1024) return v < cap ? v : cap;
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        uint64_t usable = free_b > reserve_bytes ? free_b - reserve_bytes : 0;
        uint64_t by_mem = usable / 30;  // 8 + 8 + 4 bytes per key + sort status words + run-length scratch + slack
        if (by_mem < cap) cap = by_mem;
    }
    return cap;
}

static int c23_single_pass(aix_ctx *ctx, const uint8_t *reads_dev, uint64_t len, uint64_t n_win, uint64_t *n_out) {
    AixTrace trace(ctx->stream, "canonical23 count");
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    uint32_t *cnts = nullptr;
    auto cleanup = [&]() { cudaFree(keys); cudaFree(keys_alt); cudaFree(cnts); };
    cudaError_t e = cudaMalloc(&keys, n_win * 8);
    if (e == cudaSuccess) e = cudaMalloc(&keys_alt, n_win * 8);
    if (e == cudaSuccess) e = cudaMalloc(&cnts, n_win * 4);
    if (e != cudaSuccess) {
        cudaGetLastError(); cleanup();
        return ctx->fail(AIX_ERR_NOMEM, "canonical23 buffers: %s", cudaGetErrorString(e));
    }
    trace.mark("allocate keys + spare + counts");
    canonical23_kernel<<<aix_grid((n_win + kCanRoll - 1) / kCanRoll, 128), 128, 0, ctx->stream>>>(reads_dev, len, keys);
    ctx->launches++;
    trace.mark("emit canonical k-mers");
    uint64_t *uniq = nullptr, n = 0;
    int rc = c23_sort_rle(ctx, keys, keys_alt, cnts, n_win, 47, &uniq, &n)  /* 46 k-mer bits + the bit that sets the all-ones sentinel apart */;
    if (rc != AIX_OK) { cleanup(); return rc; }
    uint64_t last_key = 0;
    if (n) e = cudaMemcpy(&last_key, uniq + (n - 1), 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && n && last_key == ~0ULL) --n;  // the invalid-window sentinel sorts last
    if (e == cudaSuccess && n) {
        e = cudaMalloc(&ctx->c23_kmers_dev, n * 8);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->c23_counts_dev, n * 4);
        if (e == cudaSuccess) e = cudaMemcpy(ctx->c23_kmers_dev, uniq, n * 8, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(ctx->c23_counts_dev, cnts, n * 4, cudaMemcpyDeviceToDevice);
    }
    trace.mark("sort + run-length + copy out");
    cleanup();
    trace.mark("free buffers");
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_NOMEM, "canonical23 result: %s", cudaGetErrorString(e));
    }
    ctx->c23_n = n;
    *n_out = n;
    return AIX_OK;
}

static int c23_multi_pass(aix_ctx *ctx, const uint8_t *reads_dev, uint64_t len, uint64_t n_win, uint64_t cap, uint64_t *n_out) {
    cudaStream_t st = ctx->stream;
    const unsigned grid = aix_grid((n_win + kCanRoll - 1) / kCanRoll, 128);
    unsigned long long *bins_dev = nullptr, *cursor = nullptr;
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    uint32_t *cnts = nullptr;
    struct Seg { uint64_t *k; uint32_t *c; uint64_t n; };
    std::vector<Seg> segs;
    auto cleanup = [&]() {
        cudaFree(bins_dev); cudaFree(cursor); cudaFree(keys); cudaFree(keys_alt); cudaFree(cnts);
        for (auto &g : segs) { cudaFree(g.k); cudaFree(g.c); }
    };
#define C23_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            cudaGetLastError(); cleanup();                                                          \
            return ctx->fail(AIX_ERR_CUDA, "canonical23 %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
        }                                                                                           \
    } while (0)
    C23_CUDA(cudaMalloc(&bins_dev, kCanBins * 8));
    C23_CUDA(cudaMalloc(&cursor, 8));
    C23_CUDA(cudaMemsetAsync(bins_dev, 0, kCanBins * 8, st));
    canonical23_bins_kernel<<<grid, 128, 0, st>>>(reads_dev, len, bins_dev);
    ctx->launches++;
    std::vector<unsigned long long> bins(kCanBins);
    C23_CUDA(cudaMemcpyAsync(bins.data(), bins_dev, kCanBins * 8, cudaMemcpyDeviceToHost, st));
    C23_CUDA(cudaStreamSynchronize(st));
    uint64_t max_bin = 0;
    for (auto b : bins) max_bin = b > max_bin ? b : max_bin;
    if (max_bin > cap) {
        cleanup();
        return ctx->fail(AIX_ERR_NOMEM, "canonical23: one k-mer range holds %llu windows, more than a pass can sort (%llu)",
                         (unsigned long long)max_bin, (unsigned long long)cap);
    }
    // ranges of whole bins with at most `cap` k-mers each
    std::vector<std::pair<int, int>> ranges;
    uint64_t largest = 0;
    for (int b = 0; b < kCanBins;) {
        uint64_t acc = 0;
        int e = b;
        while (e < kCanBins && acc + bins[e] <= cap) acc += bins[e++];
        if (acc) ranges.push_back({b, e});
        largest = acc > largest ? acc : largest;
        b = e;
    }
    C23_CUDA(cudaMalloc(&keys, (largest ? largest : 1) * 8));
    C23_CUDA(cudaMalloc(&keys_alt, (largest ? largest : 1) * 8));
    C23_CUDA(cudaMalloc(&cnts, (largest ? largest : 1) * 4));
    uint64_t n_total = 0;
    for (auto &rg : ranges) {
        const uint64_t lo = (uint64_t)rg.first << (46 - kCanBinsLog2), hi = (uint64_t)rg.second << (46 - kCanBinsLog2);
        C23_CUDA(cudaMemsetAsync(cursor, 0, 8, st));
        canonical23_range_kernel<<<grid, 128, 0, st>>>(reads_dev, len, lo, hi, keys, cursor);
        ctx->launches++;
        unsigned long long n_keys = 0;
        C23_CUDA(cudaMemcpyAsync(&n_keys, cursor, 8, cudaMemcpyDeviceToHost, st));
        C23_CUDA(cudaStreamSynchronize(st));
        uint64_t *uniq = nullptr, n = 0;
        int rc = c23_sort_rle(ctx, keys, keys_alt, cnts, n_keys, 46, &uniq, &n);
        if (rc != AIX_OK) { cleanup(); return rc; }
        if (n) {
            Seg g = {nullptr, nullptr, n};
            segs.push_back(g);
            C23_CUDA(cudaMalloc(&segs.back().k, n * 8));
            C23_CUDA(cudaMalloc(&segs.back().c, n * 4));
            C23_CUDA(cudaMemcpyAsync(segs.back().k, uniq, n * 8, cudaMemcpyDeviceToDevice, st));
            C23_CUDA(cudaMemcpyAsync(segs.back().c, cnts, n * 4, cudaMemcpyDeviceToDevice, st));
            C23_CUDA(cudaStreamSynchronize(st));
            n_total += n;
        }
    }
    cudaFree(keys); keys = nullptr;
    cudaFree(keys_alt); keys_alt = nullptr;
    cudaFree(cnts); cnts = nullptr;
    if (n_total) {
        C23_CUDA(cudaMalloc(&ctx->c23_kmers_dev, n_total * 8));
        C23_CUDA(cudaMalloc(&ctx->c23_counts_dev, n_total * 4));
        uint64_t at = 0;
        for (auto &g : segs) {  // ranges are ascending, so the concatenation is sorted
            C23_CUDA(cudaMemcpyAsync(ctx->c23_kmers_dev + at, g.k, g.n * 8, cudaMemcpyDeviceToDevice, st));
            C23_CUDA(cudaMemcpyAsync(ctx->c23_counts_dev + at, g.c, g.n * 4, cudaMemcpyDeviceToDevice, st));
            at += g.n;
        }
        C23_CUDA(cudaStreamSynchronize(st));
    }
#undef C23_CUDA
    cleanup();
    ctx->c23_n = n_total;
    *n_out = n_total;
    return AIX_OK;
}

// distinct canonical 23-mers + counts of a reads image already in HBM; the result stays in
// ctx (c23_*), sorted ascending
int aix_canonical23_count_dev(aix_ctx *ctx, const uint8_t *reads_dev, uint64_t len, uint64_t *n_out) {
    if (!ctx || !n_out) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->l2_unpin();
    if (ctx->c23_kmers_dev) { cudaFree(ctx->c23_kmers_dev); ctx->c23_kmers_dev = nullptr; }
    if (ctx->c23_counts_dev) { cudaFree(ctx->c23_counts_dev); ctx->c23_counts_dev = nullptr; }
    ctx->c23_n = 0;
    *n_out = 0;
    if (len < 23) return AIX_OK;
    const uint64_t n_win = len - 22;
    const uint64_t cap = c23_pass_capacity(n_win);  // leave room for the result (12 B per distinct k-mer, twice)
    if (cap < 1024) return ctx->fail(AIX_ERR_NOMEM, "canonical23: not enough free HBM");
    if (n_win <= cap) return c23_single_pass(ctx, reads_dev, len, n_win, n_out);
    return c23_multi_pass(ctx, reads_dev, len, n_win, cap, n_out);
}

int aix_canonical23_result_dev(aix_ctx *ctx, const uint64_t **kmers_dev, const uint32_t **counts_dev, uint64_t *n) {
    if (!ctx) return AIX_ERR_ARG;
    if (kmers_dev) *kmers_dev = ctx->c23_kmers_dev;
    if (counts_dev) *counts_dev = ctx->c23_counts_dev;
    if (n) *n = ctx->c23_n;
    return AIX_OK;
}

int aix_canonical23_count(aix_ctx *ctx, const uint8_t *reads, uint64_t len, uint64_t *n_out, uint64_t *kmers_out,
                          uint32_t *counts_out) {
    if (!ctx || !n_out) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!kmers_out) {
        if (len && !reads) return ctx->fail(AIX_ERR_ARG, "null reads");
        uint8_t *d = nullptr;
        cudaError_t e = cudaMalloc(&d, len ? len : 1);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(AIX_ERR_NOMEM, "canonical23 reads: %s", cudaGetErrorString(e));
        }
        cudaMemcpyAsync(d, reads, len, cudaMemcpyHostToDevice, ctx->stream);
        int rc = aix_canonical23_count_dev(ctx, d, len, n_out);
        cudaFree(d);
        return rc;
    }
    *n_out = ctx->c23_n;
    if (ctx->c23_n) {
        AIX_CUDA(ctx, cudaMemcpy(kmers_out, ctx->c23_kmers_dev, ctx->c23_n * 8, cudaMemcpyDeviceToHost));
        if (counts_out) AIX_CUDA(ctx, cudaMemcpy(counts_out, ctx->c23_counts_dev, ctx->c23_n * 4, cudaMemcpyDeviceToHost));
    }
    return AIX_OK;
}

// the text files the reference's index build consumes (scripts/compute_aindex.py:140-200): `.dat` = "KMER\tCOUNT\n"
// (input of compute_index, hash.cpp:696-701) and the key file = "KMER\n" (`cut -f1`, input of compute_mphf_seq).
// The k-mers are decoded on the GPU (get_bitset_dna23, kmers.cpp:89-114); the host only formats the lines.
int aix_write_dat(aix_ctx *ctx, const uint64_t *kmers, const uint32_t *counts, uint64_t n, const char *dat_path,
                  const char *keys_path) {
    if (!ctx || (n && !kmers) || (!dat_path && !keys_path)) return AIX_ERR_ARG;
    if (dat_path && n && !counts) return ctx->fail(AIX_ERR_ARG, "write_dat: counts are required for the .dat file");
    FILE *fd = dat_path ? fopen(dat_path, "wb") : nullptr, *fk = keys_path ? fopen(keys_path, "wb") : nullptr;
    if ((dat_path && !fd) || (keys_path && !fk)) {
        if (fd) fclose(fd);
        if (fk) fclose(fk);
        return ctx->fail(AIX_ERR_IO, "write_dat: cannot create %s", (dat_path && !fd) ? dat_path : keys_path);
    }
    const uint64_t chunk = 1u << 22;
    std::vector<uint8_t> ascii(chunk * 23);
    std::vector<char> lines(chunk * 36);
    bool ok = true;
    int rc = AIX_OK;
    for (uint64_t i0 = 0; i0 < n && ok && rc == AIX_OK; i0 += chunk) {
        const uint64_t m = n - i0 < chunk ? n - i0 : chunk;
        rc = aix_decode_kmers(ctx, kmers + i0, m, 23, ascii.data());
        if (rc != AIX_OK) break;
        if (fk) {
            char *w = lines.data();
            for (uint64_t i = 0; i < m; ++i) {
                memcpy(w, ascii.data() + i * 23, 23);
                w[23] = '\n';
                w += 24;
            }
            ok = fwrite(lines.data(), 1, (size_t)(w - lines.data()), fk) == (size_t)(w - lines.data());
        }
        if (fd && ok) {
            char *w = lines.data();
            for (uint64_t i = 0; i < m; ++i) {
                memcpy(w, ascii.data() + i * 23, 23);
                w += 23;
                *w++ = '\t';
                char tmp[12];
                int d = 0;
                uint32_t c = counts[i0 + i];
                do { tmp[d++] = (char)('0' + c % 10); c /= 10; } while (c);
                while (d) *w++ = tmp[--d];
                *w++ = '\n';
            }
            ok = fwrite(lines.data(), 1, (size_t)(w - lines.data()), fd) == (size_t)(w - lines.data());
        }
    }
    if (fd) ok = (fclose(fd) == 0) && ok;
    if (fk) ok = (fclose(fk) == 0) && ok;
    if (rc != AIX_OK) return rc;
    return ok ? AIX_OK : ctx->fail(AIX_ERR_IO, "write_dat: short write");
}

}  // extern "C"
