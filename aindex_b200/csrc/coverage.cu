// coverage.cu -- K5: per-position term-frequency profile of sequences.
//
// Reference: AIndex.get_sequence_coverage (aindex/core/aindex.py:314-322): for every window
// start i of a sequence, tf = get_tf_value(seq[i:i+k]) (python_wrapper.cpp:644-650 ->
// get_tf_value_23mer :610-627 or get_tf_value_13mer :482-503); coverage[i] = tf if
// tf >= cutoff else 0.  No newline / '~' / 'N' skipping: such windows are simply looked up.
//
// One thread per output position; a CTA owns 256 consecutive outputs, locates the sequences
// they belong to by binary search over the output offsets (one coarse search per CTA, one
// short search per thread) and reads its 23 bytes as six/seven aligned words from L1/L2
// (neighbouring threads share all but one byte).
#include "aix_internal.cuh"
#include "query23.cuh"

namespace aix {

constexpr int kCovBlock = 256;

// out_offs[s] = sum_{t<s} max(0, len_t - k + 1), single CTA, n_seq+1 entries
__global__ void coverage_offsets_kernel(const int64_t *__restrict__ offs, uint64_t n_seq, int k,
                                        unsigned long long *__restrict__ out_offs) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (uint64_t s0 = 0; s0 < n_seq; s0 += blockDim.x) {
        uint64_t s = s0 + threadIdx.x;
        unsigned long long v = 0;
        if (s < n_seq) {
            int64_t len = offs[s + 1] - offs[s];
            v = len >= k ? (unsigned long long)(len - k + 1) : 0ull;
        }
        unsigned long long x = v;
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
        if (s < n_seq) out_offs[s] = base + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) out_offs[n_seq] = carry;
}

// largest s in [lo, hi] with out_offs[s] <= o  (out_offs[lo] <= o guaranteed)
__device__ __forceinline__ uint64_t find_seq(const unsigned long long *__restrict__ out_offs, uint64_t lo, uint64_t hi,
                                             unsigned long long o) {
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo + 1) / 2;
        if (__ldg(out_offs + mid) <= o) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// cta_seq[b] = sequence that holds output b * 256 (one thread per CTA of the main kernel, b = 0 .. n_ctas;
// the last entry is n_seq - 1).  Doing these ~20-step searches here, all in parallel, instead of by two threads
// of every CTA in front of a barrier removed the largest stall of the first version (profiles/r01_c4_ncu_before.txt:
// barrier 9.5 of 24 stall cycles per issue).
__global__ void coverage_cta_seq_kernel(const unsigned long long *__restrict__ out_offs, uint64_t n_seq, uint64_t total_out,
                                        uint64_t n_ctas, uint32_t *__restrict__ cta_seq) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_ctas) return;
    unsigned long long o = b * kCovBlock;
    cta_seq[b] = (b == n_ctas || o >= total_out) ? (uint32_t)(n_seq - 1) : (uint32_t)find_seq(out_offs, 0, n_seq - 1, o);
}

template <int K, bool kCanon>
__global__ void __launch_bounds__(kCovBlock) coverage_kernel(Index23Dev ix, MphfDev m, const uint64_t *__restrict__ tf13_direct,
                                                           const uint8_t *__restrict__ seqs, const int64_t *__restrict__ offs,
                                                           const unsigned long long *__restrict__ out_offs,
                                                           const uint32_t *__restrict__ cta_seq, uint64_t total_out, uint32_t cutoff,
                                                           uint32_t *__restrict__ out) {
    const uint64_t o = (uint64_t)blockIdx.x * kCovBlock + threadIdx.x;
    if (o >= total_out) return;
    const uint64_t s_lo = __ldg(cta_seq + blockIdx.x), s_hi = __ldg(cta_seq + blockIdx.x + 1);
    const uint64_t s = s_lo == s_hi ? s_lo : find_seq(out_offs, s_lo, s_hi, o);
    const uint8_t *p = seqs + __ldg(offs + s) + (o - __ldg(out_offs + s));
    uint32_t tf;
    if (K == 23) {
        uint64_t r0, r1, r2;
        load_window23(p, r0, r1, r2);
        tf = find23_window<kCanon>(ix, m, r0, r1, r2).tf;
    } else {
        // get_tf_value_13mer: upper-case ACGT only, u64 count narrowed to u32.  13 bytes at any alignment are
        // inside four aligned words (offset <= 3, so the last byte is at most byte 15); SIMD encode + validity as in the batch query kernel
        const uintptr_t a = (uintptr_t)p;
        const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8u;
        const uint32_t x0 = __ldg(w), x1 = __ldg(w + 1), x2 = __ldg(w + 2), x3 = __ldg(w + 3);
        const uint32_t y[4] = {__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh),
                               (x3 >> sh) & 0xFFu};
        uint32_t pk[4], bad = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c4 = ((y[j] >> 1) ^ (y[j] >> 2)) & 0x03030303u;
            pk[j] = j == 3 ? (c4 & 3u) : ((c4 * 0x40100401u) >> 24);
            const uint32_t diff = expect_acgt4(c4) ^ y[j];
            bad |= j == 3 ? (diff & 0xFFu) : diff;
        }
        const uint32_t v = (pk[0] << 18) | (pk[1] << 10) | (pk[2] << 2) | pk[3];
        tf = bad == 0 ? (uint32_t)__ldg(tf13_direct + v) : 0u;
    }
    __stcs(out + o, tf >= cutoff ? tf : 0u);
}

}  // namespace aix

using namespace aix;

extern "C" {

// the device form on a given stream; `slot` = scratch slot of the per-call offset tables (one per stream in flight)
static int coverage_on(aix_ctx *ctx, cudaStream_t st, int slot, const aix_index23 *ix23, const aix_index13 *ix13,
                       const uint8_t *seqs_dev, const int64_t *offs_dev, uint64_t n_seq, uint64_t total_out, int k,
                       uint32_t cutoff, uint32_t *out_dev) {
    if (!ctx) return AIX_ERR_ARG;
    if (k == 23 && !ix23) return ctx->fail(AIX_ERR_STATE, "23-mer index not loaded");
    if (k == 13 && !ix13) return ctx->fail(AIX_ERR_STATE, "13-mer index not loaded");
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (n_seq == 0 || total_out == 0) return AIX_OK;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->l2_unpin();
    if (n_seq >= (1ull << 32)) return ctx->fail(AIX_ERR_ARG, "coverage: at most 2^32-1 sequences per call");
    const uint64_t n_ctas = (total_out + kCovBlock - 1) / kCovBlock;
    void *oo;
    const size_t oo_bytes = ((n_seq + 1) * 8 + 255) & ~(size_t)255;
    AIX_TRY(ctx->reserve(slot, oo_bytes + (n_ctas + 1) * 4, &oo));
    uint32_t *cta_seq = reinterpret_cast<uint32_t *>((char *)oo + oo_bytes);
    coverage_offsets_kernel<<<1, 1024, 0, st>>>(offs_dev, n_seq, k, (unsigned long long *)oo);
    AIX_LAUNCH_CHECK(ctx);
    coverage_cta_seq_kernel<<<aix_grid(n_ctas + 1, 256), 256, 0, st>>>((unsigned long long *)oo, n_seq, total_out, n_ctas, cta_seq);
    AIX_LAUNCH_CHECK(ctx);
    unsigned grid = (unsigned)n_ctas;
    Index23Dev id = {};
    MphfDev md = {};
    if (k == 23) {
        id = ix23->dev();
        md = ix23->mphf_dev();
        // coverage is hit-dominated (sequences of the indexed organism): the fingerprint tier would add a
        // dependent L2 round trip in front of nearly every HBM record load.  AIX_COVERAGE_TIER=1 keeps it.
        const char *e = getenv("AIX_COVERAGE_TIER");
        if (!(e && atoi(e) != 0)) id.fp = nullptr;
        if (ix23->canonical_only)
            coverage_kernel<23, true><<<grid, kCovBlock, 0, st>>>(id, md, nullptr, seqs_dev, offs_dev, (unsigned long long *)oo, cta_seq, total_out, cutoff, out_dev);
        else
            coverage_kernel<23, false><<<grid, kCovBlock, 0, st>>>(id, md, nullptr, seqs_dev, offs_dev, (unsigned long long *)oo, cta_seq, total_out, cutoff, out_dev);
    } else {
        coverage_kernel<13, true><<<grid, kCovBlock, 0, st>>>(id, md, ix13->tf_direct_dev, seqs_dev, offs_dev, (unsigned long long *)oo, cta_seq, total_out, cutoff, out_dev);
    }
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_coverage_dev(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13, const uint8_t *seqs_dev,
                     const int64_t *offs_dev, uint64_t n_seq, uint64_t total_bytes, uint64_t total_out, int k,
                     uint32_t cutoff, uint32_t *out_dev) {
    (void)total_bytes;
    if (!ctx) return AIX_ERR_ARG;
    return coverage_on(ctx, ctx->stream, SCR_TMP1, ix23, ix13, seqs_dev, offs_dev, n_seq, total_out, k, cutoff, out_dev);
}

// Host buffers: sequences are processed in groups of whole sequences (<= ~256 MiB of
// output per group) on two alternating streams, offsets rebased per group.
int aix_coverage(aix_ctx *ctx, const aix_index23 *ix23, const aix_index13 *ix13, const uint8_t *seqs, const int64_t *offs,
                 uint64_t n_seq, int k, uint32_t cutoff, uint32_t *out) {
    if (!ctx) return AIX_ERR_ARG;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (n_seq == 0) return AIX_OK;
    if (!seqs || !offs || !out) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint64_t s = 0; s < n_seq; ++s)
        if (offs[s + 1] < offs[s]) return ctx->fail(AIX_ERR_ARG, "offsets must be non-decreasing");
    const uint64_t group_out_target = 64ull << 20;  // output values per group
    uint64_t s0 = 0, out_done = 0;
    // two groups in flight on the two transfer streams, each with its own device buffers: the H2D copy of group
    // g+1 and the D2H copy of group g-1 overlap the kernel of group g.  Device buffers are reused in stream order;
    // the host waits only at the end.
    std::vector<int64_t> rel[2];
    int g = 0;
    while (s0 < n_seq) {
        uint64_t s1 = s0, g_out = 0;
        while (s1 < n_seq) {
            int64_t len = offs[s1 + 1] - offs[s1];
            uint64_t w = len >= k ? (uint64_t)(len - k + 1) : 0;
            if (g_out && g_out + w > group_out_target) break;
            g_out += w;
            ++s1;
        }
        const uint64_t g_seq = s1 - s0;
        const uint64_t byte0 = (uint64_t)offs[s0], g_bytes = (uint64_t)offs[s1] - byte0;
        if (g_out) {
            const int b = g & 1;
            cudaStream_t st = ctx->xfer[b];
            if (g >= 2) AIX_CUDA(ctx, cudaStreamSynchronize(st));  // rel[b] and the buffers of group g-2 are free again
            rel[b].resize(g_seq + 1);
            for (uint64_t i = 0; i <= g_seq; ++i) rel[b][i] = offs[s0 + i] - (int64_t)byte0;
            void *d_seq, *d_offs, *d_out;
            AIX_TRY(ctx->reserve(SCR_IN0 + b, g_bytes + 64, &d_seq));
            AIX_TRY(ctx->reserve(SCR_LEN0 + b, (g_seq + 1) * 8, &d_offs));
            AIX_TRY(ctx->reserve(SCR_OUT0 + b, g_out * 4, &d_out));
            AIX_CUDA(ctx, cudaMemcpyAsync(d_seq, seqs + byte0, g_bytes, cudaMemcpyHostToDevice, st));
            AIX_CUDA(ctx, cudaMemcpyAsync(d_offs, rel[b].data(), (g_seq + 1) * 8, cudaMemcpyHostToDevice, st));
            AIX_TRY(coverage_on(ctx, st, SCR_TMP0 + b, ix23, ix13, (const uint8_t *)d_seq, (const int64_t *)d_offs, g_seq, g_out, k,
                                cutoff, (uint32_t *)d_out));
            AIX_CUDA(ctx, cudaMemcpyAsync(out + out_done, d_out, g_out * 4, cudaMemcpyDeviceToHost, st));
            ++g;
        }
        out_done += g_out;
        s0 = s1;
    }
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[0]));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[1]));
    return AIX_OK;
}

}  // extern "C"
