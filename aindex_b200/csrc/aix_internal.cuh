// aix_internal.cuh -- host-side object definitions behind the opaque C-ABI handles.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/aindex_cuda.h"
#include "device_common.cuh"

struct aix_mphf {
    uint64_t n = 0, hash_domain = 0, seed = 0, bv_size = 0, n_words = 0, n_blocks = 0;
    std::vector<uint64_t> words;        // .pf order (host copy, for save / inspection)
    std::vector<uint64_t> block_ranks;  // .pf order
    ulonglong2 *recs_dev = nullptr;     // B200 layout (wide records), see device_common.cuh
    uint4 *crecs_dev = nullptr;         // B200 layout (compact records); exactly one of the two is set
    uint64_t layout_bytes = 0;          // size of the device structure (L2 budget of the fingerprint tier)
    aix::MphfDev dev() const {
        aix::MphfDev d;
        d.n = n; d.hash_domain = hash_domain; d.seed = seed;
        // floor(2^64 / d); for d == 1 the quotient does not fit: 2^64-1 still gives q in {q_true-1, q_true}
        d.magic = hash_domain > 1 ? (uint64_t)((((unsigned __int128)1) << 64) / hash_domain) : ~0ULL;
        d.recs = recs_dev;
        d.crecs = crecs_dev;
        d.frecs = nullptr;
        return d;
    }
};

struct aix_index23 {
    uint64_t n = 0;
    int canonical_only = 0;
    const aix_mphf *mphf = nullptr;
    uint4 *recs_dev = nullptr;  // {checker lo, checker hi, tf, 0}
    uint8_t *fp_dev = nullptr;  // fingerprint tier (may be null)
    int fp_bits = 0;            // 8 or 4 when fp_dev is set
    uint4 *frecs_dev = nullptr; // fused MPHF + 4-bit fingerprint records (device_common.cuh); replaces the tier when set
    uint64_t frecs_bytes = 0;
    // front filter (device_common.cuh: Index23Dev::bloom) and what the batch launcher has learnt about the traffic:
    // the filter kernel counts the queries it saw and those that passed the filter (sampled CTAs); the counts come back
    // with an asynchronous 16-byte copy after each of its launches and steer the next launches (tf_query.cu)
    uint2 *bloom_dev = nullptr;
    uint32_t bloom_words = 0;
    unsigned long long *qstats_dev = nullptr;            // {queries, passed}
    volatile unsigned long long *qstats_host = nullptr;  // pinned copy
    mutable unsigned long long seen_q = 0, seen_p = 0, launches_filter = 0, launches_direct = 0;
    mutable double pass_rate = -1.0;                     // < 0: nothing observed yet
    mutable int direct_since_probe = 0;
    int filter_mode = 0;                                 // 0 = decide from the pass rate, 1 = always, 2 = never (aix_index23_set_filter)
    // the MPHF as this index's kernels see it: the fused records when they exist
    aix::MphfDev mphf_dev() const {
        aix::MphfDev d = mphf->dev();
        d.frecs = frecs_dev;
        return d;
    }
    aix::Index23Dev dev() const {
        aix::Index23Dev d;
        d.n = n; d.canonical_only = canonical_only; d.recs = recs_dev; d.fp = fp_dev; d.fp_bits = fp_bits;
        d.bloom = bloom_dev; d.bloom_words = bloom_words;
        return d;
    }
};

struct aix_index13 {
    const aix_mphf *mphf = nullptr;
    uint64_t *tf_mphf_dev = nullptr;    // u64[4^13], .tf.bin order (id = mphf(kmer))
    uint64_t *tf_direct_dev = nullptr;  // u64[4^13], direct-address order (v = 2-bit value)
};

struct aix_positions {
    uint64_t n_indices = 0, n_positions = 0;
    uint64_t *indices_dev = nullptr;
    uint64_t *positions_dev = nullptr;
    bool pooled = false;  // arrays come from the ctx's memory pool (built on the device) rather than cudaMalloc (uploaded)
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// Front filter of a 23-mer index held in persisting L2 lines (tf_query.cu: l2_window_on): the set-aside configured on a
// device and whether persisting lines may still be resident.  A property of the DEVICE, shared by every ctx on it.
struct AixL2State {
    std::atomic<size_t> set_aside{0};
    std::atomic<bool> pinned{false};
};
inline AixL2State g_aix_l2[64];

struct aix_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // *_dev entry points + pipelines' compute
    cudaStream_t xfer[2] = {nullptr, nullptr};
    cudaEvent_t ev[8] = {};
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    DevBuf scratch[8];                  // grow-only device scratch, indexed by role
    // stream-ordered pool for the large, short-lived buffers of the builders (sort keys, status words, ...): freed
    // blocks stay in the pool (release threshold = max), so a second build does not pay the driver's map / unmap of
    // 100 GB again (measured: 50-700 ms per build with cudaMalloc / cudaFree, more while NVML is being polled).
    // aix_ctx_trim() hands the cached memory back.
    cudaMemPool_t pool = nullptr;
    std::unordered_map<void *, size_t> plain_allocs;        // cudaMalloc'ed buffers in use (multi-GPU exchange buffers) and their sizes
    std::vector<std::pair<void *, size_t>> plain_cache;     // ... and the freed ones, kept for the next build (aix_plain_alloc)
    void *small_host = nullptr;         // pinned + device-mapped staging of the small-batch path (batch_pipeline.cuh)
    // single-query mailbox (tf_query.cu): a one-thread resident kernel that polls a request slot in mapped host memory,
    // so that get_tf_value() costs two PCIe traversals instead of a kernel launch + a stream synchronisation
    void *mbox_host = nullptr;          // MboxSlot, cudaHostAllocMapped
    void *mbox_dev = nullptr;           // its device address
    cudaStream_t mbox_stream = nullptr;
    const void *mbox_owner = nullptr;   // the index the running kernel serves
    bool mbox_launched = false, mbox_broken = false;
    uint32_t mbox_seq = 0;
    // count13 streaming state
    uint32_t *c13_hist32 = nullptr;     // u32[4^13]
    uint64_t *c13_hist64 = nullptr;     // u64[4^13]
    uint64_t *c13_stats_dev = nullptr;  // statistics slots, layout in count13.cu (kStatBase)
    uint64_t c13_pending_windows = 0;   // upper bound of increments not yet flushed
    bool c13_active = false;
    void *c13_peer[16][3] = {};         // IPC-mapped {hist32, hist64, stats} of every rank (own entries = own buffers)
    int c13_n_peers = 0, c13_my_rank = 0;
    bool c13_peer_ipc = true;           // false: the table holds plain device pointers of other ctxs of this process (multi.cu)
    aix_count_stats c13_range_invalid = {0, 0, 0, 0};
    // canonical23 result kept between the two passes
    uint64_t *c23_kmers_dev = nullptr;
    uint32_t *c23_counts_dev = nullptr;
    uint64_t c23_n = 0;

    // persisting L2 lines of a front filter back to normal lines and the set-aside back to the device: called where another
    // kernel family starts (counting, builds, coverage, 13-mer queries), whose working sets want the whole L2 -- with the
    // 50 MB set-aside left configured, the 64 MiB histogram slices of count13 no longer fit (190 -> 126 G k-mers/s)
    void l2_unpin() const {
        AixL2State &l2 = g_aix_l2[device & 63];
        if (l2.pinned.load() || l2.set_aside.load()) {
            cudaCtxResetPersistingL2Cache();
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
            cudaGetLastError();
            l2.pinned = false;
            l2.set_aside = 0;
        }
    }

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    // grow-only scratch allocation
    int reserve(int slot, size_t bytes, void **out) {
        DevBuf &b = scratch[slot];
        if (b.cap < bytes) {
            if (b.p) cudaFree(b.p);
            b.p = nullptr; b.cap = 0;
            size_t want = bytes + (bytes >> 3) + 256;
            cudaError_t e = cudaMalloc(&b.p, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                e = cudaMalloc(&b.p, bytes);
                want = bytes;
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(AIX_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            }
            b.cap = want;
        }
        *out = b.p;
        return AIX_OK;
    }
};

// Buffers another GPU copies into / out of (multi-GPU exchange) are plain cudaMalloc memory, peer-accessible through
// cudaDeviceEnablePeerAccess.  cudaMalloc / cudaFree of tens of GB block for tens of ms each, so freed buffers are kept in a
// per-ctx cache (as the pool keeps its blocks) and handed out again when the size fits; aix_ctx_trim returns them to the
// driver, and so does any allocation that fails while the cache holds memory.
static inline void aix_plain_cache_flush(aix_ctx *ctx) {
    for (auto &b : ctx->plain_cache) cudaFree(b.first);
    ctx->plain_cache.clear();
}
static inline cudaError_t aix_pool_alloc(aix_ctx *ctx, void **p, size_t bytes, cudaStream_t st) {
    cudaError_t e = ctx->pool ? cudaMallocFromPoolAsync(p, bytes ? bytes : 1, ctx->pool, st) : cudaMalloc(p, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation && !ctx->plain_cache.empty()) {
        cudaGetLastError();
        cudaStreamSynchronize(st);
        aix_plain_cache_flush(ctx);
        e = ctx->pool ? cudaMallocFromPoolAsync(p, bytes ? bytes : 1, ctx->pool, st) : cudaMalloc(p, bytes ? bytes : 1);
    }
    return e;
}
template <typename T>
static inline cudaError_t aix_pool_alloc(aix_ctx *ctx, T **p, size_t bytes, cudaStream_t st) {
    return aix_pool_alloc(ctx, (void **)p, bytes, st);
}
template <typename T>
static inline cudaError_t aix_plain_alloc(aix_ctx *ctx, T **p, size_t bytes) {
    if (!bytes) bytes = 1;
    // best fit among the cached blocks that are large enough and at most 25 % too large
    int best = -1;
    for (int i = 0; i < (int)ctx->plain_cache.size(); ++i) {
        const size_t have = ctx->plain_cache[i].second;
        if (have >= bytes && have <= bytes + bytes / 4 && (best < 0 || have < ctx->plain_cache[best].second)) best = i;
    }
    if (best >= 0) {
        *p = (T *)ctx->plain_cache[best].first;
        ctx->plain_allocs[(void *)*p] = ctx->plain_cache[best].second;
        ctx->plain_cache.erase(ctx->plain_cache.begin() + best);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc((void **)p, bytes);
    if (e == cudaErrorMemoryAllocation && !ctx->plain_cache.empty()) {
        cudaGetLastError();
        aix_plain_cache_flush(ctx);
        e = cudaMalloc((void **)p, bytes);
    }
    if (e == cudaSuccess) ctx->plain_allocs[(void *)*p] = bytes;
    return e;
}
// frees a builder temporary: pool memory goes back to the pool (stream-ordered), a plain buffer into the cache -- after the
// stream has drained, since the next user may be another stream or another GPU
static inline void aix_pool_free(aix_ctx *ctx, void *p, cudaStream_t st) {
    if (!p) return;
    if (ctx) {
        auto it = ctx->plain_allocs.find(p);
        if (it != ctx->plain_allocs.end()) {
            const size_t bytes = it->second;
            ctx->plain_allocs.erase(it);
            cudaStreamSynchronize(st);
            ctx->plain_cache.emplace_back(p, bytes);
            return;
        }
    }
    if (ctx && ctx->pool) cudaFreeAsync(p, st);
    else cudaFree(p);
}

struct aix_multi {  // multi.cu: one ctx (and one host thread at a time) per GPU of one box, one process
    std::vector<aix_ctx *> ctx;
    bool peer_ok = false;
    std::string err;
};

#define AIX_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return (ctx)->fail(AIX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,          \
                               cudaGetErrorString(e__));                                         \
    } while (0)

#define AIX_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != AIX_OK) return rc__; \
    } while (0)

#define AIX_LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                         \
        (ctx)->launches++;                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess)                                                                  \
            return (ctx)->fail(AIX_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,      \
                               cudaGetErrorString(e__));                                         \
    } while (0)

// AIX_TRACE=1: per-phase wall times of the multi-kernel host routines on stderr (synchronises the
// stream at every mark, so only for diagnosis -- profiles/ keeps the outputs that DESIGN.md quotes)
struct AixTrace {
    bool on;
    cudaStream_t st;
    const char *what;
    double t_last;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + ts.tv_nsec * 1e-9;
    }
    AixTrace(cudaStream_t s, const char *w) : on(getenv("AIX_TRACE") != nullptr), st(s), what(w), t_last(0) {
        if (on) {
            cudaStreamSynchronize(st);
            t_last = now();
        }
    }
    void mark(const char *phase) {
        if (!on) return;
        cudaStreamSynchronize(st);
        double t = now();
        fprintf(stderr, "[aix trace] %s: %-70s %9.3f ms\n", what, phase, (t - t_last) * 1e3);
        t_last = t;
    }
};

// device time of the work enqueued between the constructor and done(), printed when AIX_TRACE is set (CUDA events on
// the stream: unlike AixTrace::mark this leaves out allocation and host-side waiting)
struct AixTraceSpan {
    bool on;
    cudaStream_t st;
    cudaEvent_t e0, e1;
    explicit AixTraceSpan(cudaStream_t s) : on(getenv("AIX_TRACE") != nullptr), st(s), e0(nullptr), e1(nullptr) {
        if (on) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, st);
        }
    }
    void done(const char *what) {
        if (!on) return;
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[aix trace] device span: %-70s %9.3f ms\n", what, ms);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        on = false;
    }
};

static inline unsigned aix_grid(uint64_t work_items, unsigned block) {
    uint64_t g = (work_items + block - 1) / block;
    return (unsigned)(g ? g : 1);
}

// scratch slot roles
enum { SCR_IN0 = 0, SCR_IN1 = 1, SCR_OUT0 = 2, SCR_OUT1 = 3, SCR_LEN0 = 4, SCR_LEN1 = 5, SCR_TMP0 = 6, SCR_TMP1 = 7 };

// host pointer kind
static inline bool aix_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// internal cross-file entry points
namespace aix {
void mbox_stop(aix_ctx *ctx);  // stops the single-query mailbox kernel (before its index goes away; tf_query.cu)
int mphf_build_layout(aix_ctx *ctx, aix_mphf *m);  // words/block_ranks (host) -> recs_dev
}
