// multi.cu -- several GPUs of one box behind the C-ABI, from ONE process: one aix_ctx and one host thread per
// GPU, peer access enabled between all pairs, plain device pointers instead of CUDA IPC handles, host-side
// joins instead of collectives.  No torch, no NCCL, no MPI.
//
// Reference: Kmer13Counter::count_kmers_from_file (src/count_kmers13.cpp:277-353) spawns N worker threads over one
// shared atomic table (:305-309); here the N workers are GPUs, each with its own direct-address histogram, and
// the one exchange step of the path is the sum over GPUs by k-mer range: GPU r reads range r of
// every GPU's histogram through NVLink (count13_reduce_peers_kernel, count13.cu) -- or, where peer access
// cannot be enabled, receives it with cudaMemcpyPeer -- then permutes its range into .tf.bin order.
#include <thread>

#include "aix_internal.cuh"

namespace aix {

__global__ void add_u64_kernel(unsigned long long *__restrict__ dst, const unsigned long long *__restrict__ src, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// line-structured cut points: cuts[r] = first byte of shard r (cuts[0] = 0, cuts[n] = len), every shard starts at
// a record start so that the per-shard windows are exactly the windows of the whole file
static void shard_image(const uint8_t *b, uint64_t len, int fmt, int n, std::vector<uint64_t> &cuts) {
    cuts.assign(n + 1, len);
    cuts[0] = 0;
    if (n <= 1 || len == 0) return;
    if (fmt == AIX_FMT_FASTQ) {
        // records are 4 lines (count_kmers13.cpp:240-257): newline counts per block, then the first line index
        // that is a multiple of 4 at or after the target
        const int T = n;
        std::vector<uint64_t> nl(T + 1, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                uint64_t lo = len * t / T, hi = len * (t + 1) / T, c = 0;
                const uint8_t *p = b + lo, *e = b + hi;
                while (p < e && (p = (const uint8_t *)memchr(p, '\n', (size_t)(e - p))) != nullptr) { ++c; ++p; }
                nl[t + 1] = c;
            });
        for (auto &x : th) x.join();
        for (int t = 0; t < T; ++t) nl[t + 1] += nl[t];
        for (int r = 1; r < n; ++r) {
            uint64_t pos = len * r / n;    // block r starts here with nl[r] newlines before it
            uint64_t lines = nl[r];        // completed lines before pos
            // advance to the end of the line that completes a record
            uint64_t cut = len;
            const uint8_t *p = b + pos, *e = b + len;
            if (pos > 0 && b[pos - 1] == '\n' && (lines & 3) == 0) cut = pos;
            else {
                while (p < e && (p = (const uint8_t *)memchr(p, '\n', (size_t)(e - p))) != nullptr) {
                    ++lines; ++p;
                    if ((lines & 3) == 0) { cut = (uint64_t)(p - b); break; }
                }
            }
            cuts[r] = cut < cuts[r - 1] ? cuts[r - 1] : cut;
        }
        return;
    }
    for (int r = 1; r < n; ++r) {
        uint64_t pos = len * r / n, cut = len;
        if (pos < cuts[r - 1]) pos = cuts[r - 1];
        if (fmt == AIX_FMT_FASTA) {
            // a shard starts at a header line ('>' at a line start): records are concatenated before counting (:211-235)
            const uint8_t *p = b + pos, *e = b + len;
            if (pos < len && b[pos] == '>' && (pos == 0 || b[pos - 1] == '\n')) cut = pos;
            else {
                while (p < e && (p = (const uint8_t *)memchr(p, '\n', (size_t)(e - p))) != nullptr) {
                    ++p;
                    if (p < e && *p == '>') { cut = (uint64_t)(p - b); break; }
                }
            }
        } else {
            // plain text: cut right after the newline that ends the line containing byte pos - 1
            const uint8_t *p = pos ? (const uint8_t *)memchr(b + pos - 1, '\n', (size_t)(len - pos + 1)) : b - 1;
            if (p != nullptr) cut = (uint64_t)(p - b) + 1;
        }
        cuts[r] = cut < cuts[r - 1] ? cuts[r - 1] : cut;
    }
}

}  // namespace aix

using namespace aix;

static thread_local std::string g_multi_error;

extern "C" {

int aix_multi_create(int n_dev, const int *dev_ids, aix_multi **out) {
    if (!out) return AIX_ERR_ARG;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
        cudaGetLastError();
        g_multi_error = "no usable CUDA device (libaindex_cuda has no CPU fallback)";
        return AIX_ERR_CUDA;
    }
    if (n_dev <= 0) n_dev = visible;
    if (n_dev > 16) n_dev = 16;
    aix_multi *mg = new aix_multi();
    for (int i = 0; i < n_dev; ++i) {
        aix_ctx *c = nullptr;
        int rc = aix_ctx_create(dev_ids ? dev_ids[i] : i, &c);
        if (rc != AIX_OK) {
            g_multi_error = aix_last_error(nullptr);
            aix_multi_destroy(mg);
            return rc;
        }
        mg->ctx.push_back(c);
    }
    mg->peer_ok = n_dev > 1;
    for (int i = 0; i < n_dev && mg->peer_ok; ++i) {
        cudaSetDevice(mg->ctx[i]->device);
        for (int j = 0; j < n_dev; ++j) {
            if (i == j || mg->ctx[i]->device == mg->ctx[j]->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, mg->ctx[i]->device, mg->ctx[j]->device);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(mg->ctx[j]->device, 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaErrorPeerAccessAlreadyEnabled) e = cudaSuccess;
            cudaGetLastError();
            if (e != cudaSuccess) mg->peer_ok = false;
        }
    }
    *out = mg;
    return AIX_OK;
}

void aix_multi_destroy(aix_multi *mg) {
    if (!mg) return;
    for (aix_ctx *c : mg->ctx) aix_ctx_destroy(c);
    delete mg;
}

int aix_multi_size(const aix_multi *mg) { return mg ? (int)mg->ctx.size() : 0; }
aix_ctx *aix_multi_ctx(aix_multi *mg, int i) { return (mg && i >= 0 && i < (int)mg->ctx.size()) ? mg->ctx[i] : nullptr; }
int aix_multi_peer_access(const aix_multi *mg) { return mg && mg->peer_ok ? 1 : 0; }
const char *aix_multi_last_error(const aix_multi *mg) { return mg ? mg->err.c_str() : g_multi_error.c_str(); }

}  // extern "C"

// shards[r] / lens[r]: the part of the input GPU r counts (host pointers, or device pointers on GPU r when
// src_is_device); tf_out == NULL stops after the exchange step (every GPU then holds its summed k-mer range in
// aix_count13_hist_dev()[r * 4^13 / n ...], the state a reduce-scatter leaves)
static int count13_multi_impl(aix_multi *mg, const aix_mphf *m0, const uint8_t *const *shards, const uint64_t *lens,
                              bool src_is_device, int fmt, uint64_t *tf_out, aix_count_stats *stats) {
    const int n = (int)mg->ctx.size();
    auto failed = [&](int rc, aix_ctx *c) {
        mg->err = c ? aix_last_error(c) : "aix_count13_multi failed";
        return rc;
    };
    if (AIX_TOTAL_13MERS % (uint64_t)n) { mg->err = "the number of GPUs must divide 4^13"; return AIX_ERR_ARG; }
    // one MPHF copy per GPU (20 MB each), from the host arrays of m0 (only the permutation needs it)
    std::vector<aix_mphf *> mph(n, nullptr);
    std::vector<int> rcs(n, AIX_OK);
    // ---- count: one host thread per GPU (the workers of count_kmers13.cpp:305-309) ----
    {
        std::vector<std::thread> th;
        for (int r = 0; r < n; ++r)
            th.emplace_back([&, r]() {
                aix_ctx *c = mg->ctx[r];
                int rc = AIX_OK;
                if (tf_out) rc = aix_mphf_upload(c, m0->n, m0->hash_domain, m0->seed, m0->words.data(), m0->n_words,
                                                 m0->block_ranks.data(), m0->n_blocks, &mph[r]);
                if (rc == AIX_OK) rc = aix_count13_begin(c);
                if (rc == AIX_OK && lens[r])
                    rc = src_is_device ? aix_count13_add_dev(c, shards[r], lens[r], fmt) : aix_count13_add(c, shards[r], lens[r], fmt);
                if (rc == AIX_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = c->fail(AIX_ERR_CUDA, "count13 multi: stream sync failed");
                rcs[r] = rc;
            });
        for (auto &x : th) x.join();
    }
    auto cleanup = [&]() {
        for (int r = 0; r < n; ++r) {
            aix_ctx *c = mg->ctx[r];
            if (!c->c13_peer_ipc) aix_count13_peers_close(c);
            aix_count13_end(c);
            if (mph[r]) aix_mphf_destroy(c, mph[r]);
        }
    };
    for (int r = 0; r < n; ++r)
        if (rcs[r] != AIX_OK) { int rc = failed(rcs[r], mg->ctx[r]); cleanup(); return rc; }
    const uint64_t step = AIX_TOTAL_13MERS / (uint64_t)n;
    int rc = AIX_OK;
    aix_ctx *bad = nullptr;
#define MG_TRY(c, expr) do { if (rc == AIX_OK) { rc = (expr); if (rc != AIX_OK) bad = (c); } } while (0)
#define MG_CUDA(c, expr) do { if (rc == AIX_OK) { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cudaGetLastError(); rc = (c)->fail(AIX_ERR_CUDA, "count13 multi %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); bad = (c); } } } while (0)
    // ---- the exchange step: GPU r ends up with the sum over GPUs of k-mer range r in its own hist64 ----
    if (n == 1) {
        MG_TRY(mg->ctx[0], aix_count13_flush(mg->ctx[0]));
    } else if (mg->peer_ok) {
        for (int r = 0; r < n; ++r) {
            aix_ctx *c = mg->ctx[r];
            aix_count13_peers_close(c);
            c->c13_peer_ipc = false;
            c->c13_n_peers = n;
            c->c13_my_rank = r;
            for (int p = 0; p < n; ++p) {
                c->c13_peer[p][0] = mg->ctx[p]->c13_hist32; c->c13_peer[p][1] = mg->ctx[p]->c13_hist64; c->c13_peer[p][2] = mg->ctx[p]->c13_stats_dev;
            }
            // in place: a thread reads entry v of every GPU, then writes entry v of its own hist64; nobody else reads that entry
            MG_TRY(c, aix_count13_reduce_peers_dev(c, r * step, (r + 1) * step, c->c13_hist64 + r * step));
        }
    } else {
        for (int r = 0; r < n; ++r) MG_TRY(mg->ctx[r], aix_count13_flush(mg->ctx[r]));
        for (int r = 0; r < n; ++r) MG_CUDA(mg->ctx[r], (cudaSetDevice(mg->ctx[r]->device), cudaStreamSynchronize(mg->ctx[r]->stream)));
        for (int r = 0; r < n && rc == AIX_OK; ++r) {
            aix_ctx *c = mg->ctx[r];
            void *stage = nullptr;
            MG_TRY(c, c->reserve(SCR_TMP1, step * 8, &stage));
            for (int p = 0; p < n && rc == AIX_OK; ++p) {
                if (p == r) continue;
                MG_CUDA(c, cudaSetDevice(c->device));
                MG_CUDA(c, cudaMemcpyPeerAsync(stage, c->device, mg->ctx[p]->c13_hist64 + r * step, mg->ctx[p]->device, step * 8, c->stream));
                if (rc == AIX_OK) {
                    add_u64_kernel<<<aix_grid(step, 256), 256, 0, c->stream>>>((unsigned long long *)(c->c13_hist64 + r * step), (const unsigned long long *)stage, step);
                    c->launches++;
                }
            }
        }
    }
    if (!tf_out) {  // count + exchange only
        for (int r = 0; r < n; ++r) MG_CUDA(mg->ctx[r], (cudaSetDevice(mg->ctx[r]->device), cudaStreamSynchronize(mg->ctx[r]->stream)));
        if (rc != AIX_OK) failed(rc, bad);
        for (int r = 0; r < n; ++r) {
            if (!mg->ctx[r]->c13_peer_ipc) aix_count13_peers_close(mg->ctx[r]);
            if (mph[r]) aix_mphf_destroy(mg->ctx[r], mph[r]);
        }
        return rc;
    }
    // ---- every GPU permutes its range into .tf.bin order (own 4^13 x u64 array), GPU 0 adds the arrays up ----
    std::vector<void *> tf_dev(n, nullptr);
    for (int r = 0; r < n; ++r) {
        aix_ctx *c = mg->ctx[r];
        MG_TRY(c, c->reserve(SCR_TMP0, AIX_TOTAL_13MERS * 8, &tf_dev[r]));
        MG_CUDA(c, cudaSetDevice(c->device));
        MG_CUDA(c, cudaMemsetAsync(tf_dev[r], 0, AIX_TOTAL_13MERS * 8, c->stream));
        MG_TRY(c, aix_count13_finish_dev(c, mph[r], r * step, (r + 1) * step, (uint64_t *)tf_dev[r]));
    }
    for (int r = 0; r < n; ++r) MG_CUDA(mg->ctx[r], (cudaSetDevice(mg->ctx[r]->device), cudaStreamSynchronize(mg->ctx[r]->stream)));
    {
        aix_ctx *c = mg->ctx[0];
        void *stage = nullptr;
        MG_TRY(c, c->reserve(SCR_OUT0, AIX_TOTAL_13MERS * 8, &stage));
        for (int p = 1; p < n && rc == AIX_OK; ++p) {
            MG_CUDA(c, cudaSetDevice(c->device));
            MG_CUDA(c, cudaMemcpyPeerAsync(stage, c->device, tf_dev[p], mg->ctx[p]->device, AIX_TOTAL_13MERS * 8, c->stream));
            if (rc == AIX_OK) {
                add_u64_kernel<<<aix_grid(AIX_TOTAL_13MERS, 256), 256, 0, c->stream>>>((unsigned long long *)tf_dev[0], (const unsigned long long *)stage, AIX_TOTAL_13MERS);
                c->launches++;
            }
        }
        MG_CUDA(c, cudaMemcpyAsync(tf_out, tf_dev[0], AIX_TOTAL_13MERS * 8, cudaMemcpyDeviceToHost, c->stream));
        MG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (rc == AIX_OK && stats) {
        aix_count_stats tot = {0, 0, 0, 0};
        for (int r = 0; r < n && rc == AIX_OK; ++r) {
            aix_ctx *c = mg->ctx[r];
            aix_count_stats st;
            MG_CUDA(c, cudaSetDevice(c->device));
            MG_TRY(c, aix_count13_stats(c, &st));
            uint64_t oor = 0;
            MG_CUDA(c, cudaMemcpy(&oor, c->c13_stats_dev + 3, 8, cudaMemcpyDeviceToHost));
            tot.sequences += st.sequences; tot.windows += st.windows;
            tot.valid += st.valid - oor;   // count_kmers13.cpp:153-156
            tot.invalid += st.invalid + oor;
        }
        *stats = tot;
    }
#undef MG_TRY
#undef MG_CUDA
    if (rc != AIX_OK) failed(rc, bad);
    cleanup();
    return rc;
}

extern "C" {

int aix_count13_multi(aix_multi *mg, const aix_mphf *m0, const uint8_t *bytes, uint64_t len, int fmt, uint64_t *tf_out,
                      aix_count_stats *stats) {
    if (!mg || !m0 || !tf_out || mg->ctx.empty() || (len && !bytes)) return AIX_ERR_ARG;
    const int n = (int)mg->ctx.size();
    if (n == 1) {
        int rc = aix_count13(mg->ctx[0], m0, bytes, len, fmt, tf_out, stats);
        if (rc != AIX_OK) mg->err = aix_last_error(mg->ctx[0]);
        return rc;
    }
    if (fmt == AIX_FMT_DETECT) fmt = (len == 0 || bytes[0] == '\n') ? AIX_FMT_PLAIN : bytes[0] == '>' ? AIX_FMT_FASTA : bytes[0] == '@' ? AIX_FMT_FASTQ : AIX_FMT_PLAIN;
    std::vector<uint64_t> cuts;
    shard_image(bytes, len, fmt, n, cuts);
    std::vector<const uint8_t *> shards(n);
    std::vector<uint64_t> lens(n);
    for (int r = 0; r < n; ++r) { shards[r] = bytes + cuts[r]; lens[r] = cuts[r + 1] - cuts[r]; }
    return count13_multi_impl(mg, m0, shards.data(), lens.data(), false, fmt, tf_out, stats);
}

int aix_count13_multi_dev(aix_multi *mg, const aix_mphf *m0, const uint8_t *const *shards_dev, const uint64_t *lens, int fmt,
                          uint64_t *tf_out, aix_count_stats *stats) {
    if (!mg || !m0 || !shards_dev || !lens || mg->ctx.empty()) return AIX_ERR_ARG;
    if (fmt == AIX_FMT_DETECT) { mg->err = "aix_count13_multi_dev needs an explicit format"; return AIX_ERR_ARG; }
    return count13_multi_impl(mg, m0, shards_dev, lens, true, fmt, tf_out, stats);
}

}  // extern "C"
