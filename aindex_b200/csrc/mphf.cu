// mphf.cu -- emphf minimal perfect hash: .pf (de)serialisation, upload into the B200
// layout, batched lookup / hash kernels, the 4^13 permutation table.
//
// Reference: src/emphf/mphf.hpp:79-113, base_hash.hpp:38-145, bitpair_vector.hpp:46-107,
// ranked_bitpair_vector.hpp:47-84; callers python_wrapper.cpp:629-642 (get_hash_values).
#include "aix_internal.cuh"
#include "batch_pipeline.cuh"

namespace aix {

// recs[w] = { words[w], block_ranks[w/16] + nonzero pairs of words[16*(w/16) .. w) }
__global__ void mphf_layout_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ block_ranks,
                                   uint64_t n_words, ulonglong2 *__restrict__ recs) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint64_t blk = w >> 4;
    uint64_t r = block_ranks[blk];
    for (uint64_t i = blk << 4; i < w; ++i) r += nonzero_pairs64(words[i]);
    recs[w] = make_ulonglong2(words[w], r);
}

// compact records: crecs[r] = { half-words 3r, 3r+1, 3r+2 of the bit-pair vector, rank of pair 48*r }
__global__ void mphf_layout_compact_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ block_ranks,
                                           uint64_t n_words, uint64_t n_recs, uint4 *__restrict__ crecs) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_recs) return;
    const uint64_t j = 3 * r;  // first half-word
    auto half = [&](uint64_t k) -> uint32_t {
        uint64_t w = k >> 1;
        if (w >= n_words) return 0u;
        return (k & 1) ? (uint32_t)(words[w] >> 32) : (uint32_t)words[w];
    };
    const uint64_t w = j >> 1, blk = w >> 4;
    uint64_t rank = w < n_words ? block_ranks[blk] : 0;
    for (uint64_t i = blk << 4; i < w && i < n_words; ++i) rank += nonzero_pairs64(words[i]);
    if (j & 1) rank += nonzero_pairs32(half(j - 1));
    crecs[r] = make_uint4(half(j), half(j + 1), half(j + 2), (uint32_t)rank);
}

static bool mphf_force_wide() {
    const char *e = getenv("AIX_MPHF_WIDE");  // test hook: exercise the wide records on small structures
    return e && atoi(e) != 0;
}

int mphf_build_layout(aix_ctx *ctx, aix_mphf *m) {
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t nw = m->n_words ? m->n_words : 1;
    uint64_t *words_dev = nullptr, *ranks_dev = nullptr;
    struct Staging {  // freed on every exit path (AIX_CUDA returns early)
        uint64_t *&a, *&b;
        ~Staging() { if (a) cudaFree(a); if (b) cudaFree(b); }
    } staging{words_dev, ranks_dev};
    const bool compact = m->bv_size < (1ull << 32) && m->hash_domain < (1ull << 31) && m->n < (1ull << 32) && !mphf_force_wide();
    if (compact) {
        const uint64_t n_recs = (2 * nw + 2) / 3 + 1;
        m->layout_bytes = n_recs * sizeof(uint4);
        AIX_CUDA(ctx, cudaMalloc(&m->crecs_dev, n_recs * sizeof(uint4)));
        AIX_CUDA(ctx, cudaMemsetAsync(m->crecs_dev, 0, n_recs * sizeof(uint4), ctx->stream));
        if (m->n_words) {
            AIX_CUDA(ctx, cudaMalloc(&words_dev, m->n_words * 8));
            AIX_CUDA(ctx, cudaMalloc(&ranks_dev, (m->n_blocks ? m->n_blocks : 1) * 8));
            AIX_CUDA(ctx, cudaMemcpyAsync(words_dev, m->words.data(), m->n_words * 8, cudaMemcpyHostToDevice, ctx->stream));
            AIX_CUDA(ctx, cudaMemcpyAsync(ranks_dev, m->block_ranks.data(), m->n_blocks * 8, cudaMemcpyHostToDevice, ctx->stream));
            mphf_layout_compact_kernel<<<aix_grid(n_recs, 256), 256, 0, ctx->stream>>>(words_dev, ranks_dev, m->n_words, n_recs, m->crecs_dev);
            AIX_LAUNCH_CHECK(ctx);
        }
        AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return AIX_OK;
    }
    m->layout_bytes = nw * sizeof(ulonglong2);
    AIX_CUDA(ctx, cudaMalloc(&m->recs_dev, nw * sizeof(ulonglong2)));
    AIX_CUDA(ctx, cudaMemsetAsync(m->recs_dev, 0, nw * sizeof(ulonglong2), ctx->stream));
    if (m->n_words) {
        AIX_CUDA(ctx, cudaMalloc(&words_dev, m->n_words * 8));
        AIX_CUDA(ctx, cudaMalloc(&ranks_dev, (m->n_blocks ? m->n_blocks : 1) * 8));
        AIX_CUDA(ctx, cudaMemcpyAsync(words_dev, m->words.data(), m->n_words * 8, cudaMemcpyHostToDevice, ctx->stream));
        AIX_CUDA(ctx, cudaMemcpyAsync(ranks_dev, m->block_ranks.data(), m->n_blocks * 8, cudaMemcpyHostToDevice, ctx->stream));
        mphf_layout_kernel<<<aix_grid(m->n_words, 256), 256, 0, ctx->stream>>>(words_dev, ranks_dev, m->n_words, m->recs_dev);
        AIX_LAUNCH_CHECK(ctx);
    }
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

// generic records (any stride / length): hashes the raw bytes straight from global memory
template <bool kLookup>
__global__ void records_hash_kernel(MphfDev m, uint64_t seed, const uint8_t *__restrict__ recs, uint32_t stride,
                                    const uint8_t *__restrict__ lens, uint64_t q, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t len = lens ? lens[i] : stride;
    if (len > stride) len = stride;
    uint64_t a, b, c;
    jenkins_bytes(kLookup ? m.seed : seed, recs + i * stride, len, a, b, c);
    if (kLookup) {
        out[i] = mphf_eval(m, a, b, c);
    } else {
        out[3 * i] = a; out[3 * i + 1] = b; out[3 * i + 2] = c;
    }
}

// perm13[v] = mphf(ASCII of v): one thread per 13-mer value, hash from registers
__global__ void perm13_kernel(MphfDev m, uint32_t *__restrict__ perm, uint32_t v_begin, uint32_t count) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint32_t v = v_begin + t;
    uint64_t id = mphf_lookup13(m, revcomp13(v));
    perm[t] = id > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)id;
}

}  // namespace aix

using namespace aix;

static int read_exact(FILE *f, void *dst, size_t bytes) { return fread(dst, 1, bytes, f) == bytes ? 0 : -1; }

extern "C" {

int aix_mphf_upload(aix_ctx *ctx, uint64_t n, uint64_t hash_domain, uint64_t seed, const uint64_t *words,
                    uint64_t n_words, const uint64_t *block_ranks, uint64_t n_blocks, aix_mphf **out) {
    if (!ctx || !out) return AIX_ERR_ARG;
    *out = nullptr;
    uint64_t bv = 3 * hash_domain;
    // an empty structure (hash_domain == 0) has no node to evaluate: fastmod by 0 would index the records with the raw hash
    if (hash_domain == 0) return ctx->fail(AIX_ERR_ARG, "empty mphf (hash_domain == 0)");
    if (n_words != (bv + 31) / 32 || n_blocks != (bv + 511) / 512)
        return ctx->fail(AIX_ERR_ARG, "mphf arrays do not match hash_domain (n_words=%llu n_blocks=%llu bv=%llu)",
                         (unsigned long long)n_words, (unsigned long long)n_blocks, (unsigned long long)bv);
    if ((n_words && !words) || (n_blocks && !block_ranks)) return ctx->fail(AIX_ERR_ARG, "null mphf arrays");
    aix_mphf *m = new aix_mphf();
    m->n = n; m->hash_domain = hash_domain; m->seed = seed; m->bv_size = bv;
    m->n_words = n_words; m->n_blocks = n_blocks;
    m->words.assign(words, words + n_words);
    m->block_ranks.assign(block_ranks, block_ranks + n_blocks);
    int rc = mphf_build_layout(ctx, m);
    if (rc != AIX_OK) {
        aix_mphf_destroy(ctx, m);
        return rc;
    }
    *out = m;
    return AIX_OK;
}

int aix_mphf_load_pf(aix_ctx *ctx, const char *pf_path, aix_mphf **out) {
    if (!ctx || !out || !pf_path) return AIX_ERR_ARG;
    *out = nullptr;
    FILE *f = fopen(pf_path, "rb");
    if (!f) return ctx->fail(AIX_ERR_IO, "cannot open hash file: %s", pf_path);
    uint64_t hdr[4];
    if (read_exact(f, hdr, sizeof hdr)) {
        fclose(f);
        return ctx->fail(AIX_ERR_IO, "short .pf header: %s", pf_path);
    }
    aix_mphf *m = new aix_mphf();
    m->n = hdr[0]; m->hash_domain = hdr[1]; m->seed = hdr[2]; m->bv_size = hdr[3];
    m->n_words = (m->bv_size + 31) / 32;
    m->n_blocks = (m->bv_size + 511) / 512;
    if (m->hash_domain == 0 || m->bv_size != 3 * m->hash_domain || m->n_words > (1ull << 40)) {
        fclose(f);
        delete m;
        return ctx->fail(AIX_ERR_IO, "corrupt .pf header: %s", pf_path);
    }
    m->words.resize(m->n_words);
    m->block_ranks.resize(m->n_blocks);
    if (read_exact(f, m->words.data(), m->n_words * 8) || read_exact(f, m->block_ranks.data(), m->n_blocks * 8)) {
        fclose(f);
        delete m;
        return ctx->fail(AIX_ERR_IO, "short .pf body: %s", pf_path);
    }
    fclose(f);
    int rc = mphf_build_layout(ctx, m);
    if (rc != AIX_OK) {
        aix_mphf_destroy(ctx, m);
        return rc;
    }
    *out = m;
    return AIX_OK;
}

int aix_mphf_save_pf(aix_ctx *ctx, const aix_mphf *m, const char *pf_path) {
    if (!ctx || !m || !pf_path) return AIX_ERR_ARG;
    FILE *f = fopen(pf_path, "wb");
    if (!f) return ctx->fail(AIX_ERR_IO, "cannot create %s", pf_path);
    uint64_t hdr[4] = {m->n, m->hash_domain, m->seed, m->bv_size};
    bool ok = fwrite(hdr, 8, 4, f) == 4 && fwrite(m->words.data(), 8, m->n_words, f) == m->n_words &&
              fwrite(m->block_ranks.data(), 8, m->n_blocks, f) == m->n_blocks;
    ok = (fclose(f) == 0) && ok;
    return ok ? AIX_OK : ctx->fail(AIX_ERR_IO, "short write: %s", pf_path);
}

void aix_mphf_destroy(aix_ctx *ctx, aix_mphf *m) {
    if (!m) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (m->recs_dev) cudaFree(m->recs_dev);
    if (m->crecs_dev) cudaFree(m->crecs_dev);
    delete m;
}

int aix_mphf_info(const aix_mphf *m, uint64_t info[6]) {
    if (!m || !info) return AIX_ERR_ARG;
    info[0] = m->n; info[1] = m->hash_domain; info[2] = m->seed;
    info[3] = m->bv_size; info[4] = m->n_words; info[5] = m->n_blocks;
    return AIX_OK;
}

int aix_mphf_arrays(const aix_mphf *m, uint64_t *words_out, uint64_t *block_ranks_out) {
    if (!m) return AIX_ERR_ARG;
    if (words_out) memcpy(words_out, m->words.data(), m->n_words * 8);
    if (block_ranks_out) memcpy(block_ranks_out, m->block_ranks.data(), m->n_blocks * 8);
    return AIX_OK;
}

int aix_mphf_lookup(aix_ctx *ctx, const aix_mphf *m, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                    uint64_t q, uint64_t *ids_out) {
    if (!ctx || !m) return AIX_ERR_ARG;
    MphfDev md = m->dev();
    return run_record_batches(ctx, recs, stride, lens, q, ids_out, 8,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *l, uint64_t nq, void *o) {
                                  records_hash_kernel<true><<<aix_grid(nq, 256), 256, 0, st>>>(md, 0, r, stride, l, nq, (uint64_t *)o);
                                  AIX_LAUNCH_CHECK(ctx);
                                  return AIX_OK;
                              });
}

int aix_jenkins64(aix_ctx *ctx, uint64_t seed, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                  uint64_t q, uint64_t *triples_out) {
    if (!ctx) return AIX_ERR_ARG;
    MphfDev md = {};
    return run_record_batches(ctx, recs, stride, lens, q, triples_out, 24,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *l, uint64_t nq, void *o) {
                                  records_hash_kernel<false><<<aix_grid(nq, 256), 256, 0, st>>>(md, seed, r, stride, l, nq, (uint64_t *)o);
                                  AIX_LAUNCH_CHECK(ctx);
                                  return AIX_OK;
                              });
}

int aix_perm13(aix_ctx *ctx, const aix_mphf *m, uint32_t *perm_out) {
    if (!ctx || !m || !perm_out) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *buf;
    AIX_TRY(ctx->reserve(SCR_TMP0, AIX_TOTAL_13MERS * 4, &buf));
    perm13_kernel<<<aix_grid(AIX_TOTAL_13MERS, 256), 256, 0, ctx->stream>>>(m->dev(), (uint32_t *)buf, 0u, (uint32_t)AIX_TOTAL_13MERS);
    AIX_LAUNCH_CHECK(ctx);
    AIX_CUDA(ctx, cudaMemcpyAsync(perm_out, buf, AIX_TOTAL_13MERS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

}  // extern "C"
