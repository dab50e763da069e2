// codec.cu -- standalone 2-bit codec kernels (K1): encode / decode / reverse complement of
// k-mer batches, dna_bitset packing, rolling forward / reverse-complement k-mers of a
// reads buffer.
//
// Reference: src/kmers.cpp:12-85 (get_dna23_bitset / get_dna13_bitset), :89-257
// (get_bitset_dna23 / 13), :355-388 (reverseDNA); src/dna_bitseq.hpp:22-61; the rolling
// form replaces the per-window re-encoding of src/hash.cpp:1017-1032.
#include "aix_internal.cuh"
#include "batch_pipeline.cuh"

namespace aix {

__global__ void encode_kernel(const uint8_t *__restrict__ recs, uint32_t stride, const uint8_t *__restrict__ lens,
                              uint64_t q, int k, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t len = lens ? lens[i] : stride;
    if (len > stride) len = stride;
    const uint8_t *p = recs + i * stride;
    uint64_t u = 0;
    for (int j = 0; j < k; ++j) {
        uint32_t ch = (uint32_t)j < len ? __ldg(p + j) : 0u;  // kmers.cpp:17-23: always k characters
        u = (u << 2) | base_code_strict(ch);
    }
    out[i] = u;
}

__global__ void decode_kernel(const uint64_t *__restrict__ values, uint64_t q, int k, uint8_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint64_t x = values[i];
    for (int j = k - 1; j >= 0; --j) {  // kmers.cpp:93-113
        out[i * k + j] = (uint8_t)((0x54474341u >> (8 * (x & 3))) & 0xFFu);
        x >>= 2;
    }
}

__global__ void revcomp_kernel(const uint64_t *__restrict__ values, uint64_t q, int k, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    out[i] = k == 23 ? revcomp23(values[i]) : (uint64_t)revcomp13((uint32_t)values[i]);
}

// dna_bitset ctor: 4 bases per byte, first base in bits 7:6, anything but ACGT -> A
__global__ void pack2bit_kernel(const uint8_t *__restrict__ seq, uint64_t len, uint8_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nbytes = (len + 3) / 4;
    if (i >= nbytes) return;
    uint32_t b = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint64_t p = i * 4 + j;
        uint32_t ch = p < len ? seq[p] : 0u;
        b |= base_code_strict(ch) << (6 - 2 * j);
    }
    out[i] = (uint8_t)b;
}

// rolling k-mers: each thread owns kRoll consecutive window starts, builds the first window
// from k bytes and then rolls both strands one base at a time.
constexpr int kRoll = 16;
template <int K>
__global__ void rolling_kernel(const uint8_t *__restrict__ bytes, uint64_t len, uint64_t *__restrict__ fwd,
                               uint64_t *__restrict__ rc, uint8_t *__restrict__ valid) {
    const uint64_t n_win = len - K + 1;
    uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kRoll;
    if (i0 >= n_win) return;
    const uint64_t mask = (1ULL << (2 * K)) - 1;
    uint64_t f = 0, r = 0;
    uint32_t bad = 0;  // bit j set: one of the last K bytes (age j) is not upper-case ACGT
    for (int j = 0; j < K - 1; ++j) {
        uint32_t ch = bytes[i0 + j];
        uint64_t c = base_code_strict(ch);
        f = (f << 2) | c;
        r = (r >> 2) | ((3 - c) << (2 * (K - 1)));
        bad = (bad << 1) | (is_acgt_upper(ch) ? 0u : 1u);
    }
    for (int t = 0; t < kRoll && i0 + t < n_win; ++t) {
        uint32_t ch = bytes[i0 + t + K - 1];
        uint64_t c = base_code_strict(ch);
        f = ((f << 2) | c) & mask;
        r = (r >> 2) | ((3 - c) << (2 * (K - 1)));
        bad = ((bad << 1) | (is_acgt_upper(ch) ? 0u : 1u)) & ((1u << K) - 1);
        // reverseDNA of the strict forward value: a non-ACGT byte encodes as A and complements to T
        if (fwd) fwd[i0 + t] = f;
        if (rc) rc[i0 + t] = r;
        if (valid) valid[i0 + t] = bad ? 0 : 1;
    }
}

}  // namespace aix

using namespace aix;

extern "C" {

int aix_encode_kmers(aix_ctx *ctx, const uint8_t *recs, uint32_t stride, const uint8_t *lens, uint64_t q, int k,
                     uint64_t *out) {
    if (!ctx) return AIX_ERR_ARG;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    return run_record_batches(ctx, recs, stride, lens, q, out, 8,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *l, uint64_t nq, void *o) {
                                  encode_kernel<<<aix_grid(nq, 256), 256, 0, st>>>(r, stride, l, nq, k, (uint64_t *)o);
                                  AIX_LAUNCH_CHECK(ctx);
                                  return AIX_OK;
                              });
}

static int values_op(aix_ctx *ctx, const uint64_t *values, uint64_t q, int k, void *out, size_t out_per, bool decode) {
    if (!ctx) return AIX_ERR_ARG;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    return run_record_batches(ctx, (const uint8_t *)values, 8, nullptr, q, out, out_per,
                              [&](cudaStream_t st, const uint8_t *r, const uint8_t *, uint64_t nq, void *o) {
                                  if (decode) decode_kernel<<<aix_grid(nq, 256), 256, 0, st>>>((const uint64_t *)r, nq, k, (uint8_t *)o);
                                  else revcomp_kernel<<<aix_grid(nq, 256), 256, 0, st>>>((const uint64_t *)r, nq, k, (uint64_t *)o);
                                  AIX_LAUNCH_CHECK(ctx);
                                  return AIX_OK;
                              });
}

int aix_decode_kmers(aix_ctx *ctx, const uint64_t *values, uint64_t q, int k, uint8_t *out) {
    return values_op(ctx, values, q, k, out, (size_t)(k > 0 ? k : 1), true);
}

int aix_revcomp_kmers(aix_ctx *ctx, const uint64_t *values, uint64_t q, int k, uint64_t *out) {
    return values_op(ctx, values, q, k, out, 8, false);
}

int aix_pack_2bit(aix_ctx *ctx, const uint8_t *seq, uint64_t len, uint8_t *packed_out) {
    if (!ctx) return AIX_ERR_ARG;
    if (len == 0) return AIX_OK;
    if (!seq || !packed_out) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *in, *out;
    uint64_t nbytes = (len + 3) / 4;
    AIX_TRY(ctx->reserve(SCR_IN0, len, &in));
    AIX_TRY(ctx->reserve(SCR_OUT0, nbytes, &out));
    AIX_CUDA(ctx, cudaMemcpyAsync(in, seq, len, cudaMemcpyHostToDevice, ctx->stream));
    pack2bit_kernel<<<aix_grid(nbytes, 256), 256, 0, ctx->stream>>>((const uint8_t *)in, len, (uint8_t *)out);
    AIX_LAUNCH_CHECK(ctx);
    AIX_CUDA(ctx, cudaMemcpyAsync(packed_out, out, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

int aix_rolling_kmers(aix_ctx *ctx, const uint8_t *bytes, uint64_t len, int k, uint64_t *fwd_out, uint64_t *rc_out,
                      uint8_t *valid_out) {
    if (!ctx) return AIX_ERR_ARG;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (len < (uint64_t)k) return AIX_OK;
    if (!bytes) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_win = len - k + 1;
    void *in, *f = nullptr, *r = nullptr, *v = nullptr;
    AIX_TRY(ctx->reserve(SCR_IN0, len, &in));
    if (fwd_out) AIX_TRY(ctx->reserve(SCR_OUT0, n_win * 8, &f));
    if (rc_out) AIX_TRY(ctx->reserve(SCR_OUT1, n_win * 8, &r));
    if (valid_out) AIX_TRY(ctx->reserve(SCR_LEN0, n_win, &v));
    AIX_CUDA(ctx, cudaMemcpyAsync(in, bytes, len, cudaMemcpyHostToDevice, ctx->stream));
    unsigned grid = aix_grid((n_win + kRoll - 1) / kRoll, 128);
    if (k == 23) rolling_kernel<23><<<grid, 128, 0, ctx->stream>>>((const uint8_t *)in, len, (uint64_t *)f, (uint64_t *)r, (uint8_t *)v);
    else rolling_kernel<13><<<grid, 128, 0, ctx->stream>>>((const uint8_t *)in, len, (uint64_t *)f, (uint64_t *)r, (uint8_t *)v);
    AIX_LAUNCH_CHECK(ctx);
    if (fwd_out) AIX_CUDA(ctx, cudaMemcpyAsync(fwd_out, f, n_win * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (rc_out) AIX_CUDA(ctx, cudaMemcpyAsync(rc_out, r, n_win * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid_out) AIX_CUDA(ctx, cudaMemcpyAsync(valid_out, v, n_win, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

}  // extern "C"
