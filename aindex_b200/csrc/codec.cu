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

// "ACGT"[code] for the four letter codes of a word (two PRMTs; same construction as query23.cuh::expect_acgt4)
__device__ __forceinline__ uint32_t expect_acgt4_codec(uint32_t c4) {
    const uint32_t t = c4 | (c4 >> 4);
    return __byte_perm(0x54474341u, 0u, __byte_perm(t, 0u, 0x4420));
}

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

// ---- dna_bitset packing (dna_bitseq.hpp:22-61): 4 bases per byte, first base in bits 7:6, anything but
//      upper-case ACGT -> A --------------------------------------------------------------------------------
// strict 2-bit codes of the 4 bytes of a word, MSB-first in one byte: byte j of the word -> bits 7-2j..6-2j
__device__ __forceinline__ uint32_t pack4_strict(uint32_t w) {
    uint32_t c4 = ((w >> 1) ^ (w >> 2)) & 0x03030303u;       // letter code, meaningful for ACGT only
    const uint32_t diff = expect_acgt4_codec(c4) ^ w;         // zero byte <=> the byte is exactly "ACGT"[code]
    const uint32_t nz = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;  // 0x80 in every byte that differs
    c4 &= ~((nz >> 7) * 3u);                                  // anything else encodes as A (0)
    return (c4 * 0x40100401u) >> 24;                          // c0<<6 | c1<<4 | c2<<2 | c3
}

// generic form: one output byte per thread, byte loads (any alignment, ragged tail)
__global__ void pack2bit_kernel(const uint8_t *__restrict__ seq, uint64_t first_out, uint64_t len, uint8_t *__restrict__ out) {
    uint64_t i = first_out + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
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

// vector form: 128-bit coalesced loads (16 bases) and 32-bit coalesced stores (4 packed bytes); a warp turns 512
// contiguous input bytes into 128 contiguous output bytes per step.  A thread takes kPackUnroll vectors, one CTA width
// apart, and issues their loads together: with one 16-byte load per thread an SM has 32 KB in flight, about what HBM's
// latency x bandwidth asks of it and no more (0.60 of the copy peak measured).  seq 16-byte aligned, out 4-byte aligned.
constexpr int kPackUnroll = 4;
__global__ void __launch_bounds__(256) pack2bit_vec_kernel(const uint4 *__restrict__ seq, uint64_t n_vec, uint32_t *__restrict__ out) {
    const uint64_t i0 = (uint64_t)blockIdx.x * (256 * kPackUnroll) + threadIdx.x;
    uint4 v[kPackUnroll];
#pragma unroll
    for (int j = 0; j < kPackUnroll; ++j) {
        const uint64_t i = i0 + (uint64_t)j * 256;
        v[j] = i < n_vec ? __ldcs(seq + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < kPackUnroll; ++j) {
        const uint64_t i = i0 + (uint64_t)j * 256;
        if (i < n_vec)
            __stcs(out + i, pack4_strict(v[j].x) | (pack4_strict(v[j].y) << 8) | (pack4_strict(v[j].z) << 16) | (pack4_strict(v[j].w) << 24));
    }
}

// dna_bitset::ukmer(pos, k) (dna_bitseq.hpp:124-151): the 2k-bit big-endian substring of the packed stream that
// starts at base `pos`.  Bases at or past n_bases read as A (the reference reads past its buffer there).
__global__ void ukmer_kernel(const uint8_t *__restrict__ packed, uint64_t n_bases, const uint64_t *__restrict__ pos, uint64_t q,
                             int k, uint64_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    const uint64_t p = pos[i], nbytes = (n_bases + 3) / 4;
    const uint64_t b0 = p >> 2;
    // nine bytes cover 2k <= 64 bits at any of the four bit offsets
    uint64_t hi = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) hi = (hi << 8) | (b0 + j < nbytes ? (uint64_t)packed[b0 + j] : 0ull);
    const uint64_t lo = b0 + 8 < nbytes ? (uint64_t)packed[b0 + 8] : 0ull;
    const uint32_t off = (uint32_t)(p & 3) * 2u;                       // bits to drop at the top
    uint64_t x = off ? ((hi << off) | (lo >> (8 - off))) : hi;         // 64 bits starting at base p
    x = k >= 32 ? x : (x >> (64 - 2 * k));
    // bases past the end of the sequence (inside the last byte they are already 0 = A)
    if (p + (uint64_t)k > n_bases) {
        const uint64_t keep = p < n_bases ? n_bases - p : 0;            // bases that exist
        x = keep ? (x >> (2 * ((uint64_t)k - keep))) << (2 * ((uint64_t)k - keep)) : 0ull;
    }
    out[i] = x;
}

// ---- rolling forward / reverse-complement k-mers of a byte stream (the per-window re-encoding of
//      hash.cpp:1017-1032, rolled) ----------------------------------------------------------------------------
// Each thread owns 16 consecutive window starts = one 128-bit coalesced load; the K-1 bytes that follow come from the
// next two lanes by shuffle (the last lanes of a warp take them from two extra vectors that lanes 0 and 1 load).  The
// outputs of a warp (512 windows) are staged in shared memory and written as contiguous 256-byte rows.
constexpr int kRoll = 16;
constexpr int kRollWarps = 4;
template <int K>
__global__ void __launch_bounds__(kRollWarps * 32) rolling_vec_kernel(const uint4 *__restrict__ bytes, uint64_t len, uint64_t n_vec,
                                                                     uint64_t *__restrict__ fwd, uint64_t *__restrict__ rc,
                                                                     uint8_t *__restrict__ valid) {
    __shared__ uint64_t stage[kRollWarps][32 * (kRoll + 1)];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * kRollWarps + wid) * 32;  // first vector of this warp
    if (warp0 >= n_vec) return;
    const uint64_t n_win = len - K + 1;
    const uint4 zero = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);  // past the end: never part of a window that is written
    const uint64_t vi = warp0 + lane;
    const uint4 own = vi < n_vec ? __ldcs(bytes + vi) : zero;
    const uint4 ext = (lane < 2 && warp0 + 32 + lane < n_vec) ? __ldcs(bytes + warp0 + 32 + lane) : zero;
    uint32_t w[12];  // 48 bytes: own 16, next 16, 16 after that (K - 1 <= 22 of the last 32 are used)
    w[0] = own.x; w[1] = own.y; w[2] = own.z; w[3] = own.w;
    {
        const unsigned s1 = (lane + 1) & 31u, s2 = (lane + 2) & 31u;
        const uint32_t a0 = __shfl_sync(0xFFFFFFFFu, own.x, s1), a1 = __shfl_sync(0xFFFFFFFFu, own.y, s1),
                       a2 = __shfl_sync(0xFFFFFFFFu, own.z, s1), a3 = __shfl_sync(0xFFFFFFFFu, own.w, s1);
        const uint32_t e0 = __shfl_sync(0xFFFFFFFFu, ext.x, s1), e1 = __shfl_sync(0xFFFFFFFFu, ext.y, s1),
                       e2 = __shfl_sync(0xFFFFFFFFu, ext.z, s1), e3 = __shfl_sync(0xFFFFFFFFu, ext.w, s1);
        const bool wrap1 = lane == 31u;
        w[4] = wrap1 ? e0 : a0; w[5] = wrap1 ? e1 : a1; w[6] = wrap1 ? e2 : a2; w[7] = wrap1 ? e3 : a3;
        const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, own.x, s2), b1 = __shfl_sync(0xFFFFFFFFu, own.y, s2),
                       b2 = __shfl_sync(0xFFFFFFFFu, own.z, s2), b3 = __shfl_sync(0xFFFFFFFFu, own.w, s2);
        const uint32_t f0 = __shfl_sync(0xFFFFFFFFu, ext.x, s2), f1 = __shfl_sync(0xFFFFFFFFu, ext.y, s2),
                       f2 = __shfl_sync(0xFFFFFFFFu, ext.z, s2), f3 = __shfl_sync(0xFFFFFFFFu, ext.w, s2);
        const bool wrap2 = lane >= 30u;
        w[8] = wrap2 ? f0 : b0; w[9] = wrap2 ? f1 : b1; w[10] = wrap2 ? f2 : b2; w[11] = wrap2 ? f3 : b3;
    }
    auto byte_at = [&](int j) -> uint32_t { return (w[j >> 2] >> (8 * (j & 3))) & 0xFFu; };
    const uint64_t mask = (1ULL << (2 * K)) - 1;
    uint64_t f = 0, r = 0;
    uint32_t bad = 0;  // bit j set: one of the last K bytes (age j) is not upper-case ACGT
#pragma unroll
    for (int j = 0; j < K - 1; ++j) {
        const uint32_t ch = byte_at(j);
        const uint64_t c = base_code_strict(ch);
        f = (f << 2) | c;
        r = (r >> 2) | ((3 - c) << (2 * (K - 1)));
        bad = (bad << 1) | (is_acgt_upper(ch) ? 0u : 1u);
    }
    uint64_t fo[kRoll], ro[kRoll];
    uint32_t vbits = 0;
#pragma unroll
    for (int t = 0; t < kRoll; ++t) {
        const uint32_t ch = byte_at(t + K - 1);
        const uint64_t c = base_code_strict(ch);
        f = ((f << 2) | c) & mask;
        r = (r >> 2) | ((3 - c) << (2 * (K - 1)));
        bad = ((bad << 1) | (is_acgt_upper(ch) ? 0u : 1u)) & ((1u << K) - 1);
        fo[t] = f; ro[t] = r;
        vbits |= (bad ? 0u : 1u) << t;
    }
    // staged, coalesced writes: window e of the warp (0..511) sits at stage[(e / 16) * 17 + e % 16]
    uint64_t *st = stage[wid];
    const uint64_t win0 = warp0 * kRoll;  // first window start of the warp
    for (int pass = 0; pass < 2; ++pass) {
        uint64_t *dst = pass == 0 ? fwd : rc;
        if (dst == nullptr) continue;
#pragma unroll
        for (int t = 0; t < kRoll; ++t) st[lane * (kRoll + 1) + t] = pass == 0 ? fo[t] : ro[t];
        __syncwarp();
#pragma unroll
        for (int row = 0; row < kRoll; ++row) {
            const uint32_t e = row * 32 + lane;
            const uint64_t wi = win0 + e;
            if (wi < n_win) __stcs(dst + wi, st[(e >> 4) * (kRoll + 1) + (e & 15u)]);
        }
        __syncwarp();
    }
    if (valid != nullptr) {
        // 16 validity bytes per thread = one 128-bit store when the whole group exists
        const uint64_t wi = win0 + (uint64_t)lane * kRoll;
        if (wi + kRoll <= n_win && ((uintptr_t)(valid + wi) & 15) == 0) {
            uint32_t o[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t n4 = (vbits >> (4 * g)) & 15u;
                o[g] = (n4 & 1u) | ((n4 & 2u) << 7) | ((n4 & 4u) << 14) | ((n4 & 8u) << 21);
            }
            __stcs(reinterpret_cast<uint4 *>(valid + wi), make_uint4(o[0], o[1], o[2], o[3]));
        } else {
            for (int t = 0; t < kRoll; ++t)
                if (wi + t < n_win) valid[wi + t] = (uint8_t)((vbits >> t) & 1u);
        }
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

// device form: vector kernel over the 16-byte-aligned bulk, generic kernel for an unaligned buffer or the ragged tail
static int pack2bit_launch(aix_ctx *ctx, cudaStream_t st, const uint8_t *seq_dev, uint64_t len, uint8_t *packed_dev) {
    const uint64_t nbytes = (len + 3) / 4;
    uint64_t done_out = 0;
    if ((((uintptr_t)seq_dev) & 15) == 0 && (((uintptr_t)packed_dev) & 3) == 0 && len >= 16) {
        const uint64_t n_vec = len / 16;
        pack2bit_vec_kernel<<<aix_grid(n_vec, 256 * kPackUnroll), 256, 0, st>>>((const uint4 *)seq_dev, n_vec, (uint32_t *)packed_dev);
        AIX_LAUNCH_CHECK(ctx);
        done_out = n_vec * 4;
    }
    if (done_out < nbytes) {
        pack2bit_kernel<<<aix_grid(nbytes - done_out, 256), 256, 0, st>>>(seq_dev, done_out, len, packed_dev);
        AIX_LAUNCH_CHECK(ctx);
    }
    return AIX_OK;
}

int aix_pack_2bit_dev(aix_ctx *ctx, const uint8_t *seq_dev, uint64_t len, uint8_t *packed_dev) {
    if (!ctx) return AIX_ERR_ARG;
    if (len == 0) return AIX_OK;
    if (!seq_dev || !packed_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return pack2bit_launch(ctx, ctx->stream, seq_dev, len, packed_dev);
}

int aix_pack_2bit(aix_ctx *ctx, const uint8_t *seq, uint64_t len, uint8_t *packed_out) {
    if (!ctx) return AIX_ERR_ARG;
    if (len == 0) return AIX_OK;
    if (!seq || !packed_out) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    // chunks of 256 MiB of bases (a multiple of 16, so every chunk but the last packs to whole bytes), two streams
    const uint64_t chunk = 256ull << 20;
    uint64_t done = 0;
    int c = 0;
    while (done < len) {
        const uint64_t n = len - done < chunk ? len - done : chunk;
        const int b = c & 1;
        cudaStream_t st = ctx->xfer[b];
        void *in, *out;
        AIX_TRY(ctx->reserve(SCR_IN0 + b, n + 64, &in));
        AIX_TRY(ctx->reserve(SCR_OUT0 + b, (n + 3) / 4 + 64, &out));
        AIX_CUDA(ctx, cudaMemcpyAsync(in, seq + done, n, cudaMemcpyHostToDevice, st));
        AIX_TRY(pack2bit_launch(ctx, st, (const uint8_t *)in, n, (uint8_t *)out));
        AIX_CUDA(ctx, cudaMemcpyAsync(packed_out + done / 4, out, (n + 3) / 4, cudaMemcpyDeviceToHost, st));
        done += n;
        ++c;
    }
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[0]));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[1]));
    return AIX_OK;
}

int aix_ukmers_dev(aix_ctx *ctx, const uint8_t *packed_dev, uint64_t n_bases, const uint64_t *pos_dev, uint64_t q, int k,
                   uint64_t *out_dev) {
    if (!ctx) return AIX_ERR_ARG;
    if (k < 1 || k > 32) return ctx->fail(AIX_ERR_ARG, "ukmer: k must be in 1..32");
    if (q == 0) return AIX_OK;
    if (!packed_dev || !pos_dev || !out_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    ukmer_kernel<<<aix_grid(q, 256), 256, 0, ctx->stream>>>(packed_dev, n_bases, pos_dev, q, k, out_dev);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_ukmers(aix_ctx *ctx, const uint8_t *packed, uint64_t n_bases, const uint64_t *pos, uint64_t q, int k, uint64_t *out) {
    if (!ctx) return AIX_ERR_ARG;
    if (k < 1 || k > 32) return ctx->fail(AIX_ERR_ARG, "ukmer: k must be in 1..32");
    if (q == 0) return AIX_OK;
    if (!packed || !pos || !out) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *pk, *ps, *o;
    const uint64_t nbytes = (n_bases + 3) / 4;
    AIX_TRY(ctx->reserve(SCR_IN0, nbytes + 64, &pk));
    AIX_TRY(ctx->reserve(SCR_IN1, q * 8, &ps));
    AIX_TRY(ctx->reserve(SCR_OUT0, q * 8, &o));
    AIX_CUDA(ctx, cudaMemcpyAsync(pk, packed, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    AIX_CUDA(ctx, cudaMemcpyAsync(ps, pos, q * 8, cudaMemcpyHostToDevice, ctx->stream));
    AIX_TRY(aix_ukmers_dev(ctx, (const uint8_t *)pk, n_bases, (const uint64_t *)ps, q, k, (uint64_t *)o));
    AIX_CUDA(ctx, cudaMemcpyAsync(out, o, q * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

// bytes_dev must be 16-byte aligned and readable up to the next multiple of 16 past len (bytes past len are ignored)
static int rolling_launch(aix_ctx *ctx, cudaStream_t st, const uint8_t *bytes_dev, uint64_t len, int k, uint64_t *f, uint64_t *r,
                          uint8_t *v) {
    const uint64_t n_vec = (len + 15) / 16;
    const unsigned grid = aix_grid((n_vec + 31) / 32, kRollWarps);
    if (k == 23) rolling_vec_kernel<23><<<grid, kRollWarps * 32, 0, st>>>((const uint4 *)bytes_dev, len, n_vec, f, r, v);
    else rolling_vec_kernel<13><<<grid, kRollWarps * 32, 0, st>>>((const uint4 *)bytes_dev, len, n_vec, f, r, v);
    AIX_LAUNCH_CHECK(ctx);
    return AIX_OK;
}

int aix_rolling_kmers_dev(aix_ctx *ctx, const uint8_t *bytes_dev, uint64_t len, int k, uint64_t *fwd_dev, uint64_t *rc_dev,
                          uint8_t *valid_dev) {
    if (!ctx) return AIX_ERR_ARG;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (len < (uint64_t)k) return AIX_OK;
    if (!bytes_dev) return ctx->fail(AIX_ERR_ARG, "null buffer");
    if (((uintptr_t)bytes_dev) & 15) return ctx->fail(AIX_ERR_ARG, "rolling k-mers: the device image must be 16-byte aligned");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    return rolling_launch(ctx, ctx->stream, bytes_dev, len, k, fwd_dev, rc_dev, valid_dev);
}

int aix_rolling_kmers(aix_ctx *ctx, const uint8_t *bytes, uint64_t len, int k, uint64_t *fwd_out, uint64_t *rc_out,
                      uint8_t *valid_out) {
    if (!ctx) return AIX_ERR_ARG;
    if (k != 13 && k != 23) return ctx->fail(AIX_ERR_ARG, "k must be 13 or 23");
    if (len < (uint64_t)k) return AIX_OK;
    if (!bytes) return ctx->fail(AIX_ERR_ARG, "null buffer");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_win = len - k + 1;
    // chunks of 16 Mi window starts (17 B of output per window): chunk c covers windows [w0, w0 + nw) and needs the
    // bytes [w0, w0 + nw + k - 1); w0 is a multiple of 16, so every chunk starts at an aligned vector of the image
    uint64_t chunk = 16ull << 20;
    if (const char *e = getenv("AIX_ROLLING_CHUNK")) {  // test hook: force several chunks on a small input
        uint64_t v = strtoull(e, nullptr, 10);
        if (v >= 16) chunk = v & ~15ull;
    }
    cudaStream_t st = ctx->stream;
    for (uint64_t w0 = 0; w0 < n_win; w0 += chunk) {
        const uint64_t nw = n_win - w0 < chunk ? n_win - w0 : chunk;
        const uint64_t nb = nw + k - 1;
        void *in, *f = nullptr, *r = nullptr, *v = nullptr;
        AIX_TRY(ctx->reserve(SCR_IN0, nb + 64, &in));
        if (fwd_out) AIX_TRY(ctx->reserve(SCR_OUT0, nw * 8, &f));
        if (rc_out) AIX_TRY(ctx->reserve(SCR_OUT1, nw * 8, &r));
        if (valid_out) AIX_TRY(ctx->reserve(SCR_LEN0, nw + 64, &v));
        AIX_CUDA(ctx, cudaMemcpyAsync(in, bytes + w0, nb, cudaMemcpyHostToDevice, st));
        AIX_TRY(rolling_launch(ctx, st, (const uint8_t *)in, nb, k, (uint64_t *)f, (uint64_t *)r, (uint8_t *)v));
        if (fwd_out) AIX_CUDA(ctx, cudaMemcpyAsync(fwd_out + w0, f, nw * 8, cudaMemcpyDeviceToHost, st));
        if (rc_out) AIX_CUDA(ctx, cudaMemcpyAsync(rc_out + w0, r, nw * 8, cudaMemcpyDeviceToHost, st));
        if (valid_out) AIX_CUDA(ctx, cudaMemcpyAsync(valid_out + w0, v, nw, cudaMemcpyDeviceToHost, st));
        AIX_CUDA(ctx, cudaStreamSynchronize(st));  // the scratch buffers are reused by the next chunk
    }
    return AIX_OK;
}

}  // extern "C"
