// batch_pipeline.cuh -- host-buffer entry points: chunk the records, and run
// H2D copy -> kernel -> D2H copy of consecutive chunks on two alternating streams so the
// PCIe transfers of chunk c+1 overlap the kernel of chunk c.
#pragma once
#include "aix_internal.cuh"

namespace aix {

constexpr size_t kSmallBatchBytes = 32u << 10;  // per region (records / lengths / results); +64 bytes of slack inside

// launch(stream, recs_dev, lens_dev_or_null, nq, out_dev) must enqueue the kernel(s) on `stream`.
template <typename Launch>
int run_record_batches(aix_ctx *ctx, const uint8_t *recs, uint32_t stride, const uint8_t *lens,
                       uint64_t q, void *out, size_t out_bytes_per_rec, Launch launch) {
    if (q == 0) return AIX_OK;
    if (!recs || !out || stride == 0) return ctx->fail(AIX_ERR_ARG, "null buffer or zero stride");
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    // small batches (a single get_tf_value call is a batch of one): no staged copies at all.  The records are
    // placed in a pinned, device-mapped buffer owned by the ctx, the kernel reads them and writes its results
    // through the mapping, and the host waits once: one launch + one synchronisation instead of three
    // synchronous pageable copies around the launch.
    if (q * (uint64_t)stride + 64 <= kSmallBatchBytes && q * out_bytes_per_rec <= kSmallBatchBytes) {
        if (!ctx->small_host) {
            cudaError_t e = cudaHostAlloc(&ctx->small_host, 3 * kSmallBatchBytes, cudaHostAllocMapped);
            if (e != cudaSuccess) {
                cudaGetLastError();
                ctx->small_host = nullptr;
            }
        }
        if (ctx->small_host) {
            uint8_t *h_in = (uint8_t *)ctx->small_host, *h_len = h_in + kSmallBatchBytes, *h_out = h_len + kSmallBatchBytes;
            memcpy(h_in, recs, q * stride);
            memset(h_in + q * stride, 0, 64);  // the fixed-stride kernels read whole 16-byte vectors
            if (lens) memcpy(h_len, lens, q);
            int rc = launch(ctx->stream, h_in, lens ? h_len : nullptr, q, h_out);
            if (rc != AIX_OK) return rc;
            AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            memcpy(out, h_out, q * out_bytes_per_rec);
            return AIX_OK;
        }
    }
    const uint64_t target_bytes = 64ull << 20;
    uint64_t qc = target_bytes / stride;
    qc = (qc + 4095) & ~4095ull;
    if (qc > q) qc = q;
    void *in_dev[2], *len_dev[2] = {nullptr, nullptr}, *out_dev[2];
    for (int s = 0; s < 2; ++s) {
        AIX_TRY(ctx->reserve(SCR_IN0 + s, qc * stride + 64, &in_dev[s]));
        AIX_TRY(ctx->reserve(SCR_OUT0 + s, qc * out_bytes_per_rec + 64, &out_dev[s]));
        if (lens) AIX_TRY(ctx->reserve(SCR_LEN0 + s, qc + 64, &len_dev[s]));
        if (q <= qc) break;  // single chunk: one buffer set is enough
    }
    uint64_t done = 0;
    int c = 0;
    while (done < q) {
        uint64_t nq = q - done < qc ? q - done : qc;
        int s = c & 1;
        cudaStream_t st = ctx->xfer[s];
        AIX_CUDA(ctx, cudaMemcpyAsync(in_dev[s], recs + done * stride, nq * stride, cudaMemcpyHostToDevice, st));
        if (lens) AIX_CUDA(ctx, cudaMemcpyAsync(len_dev[s], lens + done, nq, cudaMemcpyHostToDevice, st));
        int rc = launch(st, (const uint8_t *)in_dev[s], (const uint8_t *)len_dev[s], nq, out_dev[s]);
        if (rc != AIX_OK) return rc;
        AIX_CUDA(ctx, cudaMemcpyAsync((uint8_t *)out + done * out_bytes_per_rec, out_dev[s],
                                      nq * out_bytes_per_rec, cudaMemcpyDeviceToHost, st));
        done += nq;
        ++c;
    }
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[0]));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[1]));
    return AIX_OK;
}

}  // namespace aix
