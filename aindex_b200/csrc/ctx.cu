// ctx.cu -- context, error reporting, pinned host memory.
#include "aix_internal.cuh"

static thread_local std::string g_create_error;

extern "C" {

const char *aix_version(void) { return "aindex_b200 0.1 (sm_100a)"; }

int aix_ctx_create(int device, aix_ctx **out) {
    if (!out) return AIX_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = std::string("no usable CUDA device: ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libaindex_cuda has no CPU fallback)";
        return AIX_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "device index out of range";
        return AIX_ERR_ARG;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return AIX_ERR_CUDA;
    }
    aix_ctx *ctx = new aix_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->xfer[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->xfer[1], cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "cudaStreamCreate failed";
        delete ctx;
        return AIX_ERR_CUDA;
    }
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
            uint64_t keep = ~0ULL;
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
            ctx->pool = nullptr;  // builders fall back to cudaMalloc / cudaFree
        }
    }
    for (auto &ev : ctx->ev) {
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            g_create_error = "cudaEventCreate failed";
            delete ctx;
            return AIX_ERR_CUDA;
        }
    }
    *out = ctx;
    return AIX_OK;
}

void aix_ctx_destroy(aix_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    aix::mbox_stop(ctx);
    if (ctx->mbox_stream) cudaStreamDestroy(ctx->mbox_stream);
    if (ctx->mbox_host) cudaFreeHost(ctx->mbox_host);
    cudaDeviceSynchronize();
    for (auto &b : ctx->scratch)
        if (b.p) cudaFree(b.p);
    aix_plain_cache_flush(ctx);
    aix_count13_peers_close(ctx);
    if (ctx->small_host) cudaFreeHost(ctx->small_host);
    if (ctx->c13_hist32) cudaFree(ctx->c13_hist32);
    if (ctx->c13_hist64) cudaFree(ctx->c13_hist64);
    if (ctx->c13_stats_dev) cudaFree(ctx->c13_stats_dev);
    if (ctx->c23_kmers_dev) cudaFree(ctx->c23_kmers_dev);
    if (ctx->c23_counts_dev) cudaFree(ctx->c23_counts_dev);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->xfer[0]) cudaStreamDestroy(ctx->xfer[0]);
    if (ctx->xfer[1]) cudaStreamDestroy(ctx->xfer[1]);
    delete ctx;
}

const char *aix_last_error(const aix_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int aix_ctx_device(const aix_ctx *ctx) { return ctx ? ctx->device : -1; }
void *aix_ctx_stream(const aix_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t aix_ctx_launch_count(const aix_ctx *ctx) { return ctx ? ctx->launches : 0; }

int aix_ctx_sync(aix_ctx *ctx) {
    if (!ctx) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[0]));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->xfer[1]));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AIX_OK;
}

int aix_ctx_trim(aix_ctx *ctx) {
    if (!ctx) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    AIX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pool) AIX_CUDA(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
    aix_plain_cache_flush(ctx);
    return AIX_OK;
}

int aix_host_alloc(aix_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return AIX_ERR_ARG;
    AIX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(AIX_ERR_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return AIX_OK;
}

int aix_host_free(aix_ctx *ctx, void *p) {
    if (!p) return AIX_OK;
    if (!ctx) return cudaFreeHost(p) == cudaSuccess ? AIX_OK : AIX_ERR_CUDA;  // ctx may already be gone
    AIX_CUDA(ctx, cudaFreeHost(p));
    return AIX_OK;
}

}  // extern "C"
