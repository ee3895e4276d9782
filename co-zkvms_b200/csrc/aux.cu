// Auxiliary entry points of libcozk_msm.so: plain device-memory helpers, so that callers (and the tests / bench) need no
// other CUDA binding.  Synthetic inputs, test kernels and microbenchmarks live in their own library (testlib.cu).
#include "engine.hpp"

namespace cozk {

static int get_device(cozk_ctx* ctx, int device_index, Device** out) {
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) {
        set_error("bad context or device index");
        return COZK_ERR_INVALID_ARG;
    }
    *out = ctx->devs[device_index].get();
    COZK_CUDA(cudaSetDevice((*out)->id));
    return COZK_OK;
}

}  // namespace cozk

using namespace cozk;

extern "C" {

int cozk_dev_alloc(cozk_ctx* ctx, int device_index, size_t bytes, void** out) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    if (!out) return COZK_ERR_INVALID_ARG;
    COZK_CUDA(cudaMalloc(out, bytes ? bytes : 16));
    return COZK_OK;
}
int cozk_dev_free(cozk_ctx* ctx, int device_index, void* ptr) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    COZK_CUDA(cudaFree(ptr));
    return COZK_OK;
}
int cozk_dev_upload(cozk_ctx* ctx, int device_index, void* dst, const void* src, size_t bytes) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    // on the engine's stream, then synchronised: a pageable cudaMemcpy only guarantees that the data has been staged, and
    // its DMA (legacy default stream) is not ordered before kernels on the engine's non-blocking streams
    std::lock_guard<std::mutex> lock(D->mu);
    COZK_CUDA(cudaSetDevice(D->id));
    COZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, D->stream));
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}
int cozk_dev_download(cozk_ctx* ctx, int device_index, void* dst, const void* src, size_t bytes) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(D->mu);
    COZK_CUDA(cudaSetDevice(D->id));
    COZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, D->stream));
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}
int cozk_host_alloc_pinned(size_t bytes, void** out) {
    if (!out) return COZK_ERR_INVALID_ARG;
    COZK_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
    return COZK_OK;
}
int cozk_host_free_pinned(void* ptr) {
    COZK_CUDA(cudaFreeHost(ptr));
    return COZK_OK;
}
int cozk_dev_flush_l2(cozk_ctx* ctx, int device_index) {
    Device* D;
    int rc = get_device(ctx, device_index, &D);
    if (rc) return rc;
    const size_t bytes = (size_t)256 << 20;
    if ((rc = D->flush.ensure(bytes))) return rc;
    COZK_CUDA(cudaMemsetAsync(D->flush.p, 0x5A, bytes, D->stream));
    COZK_CUDA(cudaStreamSynchronize(D->stream));
    return COZK_OK;
}

}  // extern "C"
