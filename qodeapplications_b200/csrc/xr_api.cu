// Context, memory and the small data-movement kernels of libxr_b200.so.
#include "xr_common.cuh"
#include <cstring>
#include <cmath>

static thread_local char g_error[512] = "";

void xr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* xr_last_error(void) { return g_error; }
extern "C" const char* xr_version(void) { return "xr_b200 0.1 sm_100a"; }

extern "C" int xr_ctx_create(int device, void* stream, int own_stream, xr_ctx** out) {
    XR_REQUIRE(out != nullptr, "xr_ctx_create: out is null");
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        xr_set_error("xr_ctx_create: no CUDA device (%s); libxr_b200 has no CPU fallback",
                     err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
        return XR_ERR_NO_DEVICE;
    }
    XR_REQUIRE(device >= 0 && device < count, "xr_ctx_create: device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    XR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        xr_set_error("xr_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                     prop.minor);
        return XR_ERR_NO_DEVICE;
    }
    XR_CUDA(cudaSetDevice(device));
    xr_ctx* ctx = new xr_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (!own_stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
        ctx->owns_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            xr_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
            return XR_ERR_CUDA;
        }
        ctx->owns_stream = true;
    }
    *out = ctx;
    return XR_OK;
}

extern "C" int xr_ctx_destroy(xr_ctx* ctx) {
    if (!ctx) return XR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return XR_OK;
}

extern "C" int xr_ctx_set_stream(xr_ctx* ctx, void* stream) {
    XR_REQUIRE(ctx, "xr_ctx_set_stream: null ctx");
    if (ctx->owns_stream) {
        XR_CUDA(cudaStreamSynchronize(ctx->stream));
        XR_CUDA(cudaStreamDestroy(ctx->stream));
        ctx->owns_stream = false;
    }
    ctx->stream = static_cast<cudaStream_t>(stream);
    return XR_OK;
}

extern "C" int xr_sync(xr_ctx* ctx) {
    XR_REQUIRE(ctx, "xr_sync: null ctx");
    XR_CUDA(cudaStreamSynchronize(ctx->stream));
    return XR_OK;
}

extern "C" int xr_launch_count(xr_ctx* ctx, int64_t* count) {
    XR_REQUIRE(ctx && count, "xr_launch_count: null argument");
    *count = ctx->launches;
    return XR_OK;
}

extern "C" int xr_device_info(xr_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes,
                              size_t* total_bytes) {
    XR_REQUIRE(ctx, "xr_device_info: null ctx");
    cudaDeviceProp prop;
    XR_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    size_t f = 0, t = 0;
    XR_CUDA(cudaSetDevice(ctx->device));
    XR_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return XR_OK;
}

extern "C" int xr_malloc(xr_ctx* ctx, size_t bytes, void** dptr) {
    XR_REQUIRE(ctx && dptr, "xr_malloc: null argument");
    XR_CUDA(cudaSetDevice(ctx->device));
    XR_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return XR_OK;
}

extern "C" int xr_free(xr_ctx* ctx, void* dptr) {
    XR_REQUIRE(ctx, "xr_free: null ctx");
    XR_CUDA(cudaSetDevice(ctx->device));
    XR_CUDA(cudaStreamSynchronize(ctx->stream));
    XR_CUDA(cudaFree(dptr));
    return XR_OK;
}

extern "C" int xr_memset_zero(xr_ctx* ctx, void* dptr, size_t bytes) {
    XR_REQUIRE(ctx && (dptr || bytes == 0), "xr_memset_zero: null argument");
    XR_CUDA(cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
    return XR_OK;
}

extern "C" int xr_upload(xr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    XR_REQUIRE(ctx && (bytes == 0 || (dst && src)), "xr_upload: null argument");
    XR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return XR_OK;
}

extern "C" int xr_download(xr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    XR_REQUIRE(ctx && (bytes == 0 || (dst && src)), "xr_download: null argument");
    XR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return XR_OK;
}

int xr_ensure_scratch(xr_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return XR_OK;
    XR_CUDA(cudaSetDevice(ctx->device));
    XR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) XR_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    XR_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return XR_OK;
}

// ------------------------------------------------------------------ small data-movement kernels

__global__ void copy2d_scaled_kernel(double* __restrict__ dst, int64_t dst_ld, const double* __restrict__ src,
                                     int64_t src_ld, int64_t rows, int64_t cols, double alpha) {
    int64_t total = rows * cols;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / cols, c = t - r * cols;
        dst[r * dst_ld + c] = alpha * src[r * src_ld + c];
    }
}

extern "C" int xr_copy2d_scaled(xr_ctx* ctx, double* dst, int64_t dst_ld, const double* src, int64_t src_ld, int64_t rows,
                                int64_t cols, double alpha) {
    XR_REQUIRE(ctx, "xr_copy2d_scaled: null ctx");
    if (rows <= 0 || cols <= 0) return XR_OK;
    XR_REQUIRE(dst && src && dst_ld >= cols && src_ld >= cols, "xr_copy2d_scaled: bad arguments");
    int64_t total = rows * cols;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    copy2d_scaled_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(dst, dst_ld, src, src_ld, rows, cols, alpha);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}

__global__ void scatter_const_kernel(double* __restrict__ C, const int64_t* __restrict__ idx, int64_t count, double value,
                                     int accumulate) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t at = idx[t];
        C[at] = accumulate ? C[at] + value : value;
    }
}

extern "C" int xr_scatter_const(xr_ctx* ctx, double* C, const int64_t* idx, int64_t count, double value, int accumulate) {
    XR_REQUIRE(ctx, "xr_scatter_const: null ctx");
    if (count <= 0) return XR_OK;
    XR_REQUIRE(C && idx, "xr_scatter_const: null pointer");
    int64_t blocks = (count + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    scatter_const_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(C, idx, count, value, accumulate);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}

struct PermuteParams {
    int nd;
    int64_t shape[12];
    int64_t stride[12];
    int64_t total;
};

__global__ void permute_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, const PermuteParams p, double alpha) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < p.total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = t, at = 0;
#pragma unroll
        for (int d = 11; d >= 0; --d) {
            if (d < p.nd) {
                int64_t i = rem % p.shape[d];
                rem /= p.shape[d];
                at += i * p.stride[d];
            }
        }
        dst[t] = alpha * src[at];
    }
}

extern "C" int xr_permute_copy(xr_ctx* ctx, double* dst, const double* src, int nd, const int64_t* shape,
                               const int64_t* src_strides, double alpha) {
    XR_REQUIRE(ctx && shape && src_strides, "xr_permute_copy: null argument");
    XR_REQUIRE(nd >= 1 && nd <= 12, "xr_permute_copy: nd=%d unsupported (1..12)", nd);
    PermuteParams p{};
    p.nd = nd;
    p.total = 1;
    for (int d = 0; d < nd; ++d) {
        XR_REQUIRE(shape[d] >= 0, "xr_permute_copy: negative extent");
        p.shape[d] = shape[d];
        p.stride[d] = src_strides[d];
        p.total *= shape[d];
    }
    if (p.total == 0) return XR_OK;      // an empty charge sector: nothing to copy, and the buffers may be null
    XR_REQUIRE(dst && src, "xr_permute_copy: null buffer");
    int64_t blocks = (p.total + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    permute_copy_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(dst, src, p, alpha);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}
