// xr_probe_fp64: the roofline denominator of the tensor-bound kernels, measured in the same process and on the same
// device as the run that quotes it.  sm_100a has no tcgen05 f64 kind; its FP64 tensor instruction is DMMA.8x8x4
// (mma.sync.m8n8k4.f64), which shares one pipe with DFMA.  The probe issues independent DMMA chains from 16 warps per
// SM for about `seconds` and reports the sustained rate (tools/fp64_peaks.cu is the stand-alone sweep this came from:
// 37.2 TFLOP/s = 148 SMs x 64 FMA/clk x 2 x 1.965 GHz on this pool's B200s; cuBLAS DGEMM 8192^3 reaches 35.5).
#include "xr_common.cuh"

namespace {

template <int NACC>
__global__ void __launch_bounds__(128) probe_dmma_kernel(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        c0[i] = threadIdx.x;
        c1[i] = i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma_m8n8k4(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int xr_probe_fp64(xr_ctx* ctx, double seconds, double* dmma_tflops) {
    XR_REQUIRE(ctx && dmma_tflops, "xr_probe_fp64: null argument");
    XR_REQUIRE(seconds > 0.0 && seconds <= 10.0, "xr_probe_fp64: seconds must be in (0, 10]");
    constexpr int NACC = 8, THREADS = 128, BLOCKS_PER_SM = 4, ITERS = 40000;
    const int grid = ctx->sm_count * BLOCKS_PER_SM;
    const size_t out_bytes = (size_t)grid * THREADS * sizeof(double);
    int rc = xr_ensure_scratch(ctx, out_bytes);
    if (rc != XR_OK) return rc;
    double* out = static_cast<double*>(ctx->scratch);
    cudaEvent_t e0, e1;
    XR_CUDA(cudaEventCreate(&e0));
    XR_CUDA(cudaEventCreate(&e1));
    const double flops_per_launch = (double)grid * (THREADS / 32) * ITERS * NACC * 2.0 * 8 * 8 * 4;
    // warm-up launch (clocks), then as many launches as fit `seconds` at ~37 TFLOP/s, timed as one region
    probe_dmma_kernel<NACC><<<grid, THREADS, 0, ctx->stream>>>(out, ITERS, 1.0000001, 0.9999999);
    XR_CUDA(cudaGetLastError());
    int launches = (int)(seconds * 37.0e12 / flops_per_launch);
    if (launches < 1) launches = 1;
    XR_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int l = 0; l < launches; ++l) probe_dmma_kernel<NACC><<<grid, THREADS, 0, ctx->stream>>>(out, ITERS, 1.0000001, 0.9999999);
    XR_CUDA(cudaGetLastError());
    XR_CUDA(cudaEventRecord(e1, ctx->stream));
    XR_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    XR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->launches += launches + 1;
    *dmma_tflops = flops_per_launch * launches / (ms * 1e-3) / 1e12;
    return XR_OK;
}
