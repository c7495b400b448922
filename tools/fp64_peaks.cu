// FP64 pipe microbenchmark for B200 (sm_100a): measures the achievable FP64 rate of
//   (1) DFMA (vector pipe), (2) DMMA.8x8x4 (mma.sync f64, the only FP64 tensor shape sm_100a
//   has natively: m16n8k{4,8,16}.f64 all lower to DMMA.8x8x4), (3) both interleaved,
// for several warps/SM, so that the roofline denominator for the XR kernels is a measured
// number (MEASURED_PEAKS.json only carries HBM and bf16).  Prints one JSON object.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a, double b) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NM DMMAs + NF DFMAs per inner step, independent chains
template <int NM, int NF>
__global__ void k_mixed(double *out, int iters, double a, double b) {
    double c0[NM], c1[NM], f[NF];
#pragma unroll
    for (int i = 0; i < NM; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    // warps per SM sweep: blocks of 128 threads (4 warps), k blocks per SM
    int blocks_per_sm[] = {1, 2, 4, 8};
    for (int bi = 0; bi < 4; ++bi) {
        int bps = blocks_per_sm[bi];
        int grid = sms * bps, threads = 128;
        double warps = (double)grid * threads / 32;
        {
            double ms = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, iters, 1.0000001, 0.9999999); }, 5);
            double flops = warps * iters * 8.0 * 2 * 8 * 8 * 4;
            printf(", \"dmma_tflops_w%d\": %.2f", bps * 4, flops / ms / 1e9);
        }
        {
            double ms = time_ms([&] { k_dfma<16><<<grid, threads>>>(out, iters, 1.0000001, 0.9999999); }, 5);
            double flops = warps * 32 * iters * 16.0 * 2;
            printf(", \"dfma_tflops_w%d\": %.2f", bps * 4, flops / ms / 1e9);
        }
    }
    {   // mixed at 16 warps/SM: 8 DMMA (2048 FMA/warp) + 16 DFMA (512 FMA/warp) per step
        int grid = sms * 4, threads = 128;
        double warps = (double)grid * threads / 32;
        double ms = time_ms([&] { k_mixed<8, 16><<<grid, threads>>>(out, iters, 1.0000001, 0.9999999); }, 5);
        double f_mma = warps * iters * 8.0 * 2 * 256, f_fma = warps * 32 * iters * 16.0 * 2;
        printf(", \"mixed_8dmma_16dfma_tflops\": %.2f, \"mixed_dmma_part\": %.2f, \"mixed_dfma_part\": %.2f",
               (f_mma + f_fma) / ms / 1e9, f_mma / ms / 1e9, f_fma / ms / 1e9);
        ms = time_ms([&] { k_mixed<8, 64><<<grid, threads>>>(out, iters / 4, 1.0000001, 0.9999999); }, 5);
        f_mma = warps * (iters / 4) * 8.0 * 2 * 256; f_fma = warps * 32 * (iters / 4) * 64.0 * 2;
        printf(", \"mixed_8dmma_64dfma_tflops\": %.2f", (f_mma + f_fma) / ms / 1e9);
    }
    {   // sustained DMMA (about 3 s) to see the power-capped rate
        int grid = sms * 4, threads = 128;
        double warps = (double)grid * threads / 32;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 0;
        for (; launches < 40; ++launches) k_dmma<8><<<grid, threads>>>(out, iters * 4, 1.0000001, 0.9999999);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = warps * (double)iters * 4 * launches * 8.0 * 2 * 256;
        printf(", \"dmma_tflops_sustained\": %.2f, \"sustained_seconds\": %.2f", flops / ms / 1e9, ms / 1e3);
    }
    printf("}\n");
    return 0;
}
