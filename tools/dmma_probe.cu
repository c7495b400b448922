// Probe: what limits DMMA issue in a GEMM-like inner loop? (development tool)
// Variants: operand pattern (same regs / MI x NJ outer product), shared-memory fragment loads, warps per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MI, int NJ, bool LDSF>
__global__ void __launch_bounds__(256, 1) probe(double *out, int iters, const double *src) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = src[i];
    __syncthreads();
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[MI], b[NJ];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = src[lane + 32 * i];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = src[lane + 32 * (j + MI)];
    const double *p = sm + (lane >> 2) * 20 + (lane & 3);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (LDSF) {
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = p[i * 160 + (it & 3) * 4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = p[(j + MI) * 160 + (it & 3) * 4];
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MI, int NJ, bool LDSF>
void run(const char *name, int warps_per_sm, int sms, double *out, const double *src) {
    int threads = warps_per_sm * 32;
    int iters = 4000;
    auto k = probe<MI, NJ, LDSF>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 1e30;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        k<<<sms, threads, 4096 * 8>>>(out, iters, src);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    double flops = (double)sms * warps_per_sm * iters * MI * NJ * 512.0;
    printf("%-28s warps/SM=%2d  %.2f TFLOP/s\n", name, warps_per_sm, flops / best / 1e9);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double *out, *src; CK(cudaMalloc(&out, 8 * sms * 1024)); CK(cudaMalloc(&src, 8 * 4096)); CK(cudaMemset(src, 0, 8 * 4096));
    for (int w : {4, 8}) {
        run<4, 8, false>("4x8 regs", w, sms, out, src);
        run<4, 8, true>("4x8 lds", w, sms, out, src);
        run<8, 4, false>("8x4 regs", w, sms, out, src);
        run<2, 4, false>("2x4 regs", w, sms, out, src);
        run<2, 4, true>("2x4 lds", w, sms, out, src);
        run<4, 4, false>("4x4 regs", w, sms, out, src);
        run<4, 4, true>("4x4 lds", w, sms, out, src);
    }
    for (int w : {12, 16}) {
        run<2, 4, false>("2x4 regs", w, sms, out, src);
        run<2, 4, true>("2x4 lds", w, sms, out, src);
        run<4, 4, true>("4x4 lds", w, sms, out, src);
    }
    return 0;
}
