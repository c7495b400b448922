// Double-double (~106-bit) matrix product for the Newton refinement of the overlap inverse:
// hermitian-XRCC/get_xr_result.py:165,202,246,285 call qode.math.precise_numpy_inverse(S2), an inverse polished in extended
// precision; on the GPU the polish X <- X + X (I - M X) runs with every product split exactly by FMA (two-product) and every
// addition compensated (two-sum), which is more accurate than the x87 long double the host version relies on.
// Latency-sized work (dim(S2)^3 ~ 1e8 products): a plain 16 x 16 shared-memory tiling is enough.
#include "xr_common.cuh"

namespace {

constexpr int DT = 16;

// Error-free transformations.  Written with the _rn intrinsics, which the compiler may not contract: with plain operators
// nvcc's default -fmad=true can fuse `p = a*b; s = hi + p` into fma(a, b, hi), after which s is no longer fl(hi + p) and
// the compensation term is wrong.
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    const double bb = __dadd_rn(s, -a);
    e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
}

// out[i,j] = C0[i,j] (or delta_ij when C0 == nullptr) + sign * sum_k A[i,k] B[k,j]
__global__ void __launch_bounds__(DT * DT)
gemm_dd_kernel(int64_t M, int64_t N, int64_t K, const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
               const double* __restrict__ C0, int64_t ldc0, double sign, double* __restrict__ out, int64_t ldo) {
    __shared__ double As[DT][DT + 1], Bs[DT][DT + 1];
    const int tx = threadIdx.x % DT, ty = threadIdx.x / DT;
    const int64_t i = blockIdx.y * (int64_t)DT + ty, j = blockIdx.x * (int64_t)DT + tx;
    double hi = 0.0, lo = 0.0;
    for (int64_t k0 = 0; k0 < K; k0 += DT) {
        As[ty][tx] = (i < M && k0 + tx < K) ? A[i * lda + k0 + tx] : 0.0;
        Bs[ty][tx] = (k0 + ty < K && j < N) ? B[(k0 + ty) * ldb + j] : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const double a = As[ty][k], b = Bs[k][tx];
            const double p = __dmul_rn(a, b);
            const double pe = __fma_rn(a, b, -p);     // a*b == p + pe exactly
            double s, se;
            two_sum(hi, p, s, se);
            hi = s;
            lo = __dadd_rn(lo, __dadd_rn(se, pe));
        }
        __syncthreads();
    }
    if (i < M && j < N) {
        const double c = C0 ? C0[i * ldc0 + j] : (i == j ? 1.0 : 0.0);
        double s, se;
        two_sum(c, sign * hi, s, se);          // sign = +-1: exact
        out[i * ldo + j] = __dadd_rn(s, __dadd_rn(se, sign * lo));
    }
}

}  // namespace

extern "C" int xr_gemm_dd(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B, int64_t ldb,
                          const double* C0, int64_t ldc0, double sign, double* out, int64_t ldo) {
    XR_REQUIRE(ctx, "xr_gemm_dd: null ctx");
    if (M <= 0 || N <= 0) return XR_OK;
    XR_REQUIRE(A && B && out && K >= 0, "xr_gemm_dd: null pointer or negative K");
    XR_REQUIRE(lda >= K && ldb >= N && ldo >= N && (!C0 || ldc0 >= N), "xr_gemm_dd: leading dimension too small");
    XR_REQUIRE(sign == 1.0 || sign == -1.0, "xr_gemm_dd: sign must be +1 or -1");
    XR_REQUIRE(out != A && out != B, "xr_gemm_dd: out must not alias A or B");
    dim3 grid((unsigned)((N + DT - 1) / DT), (unsigned)((M + DT - 1) / DT));
    XR_REQUIRE((M + DT - 1) / DT < 65536, "xr_gemm_dd: too many rows (%lld)", (long long)M);
    gemm_dd_kernel<<<grid, DT * DT, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C0, ldc0, sign, out, ldo);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}
