// xr_gemm_scatter: the FP64 contraction engine of libxr_b200.so.
//
//   C[offM(m) + offN(n)] (= | +=) alpha * sum_k A[m*lda+k] * B[n*ldb+k]
//
// Roofline: FP64 tensor pipe (DMMA.8x8x4, 64 FMA/clk/SM = 37.2 TFLOP/s measured on B200) for
// K >= ~64; HBM write bandwidth (8 B per output element) for the K = 2n charge-transfer classes.
// Two operand-staging paths: 16-byte-aligned operands (the normal case) go through tensor-map TMA (xr_gemm_tma.cu,
// 33.1 / 35.0 TFLOP/s at K = 326 / 2304); this file is the cp.async path for everything else (odd leading dimensions,
// unaligned bases) and the reference point the TMA path was validated against (30.7 / 32.6 TFLOP/s).
// Layout: CTA tile BM x BN (64 x 64 as shipped), K streamed in BK=16 slices through a STAGES-deep cp.async ring in
// shared memory; rows are padded to BK+4 doubles so the 8-row x 4-k DMMA fragment reads hit 16
// distinct 8-byte bank slots per half warp (stride = 4 mod 16 doubles).  Each warp owns a
// (BM/WARPS_M) x (BN/WARPS_N) sub-tile as 8x8 DMMA accumulators in registers.  The epilogue adds
// two int64 offset tables, which is how results land directly in the Hamiltonian's final layout.
#include "xr_common.cuh"

int xr_gemm_scatter_tma(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double* C, const int64_t* offM, int64_t ldc, const int64_t* offN,
                        int accumulate);

namespace {

struct GemmParams {
    int64_t M, N, K;
    double alpha;
    const double* A;
    int64_t lda;
    const double* B;
    int64_t ldb;
    double* C;
    const int64_t* offM;
    int64_t ldc;
    const int64_t* offN;
    int accumulate;
    int64_t tiles_n;
};

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, bool VEC16>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, 1) gemm_nt_scatter_kernel(const GemmParams p) {
    constexpr int THREADS = WARPS_M * WARPS_N * 32;
    constexpr int LDS = BK + 4;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int MI = WM / 8, NJ = WN / 8;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + (size_t)STAGES * BM * LDS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
    const int64_t tile = blockIdx.x;
    const int64_t m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
    const int KT = (int)((p.K + BK - 1) / BK);

    auto load_tile = [&](int stage, int kt) {
        const int64_t k0 = (int64_t)kt * BK;
        double* as = As + (size_t)stage * BM * LDS;
        double* bs = Bs + (size_t)stage * BN * LDS;
        if (VEC16) {
            constexpr int CPR = BK / 2;   // 16-byte chunks per row
            for (int c = tid; c < BM * CPR; c += THREADS) {
                int row = c / CPR, kc = c % CPR;
                int64_t k = k0 + 2 * kc, rem = p.K - k;
                int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
                int64_t grow = m0 + row < p.M ? m0 + row : p.M - 1;
                cp_async_16(smem_u32(as + row * LDS + 2 * kc), p.A + grow * p.lda + (rem > 0 ? k : 0), bytes);
            }
            for (int c = tid; c < BN * CPR; c += THREADS) {
                int row = c / CPR, kc = c % CPR;
                int64_t k = k0 + 2 * kc, rem = p.K - k;
                int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
                int64_t grow = n0 + row < p.N ? n0 + row : p.N - 1;
                cp_async_16(smem_u32(bs + row * LDS + 2 * kc), p.B + grow * p.ldb + (rem > 0 ? k : 0), bytes);
            }
        } else {
            for (int c = tid; c < BM * BK; c += THREADS) {
                int row = c / BK, kc = c % BK;
                int64_t k = k0 + kc;
                int64_t grow = m0 + row < p.M ? m0 + row : p.M - 1;
                cp_async_8(smem_u32(as + row * LDS + kc), p.A + grow * p.lda + (k < p.K ? k : 0), k < p.K ? 8 : 0);
            }
            for (int c = tid; c < BN * BK; c += THREADS) {
                int row = c / BK, kc = c % BK;
                int64_t k = k0 + kc;
                int64_t grow = n0 + row < p.N ? n0 + row : p.N - 1;
                cp_async_8(smem_u32(bs + row * LDS + kc), p.B + grow * p.ldb + (k < p.K ? k : 0), k < p.K ? 8 : 0);
            }
        }
    };

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < KT) load_tile((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        cp_async_commit();
        const double* as = As + (size_t)(kt % STAGES) * BM * LDS + (warp_m * WM + g) * LDS + t;
        const double* bs = Bs + (size_t)(kt % STAGES) * BN * LDS + (warp_n * WN + g) * LDS + t;
        // (k-steps fully unrolled: measured faster than a rolled k loop with prefetched fragments, 25.5 vs 24.0
        //  TFLOP/s at K=326 on B200 -- ptxas interleaves the fragment loads with the DMMA stream by itself)
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * LDS + ks * 4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = bs[j * 8 * LDS + ks * 4];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane holds C[8i+g][8j+2t], C[8i+g][8j+2t+1] of every 8x8 block of the warp tile
    int64_t on[NJ][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        int64_t col = n0 + warp_n * WN + j * 8 + 2 * t;
        on[j][0] = col < p.N ? (p.offN ? p.offN[col] : col) : -1;
        on[j][1] = col + 1 < p.N ? (p.offN ? p.offN[col + 1] : col + 1) : -1;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int64_t row = m0 + warp_m * WM + i * 8 + g;
        if (row >= p.M) continue;
        int64_t om = p.offM ? p.offM[row] : row * p.ldc;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            // the lane's two columns are adjacent in C more often than not (runs of ket states): one 16-byte store
            double* dst0 = p.C + om + on[j][0];
            if (on[j][0] >= 0 && on[j][1] == on[j][0] + 1 && (reinterpret_cast<uintptr_t>(dst0) & 15) == 0) {
                double2 v = make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
                if (p.accumulate) {
                    const double2 old = *reinterpret_cast<double2*>(dst0);
                    v.x += old.x;
                    v.y += old.y;
                }
                *reinterpret_cast<double2*>(dst0) = v;
                continue;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (on[j][e] < 0) continue;
                double* dst = p.C + om + on[j][e];
                double v = p.alpha * acc[i][j][e];
                *dst = p.accumulate ? *dst + v : v;
            }
        }
    }
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, bool VEC16>
int launch_gemm(xr_ctx* ctx, GemmParams p) {
    constexpr int THREADS = WARPS_M * WARPS_N * 32;
    constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
    auto kernel = gemm_nt_scatter_kernel<BM, BN, BK, WARPS_M, WARPS_N, STAGES, VEC16>;
    XR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    int64_t tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    int64_t tiles = tiles_m * p.tiles_n;
    XR_REQUIRE(tiles < (1ll << 31), "xr_gemm_scatter: too many tiles (%lld)", (long long)tiles);
    kernel<<<(unsigned)tiles, THREADS, SMEM, ctx->stream>>>(p);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}

}  // namespace

extern "C" int xr_gemm_scatter(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                               const double* B, int64_t ldb, double* C, const int64_t* offM, int64_t ldc,
                               const int64_t* offN, int accumulate) {
    XR_REQUIRE(ctx, "xr_gemm_scatter: null ctx");
    if (M <= 0 || N <= 0) return XR_OK;
    XR_REQUIRE(K >= 0 && A && B && C, "xr_gemm_scatter: null pointer or negative K");
    XR_REQUIRE(lda >= K && ldb >= K, "xr_gemm_scatter: lda/ldb smaller than K");
    XR_REQUIRE(offM || ldc >= 1, "xr_gemm_scatter: need offM or ldc");
    GemmParams p{M, N, K, alpha, A, lda, B, ldb, C, offM, ldc, offN, accumulate, 0};
    const bool vec16 = (lda % 2 == 0) && (ldb % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    // 64x64 CTA tile, 4 warps x (32x32), 2-stage ring: 120 registers and 41 KB of shared memory per CTA, so 4 CTAs
    // (16 warps) share an SM and one CTA's barriers / epilogue hide behind the others' DMMA streams.  Measured on
    // B200 against 128x128 (1 CTA/SM) and 128x64 (2 CTAs/SM) tiles: 30.7 vs 26.1 / 28.2 TFLOP/s at K=326,
    // 32.6 vs 29.9 / 32.6 at K=2304, 1.97 vs 1.48 / 1.57 TB/s of C written at K=36.
    if (vec16) {
        // tensor-map TMA staging (xr_gemm_tma.cu); falls through to cp.async staging if no tensor map can describe the operands
        int rc = xr_gemm_scatter_tma(ctx, M, N, K, alpha, A, lda, B, ldb, C, offM, ldc, offN, accumulate);
        if (rc != XR_ERR_UNSUPPORTED) return rc;
    }
    return vec16 ? launch_gemm<64, 64, 16, 2, 2, 2, true>(ctx, p) : launch_gemm<64, 64, 16, 2, 2, 2, false>(ctx, p);
}
