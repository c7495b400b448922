// xr_gemm_scatter, tensor-map TMA variant (used for 16-byte-aligned operands).
//
// Same math and epilogue as gemm_nt_scatter_kernel (xr_gemm.cu); the difference is how the A[BM x BK] and
// B[BN x BK] operand tiles reach shared memory: ONE elected thread issues two cp.async.bulk.tensor.2d loads
// (SASS UTMALDG) per k-tile against tensor maps that describe A as [M rows][K] and B as [N rows][K]; the
// copies complete on an mbarrier with a transaction count.  TMA zero-fills everything outside the tensor, so
// the K tail (K = 326 = 20*16 + 6, K = 36, ...) and the M/N tails need no per-thread predication or source
// clamping at all.  Tiles are 64 rows x 16 doubles = 64 x 128 bytes and are stored with the 128-byte swizzle:
// the 16-byte chunk c of row r sits at chunk (c ^ (r & 7)), which keeps the 8-row x 4-k DMMA fragment reads to
// at most 2-way bank conflicts without padding bytes.
#include "xr_common.cuh"
#include <cuda.h>

namespace {

struct GemmTmaParams {
    int64_t M, N, K;
    double alpha;
    double* C;
    const int64_t* offM;
    int64_t ldc;
    const int64_t* offN;
    int accumulate;
    int64_t tiles_m, tiles_n;
    double* partials;      // REDUCE mode: [tiles][2] per-CTA (sum, sum of squares); nothing is stored to C
    int kt_per_split;      // SPLITK mode: k-tiles per blockIdx.y slice; raw partial tiles go to ws[split][M][N]
    double* ws;
    unsigned long long* tile_counter;      // persistent kernel, dynamic scheduling: next tile to hand out (zeroed per launch)
};

enum { MODE_SCATTER = 0, MODE_REDUCE = 1, MODE_SPLITK = 2 };

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// XR_GEMM_VARIANT (build-time only; tools/gemm_variants.py builds one library per value and times them in one GPU call):
//   0  one output tile per CTA (gemm_tma_scatter_kernel<MODE_SCATTER>)
//   1  persistent CTAs, static tile striding, thread 0 issues every TMA load
//   2  persistent, static, the issuing duty rotates over the four warps (one sub-core is not always the late one)
//   3  persistent, tiles handed out by an atomic counter (thread 0 issues), so faster SMs take more tiles
#ifndef XR_GEMM_VARIANT
#define XR_GEMM_VARIANT 3
#endif
// ... and the K range the persistent kernel is used for (k-tiles of 16): below/above, the one-tile kernel
#ifndef XR_GEMM_PERSISTENT_MAX_KT
#define XR_GEMM_PERSISTENT_MAX_KT 4
#endif

// ... whether K <= 48 goes to the whole-K kernel (one stage per tile) instead:
#ifndef XR_GEMM_WHOLEK
#define XR_GEMM_WHOLEK 1
#endif
// ... and the ring depth of the persistent kernel: deeper rings prefetch further across tile boundaries but fewer CTAs fit an SM
#ifndef XR_GEMM_PERSISTENT_STAGES
#define XR_GEMM_PERSISTENT_STAGES 3
#endif

constexpr int TBM = 64, TBN = 64, TBK = 16, TSTAGES = 3, TTHREADS = 128;
// ... and its warp count: 2 x WARPS_N warps on the 64 x 64 tile (2: 32 x 32 warp tiles, 64 accumulator registers; 4: 32 x 16 warp
// tiles, half the registers per thread, twice the resident warps)
#ifndef XR_GEMM_PERSISTENT_WARPS_N
#define XR_GEMM_PERSISTENT_WARPS_N 2
#endif
constexpr int PWN = XR_GEMM_PERSISTENT_WARPS_N, PTHREADS = 2 * PWN * 32;
constexpr int PSTAGES = XR_GEMM_PERSISTENT_STAGES;
constexpr size_t PERSISTENT_SMEM = (size_t)PSTAGES * (TBM + TBN) * TBK * 8 + 1024;
constexpr int PCTAS = (int)((227 * 1024) / (PERSISTENT_SMEM + 1024 + 128)) < 4 ? (int)((227 * 1024) / (PERSISTENT_SMEM + 1024 + 128)) : 4;
constexpr int TILE_A_BYTES = TBM * TBK * 8, TILE_B_BYTES = TBN * TBK * 8;     // 8 KB each, 1024-byte aligned
constexpr int GROUP_M = 16;
constexpr size_t TMA_SMEM = (size_t)TSTAGES * (TILE_A_BYTES + TILE_B_BYTES) + 1024;

// byte offset of element (row r, k) inside a 128-byte-swizzled [rows][16 doubles] tile
__device__ __forceinline__ int swz(int r, int k) { return r * 128 + ((((k >> 1) ^ (r & 7)) << 4) | ((k & 1) << 3)); }

template <int MODE>
__global__ void __launch_bounds__(TTHREADS, 4)
gemm_tma_scatter_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTmaParams p) {
    constexpr int MI = 4, NJ = 4;      // 2 x 2 warps, each 32 x 32
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[TSTAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp & 1, warp_n = warp >> 1;
    // Grouped rasterisation: consecutive CTAs walk GROUP_M row tiles before moving to the next column tile, so one resident
    // wave (592 CTAs) touches ~16 + 37 tile strips instead of 1 + 592 and both operands stay in L2 for large N.
    const int64_t tile = blockIdx.x;
    const int64_t per_group = GROUP_M * p.tiles_n;
    const int64_t first_m = (tile / per_group) * GROUP_M;
    const int64_t group_m = p.tiles_m - first_m < GROUP_M ? p.tiles_m - first_m : GROUP_M;
    const int64_t m0 = (first_m + (tile % per_group) % group_m) * TBM, n0 = ((tile % per_group) / group_m) * TBN;
    int KT = (int)((p.K + TBK - 1) / TBK), kt_begin = 0;
    if (MODE == MODE_SPLITK) {      // this CTA owns k-tiles [kt_begin, kt_begin + KT) of its output tile
        kt_begin = (int)blockIdx.y * p.kt_per_split;
        KT = KT - kt_begin < p.kt_per_split ? KT - kt_begin : p.kt_per_split;
    }

    const int64_t k_last = (int64_t)(kt_begin + KT - 1) * TBK;            // first k of this CTA's last k-tile
    const int last_ksteps = p.K - k_last >= TBK ? TBK / 4 : (int)((p.K - k_last + 3) / 4);

    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int kt) {      // one thread: arm the barrier with the byte count, then the two tensor copies
        const int s = kt % TSTAGES;
        unsigned char* a = smem + (size_t)s * (TILE_A_BYTES + TILE_B_BYTES);
        mbar_expect_tx(&full[s], TILE_A_BYTES + TILE_B_BYTES);
        tma_load_2d(a, &mapA, (kt_begin + kt) * TBK, (int)m0, &full[s]);
        tma_load_2d(a + TILE_A_BYTES, &mapB, (kt_begin + kt) * TBK, (int)n0, &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < TSTAGES - 1 && s < KT; ++s) issue(s);
    }

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % TSTAGES;
        mbar_wait(&full[s], (uint32_t)(kt / TSTAGES) & 1);
        __syncthreads();                                   // every warp has finished reading stage (kt-1) % TSTAGES
        if (tid == 0 && kt + TSTAGES - 1 < KT) issue(kt + TSTAGES - 1);
        const unsigned char* as = smem + (size_t)s * (TILE_A_BYTES + TILE_B_BYTES);
        const unsigned char* bs = as + TILE_A_BYTES;
        // the last k-tile may hold fewer than 16 real k (K = 2n = 36 -> 32 + 4): skip the DMMA steps that would only
        // multiply the zeros the TMA unit filled in (a quarter of all tensor work for the d=+-1 dimer classes)
        const int ksteps = kt == KT - 1 ? last_ksteps : TBK / 4;
#pragma unroll
        for (int ks = 0; ks < TBK / 4; ++ks) {
            if (ks >= ksteps) break;
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double*>(as + swz(warp_m * 32 + i * 8 + g, ks * 4 + t));
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double*>(bs + swz(warp_n * 32 + j * 8 + g, ks * 4 + t));
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }

    if (MODE == MODE_SPLITK) {
        double* ws = p.ws + (size_t)blockIdx.y * (size_t)p.M * (size_t)p.N;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int64_t row = m0 + warp_m * 32 + i * 8 + g;
            if (row >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int64_t col = n0 + warp_n * 32 + j * 8 + 2 * t;
                if (col < p.N) ws[row * p.N + col] = acc[i][j][0];
                if (col + 1 < p.N) ws[row * p.N + col + 1] = acc[i][j][1];
            }
        }
        return;
    }

    if (MODE == MODE_REDUCE) {
        // streaming consumer: the tile never leaves the SM.  Rows/columns beyond M/N were zero-filled by the TMA unit and add 0.
        __shared__ double red[2][TTHREADS / 32];
        double s1p[NJ], s2p[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) s1p[j] = s2p[j] = 0.0;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                s1p[j] += acc[i][j][0] + acc[i][j][1];
                s2p[j] = fma(acc[i][j][0], acc[i][j][0], s2p[j]);
                s2p[j] = fma(acc[i][j][1], acc[i][j][1], s2p[j]);
            }
        double s1 = (s1p[0] + s1p[1]) + (s1p[2] + s1p[3]), s2 = (s2p[0] + s2p[1]) + (s2p[2] + s2p[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            red[0][warp] = s1;
            red[1][warp] = s2;
        }
        __syncthreads();
        if (tid == 0) {
            p.partials[2 * tile] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
            p.partials[2 * tile + 1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
        }
        return;
    }

    // epilogue (identical to gemm_nt_scatter_kernel): lane holds C[8i+g][8j+2t], C[8i+g][8j+2t+1]
    int64_t on[NJ][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        int64_t col = n0 + warp_n * 32 + j * 8 + 2 * t;
        on[j][0] = col < p.N ? (p.offN ? p.offN[col] : col) : -1;
        on[j][1] = col + 1 < p.N ? (p.offN ? p.offN[col + 1] : col + 1) : -1;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int64_t row = m0 + warp_m * 32 + i * 8 + g;
        if (row >= p.M) continue;
        int64_t om = p.offM ? p.offM[row] : row * p.ldc;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
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

// Persistent variant of the scatter mode.  A CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... and its TMA ring runs
// ACROSS tile boundaries: thread 0 keeps PSTAGES-1 k-tiles in flight in the CTA's global k-tile sequence, so the first
// operand tiles of tile i+1 are already landing while the warps scatter tile i, and a stage is recycled through an `empty`
// mbarrier (one arrival per warp) instead of a block-wide barrier per k-tile.  For the short-K charge-transfer classes
// (d = +-1: K = 2n = 36 = three k-tiles, the whole K resident in the ring) a one-tile CTA spent more time being launched,
// fetching its tensor maps and filling its pipeline than on its 1.2 us of DMMAs.
__global__ void __launch_bounds__(PTHREADS, PCTAS)
gemm_tma_persistent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTmaParams p) {
    constexpr int MI = 4, NJ = 8 / PWN, WCOLS = NJ * 8;      // 2 x PWN warps, each 32 x WCOLS
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[PSTAGES], empty[PSTAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp & 1, warp_n = warp >> 1;
    const int KT = (int)((p.K + TBK - 1) / TBK);
    const int last_ksteps = (int)((p.K - (int64_t)(KT - 1) * TBK + 3) / 4);       // real k-steps of the last k-tile (1..4)
    const int64_t n_tiles = p.tiles_m * p.tiles_n;
    const int64_t per_group = GROUP_M * p.tiles_n;
    auto tile_origin = [&](int64_t tile, int64_t& m0, int64_t& n0) {      // grouped rasterisation, as in the one-tile kernel
        const int64_t first_m = (tile / per_group) * GROUP_M;
        const int64_t group_m = p.tiles_m - first_m < GROUP_M ? p.tiles_m - first_m : GROUP_M;
        m0 = (first_m + (tile % per_group) % group_m) * TBM;
        n0 = ((tile % per_group) / group_m) * TBN;
    };

    if (tid == 0) {
        for (int s = 0; s < PSTAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], PTHREADS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    constexpr bool ROTATE = XR_GEMM_VARIANT == 2;
    const bool DYNAMIC = XR_GEMM_VARIANT == 3 && p.tile_counter != nullptr;     // (no counter: every CTA has exactly one tile)
    __shared__ int64_t tile_ring[4];       // DYNAMIC: tile ids in hand-out order (-1 = no more), written by the issuer

    // producer state: the next k-tile to issue in this CTA's sequence.  Kept by EVERY thread (it is a function of the
    // step count alone), so that the issuing duty can move between warps; only the elected thread touches barriers / TMA.
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_q = my_tiles * KT;
    int64_t pq = 0, ptile = blockIdx.x, pm0 = 0, pn0 = 0, pseq = 0;
    int pkt = 0;
    bool drained = false;                  // DYNAMIC: the counter ran past the last tile (issuer's knowledge)
    int64_t upcoming = 0;                  // DYNAMIC: tile id fetched ahead of its use (thread 0)
    if (DYNAMIC && tid == 0) upcoming = (int64_t)atomicAdd(p.tile_counter, 1ull);
    auto issue_next = [&]() {
        if (DYNAMIC ? drained : pq >= total_q) return;
        const bool elected = ROTATE ? tid == 32 * (int)(pq % (PTHREADS / 32)) : tid == 0;
        const int s = (int)(pq % PSTAGES);
        if (elected) {
            mbar_wait(&empty[s], (uint32_t)((pq / PSTAGES) & 1) ^ 1);       // passes at once on the first lap
            if (DYNAMIC && pkt == 0) {
                // the id for THIS tile was requested one tile ago; the request for the next one goes out now, so the
                // round trip of the atomic (the issuing thread's long-scoreboard stall in profiles/r02c) is off the path
                ptile = upcoming;
                upcoming = (int64_t)atomicAdd(p.tile_counter, 1ull);
                if (ptile >= n_tiles) ptile = -1;
                tile_ring[pseq & 3] = ptile;                                // published by the barrier arrival below
            }
        }
        if (DYNAMIC && ptile < 0) {        // (only the issuer runs the dynamic producer: ptile is its own)
            if (elected) mbar_arrive_cta(&full[s]);                         // wake the consumers: they read the -1 and leave
            drained = true;
            return;
        }
        if (pkt == 0) tile_origin(ptile, pm0, pn0);
        if (elected) {
            unsigned char* a = smem + (size_t)s * (TILE_A_BYTES + TILE_B_BYTES);
            mbar_expect_tx(&full[s], TILE_A_BYTES + TILE_B_BYTES);
            tma_load_2d(a, &mapA, pkt * TBK, (int)pm0, &full[s]);
            tma_load_2d(a + TILE_A_BYTES, &mapB, pkt * TBK, (int)pn0, &full[s]);
        }
        ++pq;
        if (++pkt == KT) {
            pkt = 0;
            ptile += gridDim.x;
            ++pseq;
        }
    };
    if (ROTATE || tid == 0) {
        for (int s = 0; s < PSTAGES - 1; ++s) issue_next();
    }

    int64_t q = 0;
    for (int64_t tile = blockIdx.x, seq = 0; DYNAMIC || tile < n_tiles; tile += gridDim.x, ++seq) {
        if (DYNAMIC) {      // the tile id travels with the first k-tile of the tile: wait for it, then read the ring
            mbar_wait(&full[q % PSTAGES], (uint32_t)(q / PSTAGES) & 1);
            tile = tile_ring[seq & 3];
            if (tile < 0) break;
        }
        int64_t m0, n0;
        tile_origin(tile, m0, n0);
        double acc[MI][NJ][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // The epilogue's offset-table entries are pulled into L1 NOW, so that their L2 round trip runs under the tile's DMMAs:
        // with three k-tiles per tile those loads were a quarter of all stall samples when first touched after the k loop
        // (profiles/r02c).  Prefetches, not register loads: the kernel has no registers to spare at 4 CTAs per SM.
        if (p.offN) {
            const int64_t col = n0 + warp_n * WCOLS + 2 * t + 8 * (g % NJ);       // the lane's columns are 8j + 2t: j <-> g % NJ
            if (col < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.offN + col));
        }
        if (p.offM) {
            const int64_t row = m0 + warp_m * 32 + lane;
            if (row < p.M) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.offM + row));
        }

        for (int kt = 0; kt < KT; ++kt, ++q) {
            const int s = (int)(q % PSTAGES);
            if (ROTATE || tid == 0) issue_next();                // k-tile q + PSTAGES - 1 into the stage k-tile q - 1 used
            mbar_wait(&full[s], (uint32_t)(q / PSTAGES) & 1);
            const unsigned char* as = smem + (size_t)s * (TILE_A_BYTES + TILE_B_BYTES);
            const unsigned char* bs = as + TILE_A_BYTES;
            const int ksteps = kt == KT - 1 ? last_ksteps : TBK / 4;
#if defined(XR_GEMM_DIAG_NO_DMMA)         // timing-only diagnostic (wrong results): the ring and the epilogue without the tensor work
            if (kt >= 0) { __syncwarp(); if (lane == 0) mbar_arrive_cta(&empty[s]); continue; }
#endif
#pragma unroll
            for (int ks = 0; ks < TBK / 4; ++ks) {
                if (ks >= ksteps) break;
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double*>(as + swz(warp_m * 32 + i * 8 + g, ks * 4 + t));
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double*>(bs + swz(warp_n * WCOLS + j * 8 + g, ks * 4 + t));
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&empty[s]);           // this warp no longer reads the stage
        }

#if defined(XR_GEMM_DIAG_NO_STORE)        // timing-only diagnostic (wrong results): the tile is formed and thrown away
        {
            double sink = 0.0;
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) sink += acc[i][j][0] + acc[i][j][1];
            if (sink == 1.2345e301) p.C[0] = sink;
            continue;
        }
#endif
        // epilogue: lane holds C[8i+g][8j+2t], C[8i+g][8j+2t+1]
        int64_t on[NJ][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int64_t col = n0 + warp_n * WCOLS + j * 8 + 2 * t;
            on[j][0] = col < p.N ? (p.offN ? p.offN[col] : col) : -1;
            on[j][1] = col + 1 < p.N ? (p.offN ? p.offN[col + 1] : col + 1) : -1;
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int64_t row = m0 + warp_m * 32 + i * 8 + g;
            if (row >= p.M) continue;
            int64_t om = p.offM ? p.offM[row] : row * p.ldc;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
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
}

// Whole-K variant for K <= 48 (the d = +-1 dimer classes: K = 2n = 36; the K = n diagram products of hermitian-XRCC): the
// three k-tiles of an output tile are ONE stage with ONE barrier pair, so a warp runs its nine k-steps without a barrier or
// a dependent shared-memory round trip in between (the timing diagnostics of profiles/r02u put the ring + DMMAs of the
// per-k-tile kernel at 0.283 ms where the DMMA work is 0.193 ms).  The 48 KB stage is refilled for the CTA's next tile as soon
// as every warp has read it -- thread 0 issues that before its own epilogue -- so the loads fly under the stores; four
// CTAs per SM cover what is left of the latency.  Tiles are handed out by the atomic counter, fetched a tile ahead.
__global__ void __launch_bounds__(TTHREADS, 4)
gemm_tma_wholek_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmTmaParams p) {
    constexpr int MI = 4, NJ = 4;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full, empty;
    __shared__ int64_t next_tile[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp & 1, warp_n = warp >> 1;
    const int KT = (int)((p.K + TBK - 1) / TBK);                                   // 1..3
    const int last_ksteps = (int)((p.K - (int64_t)(KT - 1) * TBK + 3) / 4);
    const int64_t n_tiles = p.tiles_m * p.tiles_n;
    const int64_t per_group = GROUP_M * p.tiles_n;
    auto tile_origin = [&](int64_t tile, int64_t& m0, int64_t& n0) {
        const int64_t first_m = (tile / per_group) * GROUP_M;
        const int64_t group_m = p.tiles_m - first_m < GROUP_M ? p.tiles_m - first_m : GROUP_M;
        m0 = (first_m + (tile % per_group) % group_m) * TBM;
        n0 = ((tile % per_group) / group_m) * TBN;
    };
    if (tid == 0) {
        mbar_init(&full, 1);
        mbar_init(&empty, TTHREADS / 32);
        fence_barrier_init();
    }
    __syncthreads();

    // thread 0: hand-out state.  `upcoming` was requested one tile ago (static striding when there is no counter)
    const bool dynamic = p.tile_counter != nullptr;
    int64_t upcoming = blockIdx.x, issued = 0;
    if (tid == 0 && dynamic) upcoming = (int64_t)atomicAdd(p.tile_counter, 1ull);
    auto issue_tile = [&]() {      // thread 0: publish the next tile id and start its loads (or publish -1)
        int64_t tile = upcoming;
        upcoming = dynamic ? (int64_t)atomicAdd(p.tile_counter, 1ull) : upcoming + gridDim.x;
        if (tile >= n_tiles) tile = -1;
        if (issued > 0) mbar_wait(&empty, (uint32_t)((issued - 1) & 1));       // every warp has read the previous tile's stage
        next_tile[issued & 1] = tile;
        if (tile < 0) {
            mbar_arrive_cta(&full);
        } else {
            int64_t m0, n0;
            tile_origin(tile, m0, n0);
            mbar_expect_tx(&full, (uint32_t)KT * (TILE_A_BYTES + TILE_B_BYTES));
            for (int kt = 0; kt < KT; ++kt) {
                unsigned char* a = smem + (size_t)kt * (TILE_A_BYTES + TILE_B_BYTES);
                tma_load_2d(a, &mapA, kt * TBK, (int)m0, &full);
                tma_load_2d(a + TILE_A_BYTES, &mapB, kt * TBK, (int)n0, &full);
            }
        }
        ++issued;
    };
    if (tid == 0) issue_tile();

    for (int64_t seq = 0;; ++seq) {
        mbar_wait(&full, (uint32_t)(seq & 1));
        const int64_t tile = next_tile[seq & 1];
        if (tile < 0) break;
        int64_t m0, n0;
        tile_origin(tile, m0, n0);
        if (p.offN) {
            const int64_t col = n0 + warp_n * 32 + 2 * t + 8 * (g & 3);
            if (col < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.offN + col));
        }
        if (p.offM) {
            const int64_t row = m0 + warp_m * 32 + lane;
            if (row < p.M) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.offM + row));
        }
        double acc[MI][NJ][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kt = 0; kt < KT; ++kt) {
            const unsigned char* as = smem + (size_t)kt * (TILE_A_BYTES + TILE_B_BYTES);
            const unsigned char* bs = as + TILE_A_BYTES;
            const int ksteps = kt == KT - 1 ? last_ksteps : TBK / 4;
#pragma unroll
            for (int ks = 0; ks < TBK / 4; ++ks) {
                if (ks >= ksteps) break;
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double*>(as + swz(warp_m * 32 + i * 8 + g, ks * 4 + t));
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double*>(bs + swz(warp_n * 32 + j * 8 + g, ks * 4 + t));
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&empty);
        if (tid == 0) issue_tile();           // next tile's loads go out now and land under the stores below
        if (p.alpha != 1.0) {                 // (uniform) the class products have alpha = 1: 32 DMULs per lane less on the pipe the DMMAs use
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    acc[i][j][0] *= p.alpha;
                    acc[i][j][1] *= p.alpha;
                }
        }

        int64_t on[NJ][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int64_t col = n0 + warp_n * 32 + j * 8 + 2 * t;
            on[j][0] = col < p.N ? (p.offN ? p.offN[col] : col) : -1;
            on[j][1] = col + 1 < p.N ? (p.offN ? p.offN[col + 1] : col + 1) : -1;
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int64_t row = m0 + warp_m * 32 + i * 8 + g;
            if (row >= p.M) continue;
            int64_t om = p.offM ? p.offM[row] : row * p.ldc;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                double* dst0 = p.C + om + on[j][0];
                if (on[j][0] >= 0 && on[j][1] == on[j][0] + 1 && (reinterpret_cast<uintptr_t>(dst0) & 15) == 0) {
                    double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
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
                    double v = acc[i][j][e];
                    *dst = p.accumulate ? *dst + v : v;
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [rows][K] row-major doubles with leading dimension ld -> 2-D tensor map, box = 16 (k) x 64 (rows), 128-byte swizzle
bool make_map(CUtensorMap* map, const double* base, int64_t rows, int64_t K, int64_t ld) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
    cuuint32_t estride[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estride,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS;
}

}  // namespace

namespace {

// second pass of a split-K product: fixed-order sum of the K slices (bit-reproducible), then the scatter epilogue
__global__ void __launch_bounds__(256) splitk_finish_kernel(const double* __restrict__ ws, int splits, const GemmTmaParams p) {
    const int64_t total = p.M * p.N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int s = 0; s < splits; ++s) v += ws[(size_t)s * (size_t)total + e];
        const int64_t row = e / p.N, col = e % p.N;
        double* dst = p.C + (p.offM ? p.offM[row] : row * p.ldc) + (p.offN ? p.offN[col] : col);
        v *= p.alpha;
        *dst = p.accumulate ? *dst + v : v;
    }
}

}  // namespace

namespace {

// ------------------------------------------------------------------------------------------------------------------------
// Stream kernel: tall-skinny products with a long contraction, the rho x integral precontractions of hermitian-XRCC
// ([P, n^3..n^4] x [n^3..n^4, 1..n]): the work is READING the density once, so the roofline is HBM, and two things kept the
// general 64 x 64 tile kernel at a quarter of it: (1) it multiplies a full 64-column tile although N is 1..18 -- DMMA-bound
// on zeros -- and (2) densities whose free orbital index sits between contracted ones were first re-ordered by
// xr_permute_copy, a read + write of the whole tensor.  Here a CTA owns <= 64 rows x (8 or 32) columns, every warp 16 rows,
// and A is addressed through a 4-D tensor map: rows m = r1*E2 + r2 at element offset r1*s1 + r2*s2, contraction index
// k = k1*EK2 + k2 at offset k1*sk1 + k2 (k2 contiguous).  One TMA box [16 k2][r2 box][1 k1][r1 box] lands in shared
// memory as [rows][16] in exactly the layout of the dense kernel, so a density [ij, a, b, c, d, e] contracted over (a,b,d,e)
// is streamed where it lies.  K is always split over blockIdx.y (the whole GPU streams the operand); the partial tiles go
// through the same fixed-order second pass as MODE_SPLITK (bit-reproducible).
struct StreamParams {
    int64_t M, N;
    int64_t E2;                 // rows of the inner row group (1 for a plain matrix)
    int r1_box, r2_box;         // rows per tile = r1_box * r2_box <= 64
    int64_t tiles_r2;           // tiles along the inner row group
    int64_t EK2;                // extent of the contiguous contraction group
    int chunks;                 // ceil(EK2 / 16)
    int64_t total_kt;           // EK1 * chunks
    int kt_per_split;
    double* ws;                 // [splits][M][N]
};

constexpr int SSTAGES = 5, SBM = 64;

__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

template <int NJ>      // 8-column blocks per CTA: 1 (N <= 8) or 4 (N <= 32)
__global__ void __launch_bounds__(TTHREADS, 4)
gemm_tma_stream_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const StreamParams p) {
    constexpr int MI = 2, BN = NJ * 8;
    constexpr int A_BYTES = SBM * TBK * 8, B_BYTES = BN * TBK * 8, STAGE_BYTES = A_BYTES + (B_BYTES < 1024 ? 1024 : B_BYTES);
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[SSTAGES], empty[SSTAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t tile = blockIdx.x;
    const int64_t r1_0 = (tile / p.tiles_r2) * p.r1_box, r2_0 = (tile % p.tiles_r2) * p.r2_box;
    const int64_t kt_begin = (int64_t)blockIdx.y * p.kt_per_split;
    const int KT = (int)(p.total_kt - kt_begin < p.kt_per_split ? p.total_kt - kt_begin : p.kt_per_split);
    const uint32_t a_tx = (uint32_t)(p.r1_box * p.r2_box * TBK * 8);

    if (tid == 0) {
        for (int s = 0; s < SSTAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TTHREADS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    int issued = 0;
    auto issue_next = [&]() {       // thread 0
        if (issued >= KT) return;
        const int s = issued % SSTAGES;
        mbar_wait(&empty[s], (uint32_t)((issued / SSTAGES) & 1) ^ 1);
        const int64_t q = kt_begin + issued;
        const int64_t k1 = q / p.chunks;
        const int c = (int)(q % p.chunks);
        unsigned char* a = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full[s], a_tx + B_BYTES);
        tma_load_4d(a, &mapA, c * TBK, (int)r2_0, (int)k1, (int)r1_0, &full[s]);
        tma_load_2d(a + A_BYTES, &mapB, (int)(k1 * p.EK2 + c * TBK), 0, &full[s]);
        ++issued;
    };
    if (tid == 0) {
        for (int s = 0; s < SSTAGES - 1; ++s) issue_next();
    }

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % SSTAGES;
        if (tid == 0) issue_next();
        mbar_wait(&full[s], (uint32_t)(kt / SSTAGES) & 1);
        const unsigned char* as = smem + (size_t)s * STAGE_BYTES;
        const unsigned char* bs = as + A_BYTES;
#pragma unroll
        for (int ks = 0; ks < TBK / 4; ++ks) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double*>(as + swz(warp * 16 + i * 8 + g, ks * 4 + t));
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double*>(bs + swz(j * 8 + g, ks * 4 + t));
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&empty[s]);
    }

    // raw partial tile -> ws[split][m][n]; row r of the tile is (r1_0 + r / r2_box, r2_0 + r % r2_box)
    double* ws = p.ws + (size_t)blockIdx.y * (size_t)p.M * (size_t)p.N;
    const int rows_in_tile = p.r1_box * p.r2_box;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = warp * 16 + i * 8 + g;
        if (r >= rows_in_tile) continue;
        const int64_t r1 = r1_0 + r / p.r2_box, r2 = r2_0 + r % p.r2_box;
        if (r2 >= p.E2) continue;
        const int64_t m = r1 * p.E2 + r2;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int64_t col = j * 8 + 2 * t;
            if (col < p.N) ws[m * p.N + col] = acc[i][j][0];
            if (col + 1 < p.N) ws[m * p.N + col + 1] = acc[i][j][1];
        }
    }
}

// A as [EK2 | E2 | EK1 | E1] with byte strides (8 | s2*8 | sk1*8 | s1*8) and box [16 | r2_box | 1 | r1_box]
bool make_map_4d(CUtensorMap* map, const double* base, int64_t EK2, int64_t E2, int64_t s2, int64_t EK1, int64_t sk1, int64_t E1,
                 int64_t s1, int r2_box, int r1_box) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)EK2, (cuuint64_t)E2, (cuuint64_t)EK1, (cuuint64_t)E1};
    cuuint64_t gstride[3] = {(cuuint64_t)s2 * 8, (cuuint64_t)sk1 * 8, (cuuint64_t)s1 * 8};
    cuuint32_t box[4] = {(cuuint32_t)TBK, (cuuint32_t)r2_box, 1, (cuuint32_t)r1_box};
    cuuint32_t estride[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), gdim, gstride, box, estride,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map_rows(CUtensorMap* map, const double* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estride,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NJ>
int launch_stream(xr_ctx* ctx, const CUtensorMap& mapA, const CUtensorMap& mapB, StreamParams p, int64_t tiles, int64_t splits) {
    constexpr int BN = NJ * 8;
    constexpr size_t SMEM = (size_t)SSTAGES * (SBM * TBK * 8 + (BN * TBK * 8 < 1024 ? 1024 : BN * TBK * 8)) + 1024;
    XR_CUDA(cudaFuncSetAttribute(gemm_tma_stream_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    gemm_tma_stream_kernel<NJ><<<dim3((unsigned)tiles, (unsigned)splits), TTHREADS, SMEM, ctx->stream>>>(mapA, mapB, p);
    XR_CUDA(cudaGetLastError());
    return XR_OK;
}

}  // namespace

// The stream product (see gemm_tma_stream_kernel).  XR_ERR_UNSUPPORTED when a tensor map cannot describe the operands.
int xr_gemm_stream_impl(xr_ctx* ctx, int64_t E1, int64_t s1, int64_t E2, int64_t s2, int64_t EK1, int64_t sk1, int64_t EK2, int64_t N,
                        double alpha, const double* A, const double* B, int64_t ldb, double* C, const int64_t* offM, int64_t ldc,
                        const int64_t* offN, int accumulate) {
    const int64_t M = E1 * E2, K = EK1 * EK2;
    if (N > 32 || M >= (1ll << 31) || K >= (1ll << 31) || E1 < 1 || E2 < 1 || EK1 < 1 || EK2 < 1) return XR_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (ldb & 1)) return XR_ERR_UNSUPPORTED;
    if ((E2 > 1 && (s2 & 1)) || (EK1 > 1 && (sk1 & 1)) || (E1 > 1 && (s1 & 1))) return XR_ERR_UNSUPPORTED;      // 16-byte strides
    StreamParams p{};
    p.M = M;
    p.N = N;
    p.E2 = E2;
    p.r2_box = (int)(E2 < SBM ? E2 : SBM);
    p.r1_box = SBM / p.r2_box;
    if (p.r1_box > E1) p.r1_box = (int)E1;
    if (p.r1_box > 256 || p.r2_box > 256) return XR_ERR_UNSUPPORTED;
    p.tiles_r2 = (E2 + p.r2_box - 1) / p.r2_box;
    const int64_t tiles = ((E1 + p.r1_box - 1) / p.r1_box) * p.tiles_r2;
    p.EK2 = EK2;
    p.chunks = (int)((EK2 + TBK - 1) / TBK);
    p.total_kt = EK1 * p.chunks;
    if (tiles >= 65536ll * 32768 || tiles < 1) return XR_ERR_UNSUPPORTED;
    int64_t splits = (8 * (int64_t)ctx->sm_count + tiles - 1) / tiles;
    if (splits > p.total_kt / 8) splits = p.total_kt / 8;
    if (splits > 1024) splits = 1024;
    if (splits < 1) splits = 1;
    p.kt_per_split = (int)((p.total_kt + splits - 1) / splits);
    splits = (p.total_kt + p.kt_per_split - 1) / p.kt_per_split;
    CUtensorMap mapA, mapB;
    const int bn = N <= 8 ? 8 : 32;
    if (!make_map_4d(&mapA, A, EK2, E2, E2 > 1 ? s2 : 2, EK1, EK1 > 1 ? sk1 : 2, E1, E1 > 1 ? s1 : 2, p.r2_box, p.r1_box) ||      // (extent-1 dims: any 16-byte stride)
        !make_map_rows(&mapB, B, N, K, ldb, bn))
        return XR_ERR_UNSUPPORTED;
    int rc = xr_ensure_scratch(ctx, (size_t)splits * (size_t)M * (size_t)N * sizeof(double));
    if (rc != XR_OK) return rc;
    p.ws = static_cast<double*>(ctx->scratch);
    rc = N <= 8 ? launch_stream<1>(ctx, mapA, mapB, p, tiles, splits) : launch_stream<4>(ctx, mapA, mapB, p, tiles, splits);
    if (rc != XR_OK) return rc;
    GemmTmaParams f{M, N, K, alpha, C, offM, ldc, offN, accumulate, 0, 0, nullptr, 0, nullptr, nullptr};
    int64_t blocks = (M * N + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
    splitk_finish_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p.ws, (int)splits, f);
    XR_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return XR_OK;
}

extern "C" int xr_gemm_stream(xr_ctx* ctx, int64_t E1, int64_t s1, int64_t E2, int64_t s2, int64_t EK1, int64_t sk1, int64_t EK2,
                              int64_t N, double alpha, const double* A, const double* B, int64_t ldb, double* C,
                              const int64_t* offM, int64_t ldc, const int64_t* offN, int accumulate) {
    XR_REQUIRE(ctx, "xr_gemm_stream: null ctx");
    if (E1 <= 0 || E2 <= 0 || N <= 0) return XR_OK;
    XR_REQUIRE(EK1 >= 1 && EK2 >= 1 && A && B && C, "xr_gemm_stream: null pointer or empty contraction");
    XR_REQUIRE(offM || ldc >= 1, "xr_gemm_stream: need offM or ldc");
    XR_REQUIRE(ldb >= EK1 * EK2, "xr_gemm_stream: ldb smaller than K");
    int rc = xr_gemm_stream_impl(ctx, E1, s1, E2, s2, EK1, sk1, EK2, N, alpha, A, B, ldb, C, offM, ldc, offN, accumulate);
    if (rc == XR_ERR_UNSUPPORTED)
        xr_set_error("xr_gemm_stream: operands not expressible (N <= 32, 16-byte aligned bases and even strides are required)");
    return rc;
}

int xr_gemm_stream_impl(xr_ctx* ctx, int64_t E1, int64_t s1, int64_t E2, int64_t s2, int64_t EK1, int64_t sk1, int64_t EK2, int64_t N,
                        double alpha, const double* A, const double* B, int64_t ldb, double* C, const int64_t* offM, int64_t ldc,
                        const int64_t* offN, int accumulate);

// returns XR_ERR_UNSUPPORTED when the operands cannot be described by a tensor map (caller falls back to cp.async staging)
int xr_gemm_scatter_tma(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double* C, const int64_t* offM, int64_t ldc, const int64_t* offN,
                        int accumulate) {
    if (K < 1 || M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return XR_ERR_UNSUPPORTED;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, M, K, lda) || !make_map(&mapB, B, N, K, ldb)) return XR_ERR_UNSUPPORTED;
    GemmTmaParams p{M, N, K, alpha, C, offM, ldc, offN, accumulate, (M + TBM - 1) / TBM, (N + TBN - 1) / TBN, nullptr, 0, nullptr, nullptr};
    const int64_t tiles = p.tiles_m * p.tiles_n;
    XR_REQUIRE(tiles < (1ll << 31), "xr_gemm_scatter: too many tiles (%lld)", (long long)tiles);
    // Few output tiles but a long contraction (the rho x V precontractions of hermitian-XRCC: [P, n^4] x [n^4, 1..n]): split K
    // over blockIdx.y so the whole GPU streams the operand, partial tiles to scratch, fixed-order second pass.
    const int64_t KT = (K + TBK - 1) / TBK;
    int64_t splits = 1;
    if (tiles < 2 * (int64_t)ctx->sm_count && KT >= 16) {
        splits = (4 * (int64_t)ctx->sm_count + tiles - 1) / tiles;
        if (splits > KT / 8) splits = KT / 8;
        if (splits > 512) splits = 512;
    }
    if (splits >= 2 && N <= 32) {
        // skinny output: the stream kernel (8- or 32-column tiles instead of 64: the DMMA pipe no longer multiplies zeros)
        int rc = xr_gemm_stream_impl(ctx, M, lda, 1, 0, 1, 0, K, N, alpha, A, B, ldb, C, offM, ldc, offN, accumulate);
        if (rc != XR_ERR_UNSUPPORTED) return rc;
    }
    if (splits >= 2) {
        p.kt_per_split = (int)((KT + splits - 1) / splits);
        splits = (KT + p.kt_per_split - 1) / p.kt_per_split;
        int rc = xr_ensure_scratch(ctx, (size_t)splits * (size_t)M * (size_t)N * sizeof(double));
        if (rc != XR_OK) return rc;
        p.ws = static_cast<double*>(ctx->scratch);
        XR_CUDA(cudaFuncSetAttribute(gemm_tma_scatter_kernel<MODE_SPLITK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        gemm_tma_scatter_kernel<MODE_SPLITK><<<dim3((unsigned)tiles, (unsigned)splits), TTHREADS, TMA_SMEM, ctx->stream>>>(mapA, mapB, p);
        XR_CUDA(cudaGetLastError());
        int64_t blocks = (M * N + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
        splitk_finish_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p.ws, (int)splits, p);
        XR_CUDA(cudaGetLastError());
        ctx->launches += 2;
        return XR_OK;
    }
    if (XR_GEMM_WHOLEK && KT <= 3) {
        const int64_t resident = (int64_t)ctx->sm_count * 4;
        if (tiles > resident) {
            if (!ctx->counters) XR_CUDA(cudaMalloc(&ctx->counters, 256));
            XR_CUDA(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->stream));
            p.tile_counter = static_cast<unsigned long long*>(ctx->counters);
        }
        XR_CUDA(cudaFuncSetAttribute(gemm_tma_wholek_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        gemm_tma_wholek_kernel<<<(unsigned)(tiles < resident ? tiles : resident), TTHREADS, TMA_SMEM, ctx->stream>>>(mapA, mapB, p);
        XR_CUDA(cudaGetLastError());
        ctx->launches++;
        return XR_OK;
    }
    if (XR_GEMM_VARIANT == 0 || KT > XR_GEMM_PERSISTENT_MAX_KT) {
        XR_CUDA(cudaFuncSetAttribute(gemm_tma_scatter_kernel<MODE_SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        gemm_tma_scatter_kernel<MODE_SCATTER><<<(unsigned)tiles, TTHREADS, TMA_SMEM, ctx->stream>>>(mapA, mapB, p);
        XR_CUDA(cudaGetLastError());
        ctx->launches++;
        return XR_OK;
    }
    if (XR_GEMM_VARIANT == 3 && tiles > (int64_t)ctx->sm_count * PCTAS) {      // more tiles than resident CTAs: hand them out dynamically
        if (!ctx->counters) XR_CUDA(cudaMalloc(&ctx->counters, 256));
        XR_CUDA(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->stream));
        p.tile_counter = static_cast<unsigned long long*>(ctx->counters);
    }
    // persistent CTAs, PCTAS per SM (48 KB of ring + <= 128 registers each at the default depth): one launch-and-fill per CTA
    // instead of per tile
    const int64_t resident = (int64_t)ctx->sm_count * PCTAS;
    XR_CUDA(cudaFuncSetAttribute(gemm_tma_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PERSISTENT_SMEM));
    gemm_tma_persistent_kernel<<<(unsigned)(tiles < resident ? tiles : resident), PTHREADS, PERSISTENT_SMEM, ctx->stream>>>(mapA, mapB, p);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}

namespace {

// fixed-shape two-level sum of the per-CTA partials (bit-reproducible: no atomics, fixed assignment and order)
constexpr int RED_BLOCKS = 256, RED_THREADS = 256;

__global__ void __launch_bounds__(RED_THREADS) reduce_partials_kernel(const double* __restrict__ partials, int64_t count,
                                                                        double* __restrict__ block_out) {
    __shared__ double sh[2][RED_THREADS];
    const int64_t per_block = (count + RED_BLOCKS - 1) / RED_BLOCKS;
    const int64_t lo = blockIdx.x * per_block, hi = lo + per_block < count ? lo + per_block : count;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += RED_THREADS) {
        s1 += partials[2 * i];
        s2 += partials[2 * i + 1];
    }
    sh[0][threadIdx.x] = s1;
    sh[1][threadIdx.x] = s2;
    __syncthreads();
    for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        block_out[2 * blockIdx.x] = sh[0][0];
        block_out[2 * blockIdx.x + 1] = sh[1][0];
    }
}

__global__ void finalize_moments_kernel(const double* __restrict__ block_out, double alpha, double* moments) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < RED_BLOCKS; ++i) {
            t1 += block_out[2 * i];
            t2 += block_out[2 * i + 1];
        }
        moments[0] += alpha * t1;
        moments[1] += alpha * alpha * t2;
    }
}

}  // namespace

// moments[0] += alpha * sum(C), moments[1] += alpha^2 * sum(C^2) for C = A.B^T, without ever storing C: the consumer for
// dimer blocks that do not fit in HBM (cfg5: 1e12 elements).  Needs 16-byte-aligned operands (tensor maps).
extern "C" int xr_gemm_reduce(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                              const double* B, int64_t ldb, double* moments) {
    XR_REQUIRE(ctx, "xr_gemm_reduce: null ctx");
    if (M <= 0 || N <= 0) return XR_OK;
    XR_REQUIRE(K >= 1 && A && B && moments, "xr_gemm_reduce: null pointer or K < 1");
    XR_REQUIRE(lda >= K && ldb >= K, "xr_gemm_reduce: lda/ldb smaller than K");
    XR_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
               "xr_gemm_reduce: operands must be 16-byte aligned with even leading dimensions");
    XR_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "xr_gemm_reduce: dimension too large");
    CUtensorMap mapA, mapB;
    XR_REQUIRE(make_map(&mapA, A, M, K, lda) && make_map(&mapB, B, N, K, ldb), "xr_gemm_reduce: cuTensorMapEncodeTiled failed");
    GemmTmaParams p{M, N, K, alpha, nullptr, nullptr, 0, nullptr, 0, (M + TBM - 1) / TBM, (N + TBN - 1) / TBN, nullptr, 0, nullptr, nullptr};
    const int64_t tiles = p.tiles_m * p.tiles_n;
    XR_REQUIRE(tiles < (1ll << 31), "xr_gemm_reduce: too many tiles (%lld)", (long long)tiles);
    int rc = xr_ensure_scratch(ctx, (size_t)(tiles + RED_BLOCKS) * 2 * sizeof(double) + 256);
    if (rc != XR_OK) return rc;
    p.partials = static_cast<double*>(ctx->scratch);
    double* block_out = p.partials + 2 * tiles;
    XR_CUDA(cudaFuncSetAttribute(gemm_tma_scatter_kernel<MODE_REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
    gemm_tma_scatter_kernel<MODE_REDUCE><<<(unsigned)tiles, TTHREADS, TMA_SMEM, ctx->stream>>>(mapA, mapB, p);
    XR_CUDA(cudaGetLastError());
    reduce_partials_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(p.partials, tiles, block_out);
    XR_CUDA(cudaGetLastError());
    finalize_moments_kernel<<<1, 32, 0, ctx->stream>>>(block_out, alpha, moments);
    XR_CUDA(cudaGetLastError());
    ctx->launches += 3;
    return XR_OK;
}
