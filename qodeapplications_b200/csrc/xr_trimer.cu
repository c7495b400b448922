// xr_trimer_stream: streamed three-factor FP64 contraction for the trimer classes.
//
//   T[a,b,c] = alpha * sum_{r,s<n} W[a,r,s] * beta[b,r] * gamma[c,s]
//
// A persistent CTA takes a work item of TA=8 values of a times TB=16 values of b (128 "rows"),
// forms X[(a,b),s] = sum_r W[a,r,s] beta[b,r] once in shared memory (negligible: 2*128*n^2 flop
// against 2*128*Pc*n), keeps its DMMA A-fragments of X in registers, and then streams gamma in
// tiles of 128 c through a shared-memory ring filled by 1-D bulk TMA (cp.async.bulk + mbarrier).
// Every 128x128 tile of T lives only in DMMA accumulators and is handed to the consumer (moment
// reducer or scatter-store), so the 1e13-element trimer blocks never touch HBM.
//
// Warp-specialised: a producer warpgroup (one elected lane) streams gamma tiles into a SLOTS-deep
// ring with full/empty mbarriers per slot; the 8 consumer warps never meet at a block-wide barrier
// inside the tile loop.  setmaxnreg moves the producer warpgroup's registers to the consumers.
// The same lane also prefetches the NEXT work item's operands (TA rows of W, TB rows of beta) into a
// staging buffer by bulk TMA while the current item's tiles are consumed, so forming X at an item
// boundary reads shared memory only (the global-load version of that step -- 18 dependent L2 round
// trips per output -- left the FP64 pipe idle for ~8 % of the kernel: stall_barrier in profiles/r01j).
//
// Roofline: FP64 pipe (DMMA.8x8x4 and DFMA share it on sm_100: 64 FMA/clk/SM, 37.2 TFLOP/s measured).
// Algorithmic flops are 2*n per element.  k = n is split as 4*KS (DMMA k-steps) + TAIL (0..2 leftover
// k handled by DFMA on the accumulators), so n = 18 costs 18 FMA per element instead of the 20 a
// zero-padded fifth DMMA step would; the moment reducer adds 2 FP64 ops per element on the same pipe.
#include "xr_common.cuh"
#include <algorithm>
#include <utility>
#include <vector>

enum { XR_TRIMER_THRESHOLD = 2, XR_TRIMER_SAMPLE = 3 };      // internal mode numbers of the EXTRA consumers

namespace {

constexpr int TA = 8, TB = 16, ROWS = TA * TB;

// The first moment is LINEAR in the tile, so it does not need the stream: sum_abc T = sum_a sum_rs W[a,rs] (sum_b beta[b,r])
// (sum_c gamma[c,s]), three tiny reductions (trimer_sum_kernel).  Only the second moment is accumulated element by element:
// one FP64 instruction per element on the pipe the DMMAs share instead of two (+4 % on the whole kernel).  Set to 1 to
// accumulate the sum from the streamed elements as well (the two agree to rounding; tests/test_general_gpu.py).
#ifndef XR_TRIMER_STREAM_SUM
#define XR_TRIMER_STREAM_SUM 0
#endif
constexpr bool STREAM_SUM = XR_TRIMER_STREAM_SUM != 0;
// XR_TRIMER_SEP_TAIL keeps the DFMA k-tail of a tile in its own basic block (behind a branch that is always taken when
// TAIL > 0), so that ptxas does not interleave the 64-128 DFMAs with the last DMMAs of the tile: interleaved, the two
// instruction kinds contend for the one FP64 pipe in an order that leaves issue slots empty (31.5 -> 32.9 TFLOP/s at
// n = 18 on B200).  tools/trimer_variants.py builds one library per -D setting and times them in one GPU call; what was
// tried and rejected is listed in DESIGN.md (trimer kernel, "variants measured").
#ifndef XR_TRIMER_SEP_TAIL
#define XR_TRIMER_SEP_TAIL 1
#endif
constexpr bool SEP_TAIL = XR_TRIMER_SEP_TAIL != 0;

struct TrimerParams {
    int n;
    int64_t Pb, Pc;
    double alpha;
    const double* W;
    int64_t ldw;
    const double* betaP;    // [Pb][KP], zero padded in k
    const double* gammaP;   // [c_tiles*CT][GS], zero padded in k and rows
    const double* gammaT;   // [c_tiles*CT][2]: the TAIL columns gamma[c][4*KS .. 4*KS+1] again, compact (conflict-free tail loads)
    int64_t a_begin, a_end;
    int mode;
    double* partials;       // [grid][2]
    double* C;
    const int64_t* offA;
    const int64_t* offB;
    const int64_t* offC;
    int64_t tiles_b, n_items;
    int c_tiles;
    int staged;             // W rows are 16-byte aligned: the producer stages each item's W/beta rows by bulk TMA
    // ---- EXTRA consumers (template instantiations of their own: the moment / materialise stream is compiled without them)
    const int64_t* item_list;     // SAMPLE: the work items that hold a sample (n_items = its length); null = every item
    const int32_t* sample_ptr;    // SAMPLE: [n_items + 1] CSR over `samples`, by position in item_list
    const int4* samples;          // SAMPLE: {row inside the item (a_local*TB + b_local), c, index into sample_out, unused}
    double* sample_out;           // SAMPLE: [count]
    double tau;                   // THRESHOLD: keep |alpha*T| > tau
    unsigned long long* counter;  // THRESHOLD: number of kept elements (may exceed capacity: the caller re-runs)
    int64_t capacity;
    int64_t* idx_out;             // THRESHOLD: [capacity] offA[a] + offB[b] + offC[c]
    double* val_out;              // THRESHOLD: [capacity]
};

// WN = consumer warps along c: 2 -> 8 warps, each 32 rows x 64 columns (128-column gamma tiles, 240 registers);
//                              3 -> 12 warps, each 32 rows x 32 columns (96-column tiles, 152 registers, 3 warps per scheduler)
template <int KS, int TAIL, int WN>
struct TrimerCfg {
    static constexpr int NJ = WN == 2 ? 8 : 4;                                    // 8-column DMMA blocks per warp
    static constexpr int CT = WN * NJ * 8;                                        // gamma rows (= T columns) per tile
    static constexpr int CONSUMER_WARPS = 4 * WN, CONSUMER_THREADS = CONSUMER_WARPS * 32;
    static constexpr int THREADS = CONSUMER_THREADS + 128;                        // + one producer warpgroup (register allocation is per 4 warps)
    static constexpr int CONSUMER_REGS = WN == 2 ? 240 : 152, PRODUCER_REGS = 24;
    static constexpr int KP = 4 * KS + (TAIL ? 4 : 0);                            // packed row width (k, zero padded)
    static constexpr int GS = (KP % 16 == 4 || KP % 16 == 12) ? KP : KP + 4;       // conflict-free fragment stride
    static constexpr bool AREG = KS <= 5;                                         // A fragments live in registers
    static constexpr int SLOTS = KS <= 5 ? 4 : 2;                                 // gamma ring depth (shared memory bound)
    static constexpr bool STAGE = KS <= 5;                                        // item operands prefetched into shared memory
    static constexpr int WS = STAGE ? (KP * KP + 1) / 2 * 2 : 0;                  // staged W row (n*n <= KP*KP doubles, even)
    static constexpr int STAGE_DOUBLES = STAGE ? TA * WS + TB * KP : 0;
    static constexpr size_t SMEM = (size_t)(ROWS + SLOTS * CT) * GS * sizeof(double) + (size_t)SLOTS * CT * 2 * sizeof(double) +
                                   (size_t)STAGE_DOUBLES * sizeof(double) + (2 * SLOTS + 1) * sizeof(uint64_t) + 64;
};

template <int CONSUMER_THREADS>
__device__ __forceinline__ void consumer_barrier() {   // named barrier 1: the consumer warps only
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_THREADS) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// X[(a,b), s] = sum_r W[a, r*n + s] * beta[b, r] for one work item, r ascending (one rounding order for both sources).
// Consumer warp w forms the TB rows of a = a0 + w; lane <-> s.  SMEM_SRC: W rows and beta rows come from the staging
// buffer the producer filled by bulk TMA; otherwise straight from global memory with the r loop unrolled so that the
// loads of several r are in flight together.
template <bool SMEM_SRC, int KP, int GS>
__device__ __forceinline__ void build_item_X(double* __restrict__ Xs, const double* __restrict__ Wsrc, int64_t w_row_stride,
                                             const double* __restrict__ Bsrc, int n, int warp, int lane, int na, int nb) {
    for (int al = warp; al < TA; al += 8) {
        for (int s = lane; s < KP; s += 32) {
            double acc[TB];
#pragma unroll
            for (int b = 0; b < TB; ++b) acc[b] = 0.0;
            if (al < na && s < n) {
                const double* w = Wsrc + (int64_t)al * w_row_stride + s;
                if (SMEM_SRC) {
                    for (int r = 0; r < n; ++r) {
                        const double wv = w[r * n];
#pragma unroll
                        for (int b = 0; b < TB; ++b) acc[b] = fma(wv, Bsrc[b * KP + r], acc[b]);
                    }
                } else {
                    constexpr int U = 6;
                    for (int r0 = 0; r0 < n; r0 += U) {
                        double wv[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) wv[u] = r0 + u < n ? __ldg(w + (int64_t)(r0 + u) * n) : 0.0;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (r0 + u < n) {
#pragma unroll
                                for (int b = 0; b < TB; ++b) {
                                    const double bv = b < nb ? __ldg(Bsrc + b * KP + r0 + u) : 0.0;
                                    acc[b] = fma(wv[u], bv, acc[b]);
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < TB; ++b) Xs[(al * TB + b) * GS + s] = b < nb ? acc[b] : 0.0;
        }
    }
}

template <int KS, int TAIL, int WN, bool EXTRA>
__global__ void __launch_bounds__(TrimerCfg<KS, TAIL, WN>::THREADS, 1) trimer_stream_kernel(const TrimerParams p) {
    using Cfg = TrimerCfg<KS, TAIL, WN>;
    const int mode = p.mode;
    constexpr int KP = Cfg::KP, GS = Cfg::GS, SLOTS = Cfg::SLOTS, CT = Cfg::CT;
    constexpr int CONSUMER_WARPS = Cfg::CONSUMER_WARPS, CONSUMER_THREADS = Cfg::CONSUMER_THREADS;
    constexpr bool AREG = Cfg::AREG, STAGE = Cfg::STAGE;
    constexpr int MI = 4, NJ = Cfg::NJ, WCOLS = NJ * 8, WS = Cfg::WS;
    static_assert(CONSUMER_WARPS >= 8, "build_item_X spreads the TA rows of W over 8 consumer warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Gs = reinterpret_cast<double*>(smem_raw);                  // [SLOTS][CT][GS]  (bulk-copy destinations: 16B aligned)
    double* Gt = Gs + SLOTS * CT * GS;                                  // [SLOTS][CT][2]   tail columns, compact
    double* Xs = Gt + SLOTS * CT * 2;                                   // [ROWS][GS]
    double* Wst = Xs + ROWS * GS;                                       // [TA][WS]   next item's W rows   (STAGE only)
    double* Bst = Wst + TA * WS;                                        // [TB][KP]   next item's beta rows (STAGE only)
    uint64_t* full = reinterpret_cast<uint64_t*>(Xs + ROWS * GS + Cfg::STAGE_DOUBLES);   // [SLOTS]
    uint64_t* empty = full + SLOTS;                                     // [SLOTS]
    uint64_t* wfull = empty + SLOTS;                                    // staging buffer filled (one phase per work item)
    __shared__ double red[2][CONSUMER_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t TILE_BYTES = CT * GS * sizeof(double);
    constexpr uint32_t TAIL_BYTES = TAIL ? CT * 2 * sizeof(double) : 0;
    const bool staged = STAGE && p.staged;

    if (tid == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    __syncthreads();

    int64_t my_items = 0;
    if ((int64_t)blockIdx.x < p.n_items) my_items = (p.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t total_tiles = my_items * p.c_tiles;

    if (warp >= CONSUMER_WARPS) {
        // -------------------------------------------------- producer warpgroup (one lane works)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::PRODUCER_REGS));     // hand registers to the consumers
        if (warp == CONSUMER_WARPS && lane == 0 && my_items > 0) {
            // coordinates of the next item to stage; items advance by gridDim.x
            int64_t ia = (int64_t)blockIdx.x / p.tiles_b, ib = (int64_t)blockIdx.x % p.tiles_b;
            int64_t next_k = blockIdx.x;      // EXTRA: position in the item list of the next item to stage
            const uint32_t row_bytes = (uint32_t)(((p.n * p.n + 1) / 2 * 2) * sizeof(double));
            auto stage_item = [&]() {
                if (EXTRA) {
                    const int64_t id = p.item_list ? p.item_list[next_k] : next_k;
                    ia = id / p.tiles_b;
                    ib = id % p.tiles_b;
                    next_k += gridDim.x;
                }
                const int64_t a0 = p.a_begin + ia * TA, b0 = ib * TB;
                const int na = (int)(p.a_end - a0 < TA ? p.a_end - a0 : TA);
                const int nb = (int)(p.Pb - b0 < TB ? p.Pb - b0 : TB);
                const uint32_t beta_bytes = (uint32_t)(nb * KP * sizeof(double));
                mbar_expect_tx(wfull, (uint32_t)na * row_bytes + beta_bytes);
                for (int i = 0; i < na; ++i) bulk_copy_g2s(Wst + i * WS, p.W + (a0 + i) * p.ldw, row_bytes, wfull);
                bulk_copy_g2s(Bst, p.betaP + b0 * KP, beta_bytes, wfull);
                if (!EXTRA) {
                    ib += gridDim.x;
                    while (ib >= p.tiles_b) {
                        ib -= p.tiles_b;
                        ++ia;
                    }
                }
            };
            int64_t staged_items = 0, next_stage_q = SLOTS;
            if (staged) {
                stage_item();
                staged_items = 1;
            }
            int ct = 0;
            // Item it+1 is staged once tile (it, 0) has been released by every consumer warp -- they all built X(it)
            // from the staging buffer before touching that tile -- i.e. just before global tile it*c_tiles + SLOTS is
            // issued.  The loop runs SLOTS steps past the last tile so that rule also covers c_tiles < SLOTS.
            for (int64_t q = 0; q < total_tiles + SLOTS; ++q) {
                const int slot = (int)(q % SLOTS);
                const uint32_t round = (uint32_t)(q / SLOTS);
                mbar_wait(&empty[slot], (round & 1) ^ 1);     // passes at once on the first lap
                if (q == next_stage_q) {
                    if (staged && staged_items < my_items) {
                        stage_item();
                        ++staged_items;
                    }
                    next_stage_q += p.c_tiles;
                }
                if (q < total_tiles) {
                    mbar_expect_tx(&full[slot], TILE_BYTES + TAIL_BYTES);
                    bulk_copy_g2s(Gs + (size_t)slot * CT * GS, p.gammaP + (size_t)ct * CT * GS, TILE_BYTES, &full[slot]);
                    if (TAIL) bulk_copy_g2s(Gt + (size_t)slot * CT * 2, p.gammaT + (size_t)ct * CT * 2, TAIL_BYTES, &full[slot]);
                    if (++ct == p.c_tiles) ct = 0;
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    double s1p[NJ], s2p[NJ];     // NJ independent chains each: the moment epilogue must not be one serial FP64 dependency
#pragma unroll
    for (int j = 0; j < NJ; ++j) s1p[j] = s2p[j] = 0.0;
    const int n = p.n;
    int64_t q = 0;
    uint32_t item_parity = 0;

    for (int64_t item_k = blockIdx.x; item_k < p.n_items; item_k += gridDim.x) {
        const int64_t item = EXTRA && p.item_list ? p.item_list[item_k] : item_k;
        const int64_t a0 = p.a_begin + (item / p.tiles_b) * TA;
        const int64_t b0 = (item % p.tiles_b) * TB;
        const int na = (int)(p.a_end - a0 < TA ? p.a_end - a0 : TA);
        const int nb = (int)(p.Pb - b0 < TB ? p.Pb - b0 : TB);

        consumer_barrier<CONSUMER_THREADS>();    // every consumer is done with the previous item's Xs
        if (STAGE && staged) {
            mbar_wait(wfull, item_parity);
            item_parity ^= 1;
            build_item_X<true, KP, GS>(Xs, Wst, WS, Bst, n, warp, lane, na, nb);
        } else {
            build_item_X<false, KP, GS>(Xs, p.W + a0 * p.ldw, p.ldw, p.betaP + b0 * KP, n, warp, lane, na, nb);
        }
        consumer_barrier<CONSUMER_THREADS>();

        const double* xs = Xs + (32 * wm + g) * GS + t;
        double areg[AREG ? MI : 1][AREG ? KS : 1];
        if (AREG) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) areg[i][ks] = xs[i * 8 * GS + 4 * ks];
        }
        double atail[MI][TAIL ? TAIL : 1];     // X[row][4*KS + tt] for this lane's accumulator rows
        if (TAIL) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int tt = 0; tt < TAIL; ++tt) atail[i][tt] = Xs[(32 * wm + 8 * i + g) * GS + 4 * KS + tt];
        }

        for (int ct = 0; ct < p.c_tiles; ++ct, ++q) {
            const int slot = (int)(q % SLOTS);
            const double* gtile = Gs + (size_t)slot * CT * GS;
            const double* gs = gtile + (WCOLS * wn + g) * GS + t;
            mbar_wait(&full[slot], (uint32_t)(q / SLOTS) & 1);

            double acc[MI][NJ][2];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double b[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = gs[j * 8 * GS + 4 * ks];
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const double a = AREG ? areg[AREG ? i : 0][AREG ? ks : 0] : xs[i * 8 * GS + 4 * ks];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        if (ks == 0)
                            dmma_m8n8k4_zero(acc[i][j][0], acc[i][j][1], a, b[j]);
                        else
                            dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a, b[j]);
                    }
                }
            }
            if (TAIL && (!SEP_TAIL || p.n > 4 * KS)) {      // always taken when TAIL > 0 (n > 4*KS): see SEP_TAIL
                // leftover k (n - 4*KS <= 2) on the accumulator layout: lane owns rows 8i+g, columns 8j+2t+{0,1}.
                // The tail columns come from the compact [c][2] copy: a quad's four 32-byte reads cover 128
                // contiguous bytes (no bank conflicts; the strided [c][GS] rows would give 2-way conflicts).
                const double2* gt = reinterpret_cast<const double2*>(Gt + (size_t)slot * CT * 2) + WCOLS * wn + 2 * t;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double2 v = gt[8 * j + e];
#pragma unroll
                        for (int i = 0; i < MI; ++i) {
                            acc[i][j][e] = fma(atail[i][0], v.x, acc[i][j][e]);
                            if (TAIL == 2) acc[i][j][e] = fma(atail[i][TAIL - 1], v.y, acc[i][j][e]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);      // this warp no longer reads the slot
            if (EXTRA && mode == XR_TRIMER_THRESHOLD) {
                // compaction consumer: every lane counts its kept elements, one warp scan + ONE atomic per warp-tile reserves
                // a contiguous run of the output list, a second pass over the accumulators fills it
                const int64_t c_base = (int64_t)ct * CT + WCOLS * wn + 2 * t;
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int row = 32 * wm + 8 * i + g;
                    const bool row_ok = a0 + row / TB < p.a_end && b0 + row % TB < p.Pb;
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            cnt += (row_ok && c_base + 8 * j + e < p.Pc && fabs(p.alpha * acc[i][j][e]) > p.tau) ? 1 : 0;
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total > 0) {
                    unsigned long long base = 0;
                    if (lane == 31) base = atomicAdd(p.counter, (unsigned long long)total);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    int64_t slot = (int64_t)base + incl - cnt;
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        const int row = 32 * wm + 8 * i + g;
                        const int64_t a = a0 + row / TB, b = b0 + row % TB;
                        const bool row_ok = a < p.a_end && b < p.Pb;
                        const int64_t oab = row_ok ? p.offA[a] + p.offB[b] : 0;
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int64_t c = c_base + 8 * j + e;
                                const double v = p.alpha * acc[i][j][e];
                                if (row_ok && c < p.Pc && fabs(v) > p.tau) {
                                    if (slot < p.capacity) {
                                        p.idx_out[slot] = oab + p.offC[c];
                                        p.val_out[slot] = v;
                                    }
                                    ++slot;
                                }
                            }
                    }
                }
            } else if (EXTRA && mode == XR_TRIMER_SAMPLE) {
                // sampled-element consumer: the few requested elements of THIS item that fall into this gamma tile are picked
                // out of the accumulators by the lane that owns them (acc[i][j][e] <-> row 32*wm+8i+g, column WCOLS*wn+8j+2t+e)
                const int s_begin = p.sample_ptr[item_k], s_end = p.sample_ptr[item_k + 1];
                for (int sidx = s_begin; sidx < s_end; ++sidx) {
                    const int4 rec = p.samples[sidx];
                    const int col = rec.y - ct * CT;
                    if (col < 0 || col >= CT) continue;
                    const int r = rec.x;
                    if ((r >> 5) != wm || col / WCOLS != wn || (r & 7) != g || ((col & 7) >> 1) != t) continue;
                    const int si = (r & 31) >> 3, sj = (col % WCOLS) >> 3, se = col & 1;
                    double v = 0.0;
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                if (i == si && j == sj && e == se) v = acc[i][j][e];
                    p.sample_out[rec.z] = p.alpha * v;
                }
            } else if (mode == XR_TRIMER_REDUCE) {
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        if (STREAM_SUM) s1p[j] += acc[i][j][0] + acc[i][j][1];
                        s2p[j] = fma(acc[i][j][0], acc[i][j][0], s2p[j]);
                    }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) s2p[j] = fma(acc[i][j][1], acc[i][j][1], s2p[j]);
            } else {
                const int64_t c_base = (int64_t)ct * CT + WCOLS * wn + 2 * t;
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int row = 32 * wm + 8 * i + g;
                    const int64_t a = a0 + row / TB, b = b0 + row % TB;
                    if (a >= p.a_end || b >= p.Pb) continue;
                    const int64_t oab = p.offA[a] + p.offB[b];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int64_t c = c_base + 8 * j;
                        if (c < p.Pc) p.C[oab + p.offC[c]] = p.alpha * acc[i][j][0];
                        if (c + 1 < p.Pc) p.C[oab + p.offC[c + 1]] = p.alpha * acc[i][j][1];
                    }
                }
            }
        }
    }

    if (mode == XR_TRIMER_REDUCE) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            s1 += s1p[j];
            s2 += s2p[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            red[0][warp] = s1;
            red[1][warp] = s2;
        }
        consumer_barrier<CONSUMER_THREADS>();
        if (tid == 0) {
            double t1 = 0.0, t2 = 0.0;
            for (int w = 0; w < CONSUMER_WARPS; ++w) {
                t1 += red[0][w];
                t2 += red[1][w];
            }
            p.partials[2 * blockIdx.x] = t1;
            p.partials[2 * blockIdx.x + 1] = t2;
        }
    }
}

// First moment from the factor sums:  sum_{a in [a_begin,a_end), b, c} T[a,b,c] = alpha * sum_a sum_{rs} W[a,rs] * bsum[r] * gsum[s],
// bsum = column sums of beta, gsum = column sums of gamma.  Every assignment of work to threads and every reduction tree
// below is fixed by the sizes alone, so the result is bit-reproducible.
constexpr int SUM_THREADS = 256, SUM_WARPS = SUM_THREADS / 32, MAX_N = 48;

// block 0: bsum, block 1: gsum -> colsum[which][0..n).  A thread owns rows tid, tid + 256, ... and adds up whole rows.
__global__ void __launch_bounds__(SUM_THREADS) trimer_colsum_kernel(int n, int64_t Pb, int64_t Pc, const double* __restrict__ beta,
                                                                    int64_t ldbeta, const double* __restrict__ gamma, int64_t ldgamma,
                                                                    double* __restrict__ colsum) {
    __shared__ double part[SUM_WARPS][MAX_N];
    const int which = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* M = which ? gamma : beta;
    const int64_t rows = which ? Pc : Pb, ld = which ? ldgamma : ldbeta;
    double acc[MAX_N];
#pragma unroll
    for (int c = 0; c < MAX_N; ++c) acc[c] = 0.0;
    for (int64_t r = tid; r < rows; r += SUM_THREADS) {
        const double* row = M + r * ld;
#pragma unroll
        for (int c = 0; c < MAX_N; ++c)
            if (c < n) acc[c] += row[c];
    }
#pragma unroll
    for (int c = 0; c < MAX_N; ++c) {
        if (c < n) {
            double v = acc[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[warp][c] = v;
        }
    }
    __syncthreads();
    if (tid < n) {
        double v = 0.0;
        for (int w = 0; w < SUM_WARPS; ++w) v += part[w][tid];
        colsum[which * MAX_N + tid] = v;
    }
}

// block b: rows [a_begin + b*rows_per_block, ...) of W; thread <-> flat index rs (coalesced).  wpart[b] = its share of the sum.
__global__ void __launch_bounds__(SUM_THREADS) trimer_wsum_kernel(int n, const double* __restrict__ W, int64_t ldw, int64_t a_begin,
                                                                  int64_t a_end, int64_t rows_per_block,
                                                                  const double* __restrict__ colsum, double* __restrict__ wpart) {
    __shared__ double cs[2][MAX_N];
    __shared__ double red[SUM_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2 * MAX_N) cs[tid / MAX_N][tid % MAX_N] = (tid % MAX_N) < n ? colsum[tid] : 0.0;
    __syncthreads();
    const int64_t r0 = a_begin + (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < a_end ? r0 + rows_per_block : a_end;
    double d = 0.0;
    for (int e = tid; e < n * n; e += SUM_THREADS) {
        const double o = cs[0][e / n] * cs[1][e % n];
        double colw = 0.0;
        for (int64_t a = r0; a < r1; ++a) colw += W[a * ldw + e];
        d = fma(colw, o, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) red[warp] = d;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < SUM_WARPS; ++w) v += red[w];
        wpart[blockIdx.x] = v;
    }
}

// moments += (alpha * (sum of streamed partial sums + sum of wpart), alpha^2 * sum of streamed partial squares)
__global__ void trimer_finalize_kernel(const double* partials, int count, const double* wpart, int wcount, double alpha,
                                       double* moments) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0, tw = 0.0;
        for (int i = 0; i < count; ++i) {   // fixed order: bit-reproducible
            t1 += partials[2 * i];
            t2 += partials[2 * i + 1];
        }
        for (int i = 0; i < wcount; ++i) tw += wpart[i];
        moments[0] += alpha * (t1 + tw);
        moments[1] += alpha * alpha * t2;
    }
}

// host-side description of a sampled-element request: (a, b, c) triples, bucketed by work item in launch_trimer
struct SampleRequest {
    int64_t count = 0;
    const int64_t* abc = nullptr;      // HOST [count][3]
};

template <int KS, int TAIL, bool EXTRA = false, int WN = 2>
int launch_trimer(xr_ctx* ctx, TrimerParams p, const double* beta, int64_t ldbeta, const double* gamma, int64_t ldgamma,
                  double* moments, const SampleRequest* request = nullptr) {
    using Cfg = TrimerCfg<KS, TAIL, WN>;
    constexpr int CT = Cfg::CT;
    const int64_t n_a = p.a_end - p.a_begin;
    const int64_t tiles_a = (n_a + TA - 1) / TA;
    p.tiles_b = (p.Pb + TB - 1) / TB;
    p.n_items = tiles_a * p.tiles_b;
    p.c_tiles = (int)((p.Pc + CT - 1) / CT);
    // SAMPLE: only the work items that hold a requested element are streamed.  Samples are sorted by (item, c) on the host.
    std::vector<int64_t> item_list;
    std::vector<int32_t> sample_ptr;
    std::vector<int4> records;
    if (EXTRA && request) {
        std::vector<std::pair<int64_t, int64_t>> order(request->count);      // (item id, sample index)
        for (int64_t s = 0; s < request->count; ++s) {
            const int64_t a = request->abc[3 * s] - p.a_begin, b = request->abc[3 * s + 1];
            order[s] = {(a / TA) * p.tiles_b + b / TB, s};
        }
        std::sort(order.begin(), order.end());
        for (int64_t k = 0; k < request->count; ++k) {
            const int64_t s = order[k].second;
            if (k == 0 || order[k].first != order[k - 1].first) {
                item_list.push_back(order[k].first);
                sample_ptr.push_back((int32_t)k);
            }
            const int64_t a = request->abc[3 * s] - p.a_begin, b = request->abc[3 * s + 1], c = request->abc[3 * s + 2];
            records.push_back(make_int4((int)((a % TA) * TB + b % TB), (int)c, (int)s, 0));
        }
        sample_ptr.push_back((int32_t)request->count);
        p.n_items = (int64_t)item_list.size();
    }
    int grid = (int)(p.n_items < ctx->sm_count ? p.n_items : ctx->sm_count);

    // scratch: packed beta | packed gamma | partials
    const size_t beta_bytes = (size_t)p.Pb * Cfg::KP * sizeof(double);
    const size_t gamma_bytes = (size_t)p.c_tiles * CT * Cfg::GS * sizeof(double);
    const size_t beta_off = 0, gamma_off = (beta_bytes + 255) / 256 * 256;
    const size_t tail_bytes = (size_t)p.c_tiles * CT * 2 * sizeof(double);
    const size_t tail_off = gamma_off + (gamma_bytes + 255) / 256 * 256;
    const size_t part_off = tail_off + (tail_bytes + 255) / 256 * 256;
    // first-moment scratch (reduce mode): column sums of beta and gamma, then one partial per trimer_wsum_kernel block
    const int64_t wblocks_cap = (int64_t)ctx->sm_count * 4;
    const int64_t rows_per_block = (n_a + wblocks_cap - 1) / wblocks_cap;
    const int wblocks = (int)((n_a + rows_per_block - 1) / rows_per_block);
    const size_t colsum_off = part_off + ((size_t)grid * 2 * sizeof(double) + 255) / 256 * 256;
    const size_t wpart_off = colsum_off + (2 * MAX_N * sizeof(double) + 255) / 256 * 256;
    const size_t items_off = wpart_off + ((size_t)wblocks * sizeof(double) + 255) / 256 * 256;
    const size_t ptr_off = items_off + (item_list.size() * sizeof(int64_t) + 255) / 256 * 256;
    const size_t rec_off = ptr_off + (sample_ptr.size() * sizeof(int32_t) + 255) / 256 * 256;
    int rc = xr_ensure_scratch(ctx, rec_off + records.size() * sizeof(int4) + 256);
    if (rc != XR_OK) return rc;
    char* base = static_cast<char*>(ctx->scratch);
    double* betaP = reinterpret_cast<double*>(base + beta_off);
    double* gammaP = reinterpret_cast<double*>(base + gamma_off);
    double* gammaT = reinterpret_cast<double*>(base + tail_off);
    double* partials = reinterpret_cast<double*>(base + part_off);
    double* colsum = reinterpret_cast<double*>(base + colsum_off);
    double* wpart = reinterpret_cast<double*>(base + wpart_off);
    XR_CUDA(cudaMemsetAsync(base, 0, part_off, ctx->stream));
    rc = xr_copy2d_scaled(ctx, betaP, Cfg::KP, beta, ldbeta, p.Pb, p.n, 1.0);
    if (rc != XR_OK) return rc;
    rc = xr_copy2d_scaled(ctx, gammaP, Cfg::GS, gamma, ldgamma, p.Pc, p.n, 1.0);
    if (rc != XR_OK) return rc;
    if (TAIL) {
        rc = xr_copy2d_scaled(ctx, gammaT, 2, gamma + 4 * KS, ldgamma, p.Pc, TAIL, 1.0);
        if (rc != XR_OK) return rc;
    }
    p.betaP = betaP;
    p.gammaP = gammaP;
    p.gammaT = gammaT;
    p.partials = partials;
    // bulk TMA needs 16-byte aligned sources: every W row is, when the base is and ldw is even (build_H pads ldw to even)
    p.staged = Cfg::STAGE && (reinterpret_cast<uintptr_t>(p.W) % 16 == 0) && (p.ldw % 2 == 0);

    if (EXTRA && request) {
        // pageable sources: the runtime stages them before cudaMemcpyAsync returns, so the vectors may go out of scope
        XR_CUDA(cudaMemcpyAsync(base + items_off, item_list.data(), item_list.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        XR_CUDA(cudaMemcpyAsync(base + ptr_off, sample_ptr.data(), sample_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        XR_CUDA(cudaMemcpyAsync(base + rec_off, records.data(), records.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
        p.item_list = reinterpret_cast<const int64_t*>(base + items_off);
        p.sample_ptr = reinterpret_cast<const int32_t*>(base + ptr_off);
        p.samples = reinterpret_cast<const int4*>(base + rec_off);
    }

    auto kernel = trimer_stream_kernel<KS, TAIL, WN, EXTRA>;
    XR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    kernel<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(p);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    if (p.mode == XR_TRIMER_REDUCE) {
        if (!STREAM_SUM) {
            trimer_colsum_kernel<<<2, SUM_THREADS, 0, ctx->stream>>>(p.n, p.Pb, p.Pc, beta, ldbeta, gamma, ldgamma, colsum);
            XR_CUDA(cudaGetLastError());
            trimer_wsum_kernel<<<wblocks, SUM_THREADS, 0, ctx->stream>>>(p.n, p.W, p.ldw, p.a_begin, p.a_end, rows_per_block, colsum,
                                                                         wpart);
            XR_CUDA(cudaGetLastError());
            ctx->launches += 2;
        }
        trimer_finalize_kernel<<<1, 32, 0, ctx->stream>>>(partials, grid, wpart, STREAM_SUM ? 0 : wblocks, p.alpha, moments);
        XR_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    return XR_OK;
}

// k = n is covered by 4*KS DMMA k-steps + TAIL DFMA k: exactly for n <= 20 (n mod 4 = 3 rounds up to the next DMMA
// step), in steps of 4 beyond (A fragments then come from shared memory, no tail).
// (WN = 3, i.e. 12 consumer warps with 32x32 warp tiles, was measured too: 29.1 vs 29.9 TFLOP/s for WN = 2 at
//  n = 18 on B200 -- the kernel is bound by its DMMA:DFMA instruction mix, not by warp-level latency hiding)
template <bool EXTRA>
int dispatch_trimer(xr_ctx* ctx, const TrimerParams& p, int n, const double* beta, int64_t ldbeta, const double* gamma,
                    int64_t ldgamma, double* moments, const SampleRequest* request) {
#define XR_TRIMER_CASE(COND, KS, TAIL) \
    if (COND) return launch_trimer<KS, TAIL, EXTRA>(ctx, p, beta, ldbeta, gamma, ldgamma, moments, request)
    XR_TRIMER_CASE(n <= 4, 1, 0);
    XR_TRIMER_CASE(n == 5, 1, 1);
    XR_TRIMER_CASE(n == 6, 1, 2);
    XR_TRIMER_CASE(n <= 8, 2, 0);
    XR_TRIMER_CASE(n == 9, 2, 1);
    XR_TRIMER_CASE(n == 10, 2, 2);
    XR_TRIMER_CASE(n <= 12, 3, 0);
    XR_TRIMER_CASE(n == 13, 3, 1);
    XR_TRIMER_CASE(n == 14, 3, 2);
    XR_TRIMER_CASE(n <= 16, 4, 0);
    XR_TRIMER_CASE(n == 17, 4, 1);
    XR_TRIMER_CASE(n == 18, 4, 2);
    XR_TRIMER_CASE(n <= 20, 5, 0);
    XR_TRIMER_CASE(n <= 24, 6, 0);
    XR_TRIMER_CASE(n <= 28, 7, 0);
    XR_TRIMER_CASE(n <= 32, 8, 0);
    XR_TRIMER_CASE(n <= 36, 9, 0);
    XR_TRIMER_CASE(n <= 40, 10, 0);
    XR_TRIMER_CASE(n <= 44, 11, 0);
#undef XR_TRIMER_CASE
    return launch_trimer<12, 0, EXTRA>(ctx, p, beta, ldbeta, gamma, ldgamma, moments, request);
}

}  // namespace

extern "C" int xr_trimer_stream(xr_ctx* ctx, int n, int64_t Pa, int64_t Pb, int64_t Pc, double alpha, const double* W,
                                int64_t ldw, const double* beta, int64_t ldbeta, const double* gamma, int64_t ldgamma,
                                int64_t a_begin, int64_t a_end, int mode, double* moments, double* C, const int64_t* offA,
                                const int64_t* offB, const int64_t* offC) {
    XR_REQUIRE(ctx, "xr_trimer_stream: null ctx");
    XR_REQUIRE(n >= 1 && n <= 48, "xr_trimer_stream: n=%d unsupported (1..48)", n);
    XR_REQUIRE(a_begin >= 0 && a_end <= Pa && a_begin <= a_end, "xr_trimer_stream: bad a range [%lld,%lld) of %lld",
               (long long)a_begin, (long long)a_end, (long long)Pa);
    if (a_begin == a_end || Pb <= 0 || Pc <= 0) return XR_OK;
    XR_REQUIRE(W && beta && gamma, "xr_trimer_stream: null operand");
    XR_REQUIRE(ldw >= (int64_t)n * n && ldbeta >= n && ldgamma >= n, "xr_trimer_stream: leading dimension too small");
    if (mode == XR_TRIMER_REDUCE) {
        XR_REQUIRE(moments, "xr_trimer_stream: reduce mode needs moments");
    } else if (mode == XR_TRIMER_MATERIALIZE) {
        XR_REQUIRE(C && offA && offB && offC, "xr_trimer_stream: materialize mode needs C and offset tables");
    } else {
        XR_REQUIRE(false, "xr_trimer_stream: unknown mode %d", mode);
    }
    TrimerParams p{};
    p.n = n;
    p.Pb = Pb;
    p.Pc = Pc;
    p.alpha = alpha;
    p.W = W;
    p.ldw = ldw;
    p.a_begin = a_begin;
    p.a_end = a_end;
    p.mode = mode;
    p.C = C;
    p.offA = offA;
    p.offB = offB;
    p.offC = offC;
    return dispatch_trimer<false>(ctx, p, n, beta, ldbeta, gamma, ldgamma, moments, nullptr);
}

/* The compaction consumer: every element with |alpha * T[a,b,c]| > tau is appended to (idx_out, val_out) as
 * (offA[a] + offB[b] + offC[c], value); *count (device) receives the number of such elements, which may exceed
 * `capacity` -- then only the first `capacity` reservations were stored and the caller re-runs with a larger list. */
extern "C" int xr_trimer_threshold(xr_ctx* ctx, int n, int64_t Pa, int64_t Pb, int64_t Pc, double alpha, const double* W,
                                   int64_t ldw, const double* beta, int64_t ldbeta, const double* gamma, int64_t ldgamma,
                                   int64_t a_begin, int64_t a_end, double tau, const int64_t* offA, const int64_t* offB,
                                   const int64_t* offC, int64_t capacity, int64_t* idx_out, double* val_out, int64_t* count) {
    XR_REQUIRE(ctx, "xr_trimer_threshold: null ctx");
    XR_REQUIRE(count, "xr_trimer_threshold: null count");
    XR_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t), ctx->stream));
    XR_REQUIRE(n >= 1 && n <= 48, "xr_trimer_threshold: n=%d unsupported (1..48)", n);
    XR_REQUIRE(a_begin >= 0 && a_end <= Pa && a_begin <= a_end, "xr_trimer_threshold: bad a range [%lld,%lld) of %lld",
               (long long)a_begin, (long long)a_end, (long long)Pa);
    if (a_begin == a_end || Pb <= 0 || Pc <= 0) return XR_OK;
    XR_REQUIRE(W && beta && gamma && offA && offB && offC, "xr_trimer_threshold: null operand or offset table");
    XR_REQUIRE(capacity >= 0 && (capacity == 0 || (idx_out && val_out)), "xr_trimer_threshold: output list missing");
    XR_REQUIRE(tau >= 0.0, "xr_trimer_threshold: tau must be >= 0");
    XR_REQUIRE(ldw >= (int64_t)n * n && ldbeta >= n && ldgamma >= n, "xr_trimer_threshold: leading dimension too small");
    TrimerParams p{};
    p.n = n;
    p.Pb = Pb;
    p.Pc = Pc;
    p.alpha = alpha;
    p.W = W;
    p.ldw = ldw;
    p.a_begin = a_begin;
    p.a_end = a_end;
    p.mode = XR_TRIMER_THRESHOLD;
    p.offA = offA;
    p.offB = offB;
    p.offC = offC;
    p.tau = tau;
    p.counter = reinterpret_cast<unsigned long long*>(count);
    p.capacity = capacity;
    p.idx_out = idx_out;
    p.val_out = val_out;
    return dispatch_trimer<true>(ctx, p, n, beta, ldbeta, gamma, ldgamma, nullptr, nullptr);
}

/* The sampled-element consumer: out[t] = alpha * T[abc[3t], abc[3t+1], abc[3t+2]] for a caller-given HOST list of triples;
 * only the work items that hold a sample are streamed (through the same tile code as every other consumer). */
extern "C" int xr_trimer_sample(xr_ctx* ctx, int n, int64_t Pa, int64_t Pb, int64_t Pc, double alpha, const double* W, int64_t ldw,
                                const double* beta, int64_t ldbeta, const double* gamma, int64_t ldgamma, int64_t count,
                                const int64_t* abc_host, double* out) {
    XR_REQUIRE(ctx, "xr_trimer_sample: null ctx");
    if (count <= 0) return XR_OK;
    XR_REQUIRE(n >= 1 && n <= 48, "xr_trimer_sample: n=%d unsupported (1..48)", n);
    XR_REQUIRE(W && beta && gamma && abc_host && out, "xr_trimer_sample: null argument");
    XR_REQUIRE(count < (1ll << 31), "xr_trimer_sample: too many samples");
    XR_REQUIRE(ldw >= (int64_t)n * n && ldbeta >= n && ldgamma >= n, "xr_trimer_sample: leading dimension too small");
    for (int64_t s = 0; s < count; ++s) {
        const int64_t a = abc_host[3 * s], b = abc_host[3 * s + 1], c = abc_host[3 * s + 2];
        XR_REQUIRE(a >= 0 && a < Pa && b >= 0 && b < Pb && c >= 0 && c < Pc, "xr_trimer_sample: sample %lld = (%lld,%lld,%lld) out of range",
                   (long long)s, (long long)a, (long long)b, (long long)c);
    }
    TrimerParams p{};
    p.n = n;
    p.Pb = Pb;
    p.Pc = Pc;
    p.alpha = alpha;
    p.W = W;
    p.ldw = ldw;
    p.a_begin = 0;
    p.a_end = Pa;
    p.mode = XR_TRIMER_SAMPLE;
    p.sample_out = out;
    SampleRequest request;
    request.count = count;
    request.abc = abc_host;
    return dispatch_trimer<true>(ctx, p, n, beta, ldbeta, gamma, ldgamma, nullptr, &request);
}
