// Transition-density tensors from CI vectors -- the step before the H build (general-XRCC/density_tensors.c:142-556,
// driven by build_density_tensors.py:70-157):
//
//   rho[I,J,(i_0..i_{k-1})] = sum_Q  parity * z_bra[I, P] * z_ket[J, Q],   |P> = +- op_0(i_0) ... op_{k-1}(i_{k-1}) |Q>
//
// The reference walks ket configurations and scatters.  Here: phase 1 (integer/bit work) builds, per tensor index, the list
// of couplings (P, Q, parity) in the reference's ket-configuration order (count, exclusive scan, fill -- configurations are
// 64-bit occupation masks, the rank of the bra configuration (find_config_index, density_tensors.c:29-64) comes from a
// prefix table of binomials); phase 2 (FP64) lets every thread OWN output elements (one tensor index, one bra state, a tile
// of ket states) and sum its list in order -- no atomics, and the floating-point summation order and rounding (and so every
// bit of the result) are the reference's.  HBM-write bound on the output (N_bra N_ket dim^k doubles).
#include "xr_common.cuh"
#include <cmath>
#include <cub/device/device_scan.cuh>
#include <mutex>
#include <vector>

namespace {

constexpr int IT = 4, JT = 8;  // bra x ket states per thread (register accumulators)
constexpr int MAX_OPS = 4;

struct DensityParams {
    int k;                     // number of field operators
    int create[MAX_OPS];       // 1 = creation, 0 = annihilation, in string order (applied right to left)
    int64_t dim, T;            // spin orbitals, dim^k
    int64_t n_bra, n_ket, ncfg_bra, ncfg_ket;
    int n_orbs, n_core, n_val_elec_bra, S;
    unsigned long long core_mask;
    const double* z_bra;
    const double* z_ket;
    const unsigned long long* ket_masks;
    const long long* G;        // [n_val_elec_bra][S+1]: G[i][m] = sum_{n=1..m} C(S-n, n_val_elec_bra-i-1)
    double* rho;
};

__device__ __forceinline__ long long config_rank(unsigned long long mask, const DensityParams& p) {
    const int nv = p.n_orbs - p.n_core;                                  // valence orbitals per spin
    const unsigned long long lo = (mask >> p.n_core) & ((1ull << nv) - 1ull);
    const unsigned long long hi = (mask >> (p.n_orbs + p.n_core)) & ((1ull << nv) - 1ull);
    unsigned long long val = lo | (hi << nv);
    long long index = 0;
    int prev = -1;
    for (int i = 0; i < p.n_val_elec_bra; ++i) {
        const int c = __ffsll((long long)val) - 1;
        val &= val - 1ull;
        const long long* g = p.G + (size_t)i * (p.S + 1);
        index += g[c] - g[prev + 1];
        prev = c;
    }
    return index;
}

// |P> = +- op_0(orb_0) ... op_{k-1}(orb_{k-1}) |mask>, operators applied right to left.  Returns false when a Pauli or
// frozen-core violation kills the term; else the new mask and the number of transpositions (density_tensors.c:80-113).
__device__ __forceinline__ bool apply_string(const DensityParams& p, const int* orb, unsigned long long& mask, int& flips) {
    flips = 0;
    for (int o = p.k - 1; o >= 0; --o) {
        const unsigned long long bit = 1ull << orb[o];
        const int below = __popcll(mask & (bit - 1ull));
        const int n = __popcll(mask);
        if (p.create[o]) {
            if (mask & bit) return false;
            flips += n - below;                 // density_tensors.c:104-109: shifts to insert at position `below`
            mask |= bit;
        } else {
            if (!(mask & bit)) return false;
            flips += n - 1 - below;             // density_tensors.c:88-92: shifts to move it to the end
            mask &= ~bit;
        }
    }
    return (mask & p.core_mask) == p.core_mask;
}

__device__ __forceinline__ void decode_index(const DensityParams& p, int64_t idx, int* orb) {
    for (int o = p.k - 1; o >= 0; --o) {
        orb[o] = (int)(idx % p.dim);
        idx /= p.dim;
    }
}

// Phase 1 (integer): the coupling list of every tensor index, in ket-configuration order.  One thread per (tensor index,
// chunk of 64 ket configurations): the first pass records WHICH configurations couple (a 64-bit hit mask, its popcount is the
// chunk's entry count); after an exclusive scan of the counts in (index, chunk) order the second pass revisits only the
// hits and writes entries {P | sign << 31, Q}.  Lanes are consecutive indices, so ket_masks[Q] is a broadcast load.
constexpr int QCHUNK = 64;

__global__ void __launch_bounds__(256) density_hits_kernel(const DensityParams p, int64_t chunks, long long* __restrict__ counts,
                                                           unsigned long long* __restrict__ hitmasks) {
    const int64_t total = p.T * chunks;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t idx = t % p.T, chunk = t / p.T;
        int orb[MAX_OPS];
        decode_index(p, idx, orb);
        const int64_t q0 = chunk * QCHUNK;
        const int nq = (int)(p.ncfg_ket - q0 < QCHUNK ? p.ncfg_ket - q0 : QCHUNK);
        unsigned long long hits = 0ull;
        for (int q = 0; q < nq; ++q) {
            unsigned long long mask = p.ket_masks[q0 + q];
            int flips;
            if (apply_string(p, orb, mask, flips)) hits |= 1ull << q;
        }
        counts[idx * chunks + chunk] = __popcll(hits);
        hitmasks[idx * chunks + chunk] = hits;
    }
}

__global__ void __launch_bounds__(256) density_fill_kernel(const DensityParams p, int64_t chunks, const long long* __restrict__ offsets,
                                                           const unsigned long long* __restrict__ hitmasks, uint2* __restrict__ entries,
                                                           const double* __restrict__ weights, double* __restrict__ entry_weights) {
    const int64_t total = p.T * chunks;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t idx = t % p.T, chunk = t / p.T;
        unsigned long long hits = hitmasks[idx * chunks + chunk];
        if (!hits) continue;
        int orb[MAX_OPS];
        decode_index(p, idx, orb);
        long long at = offsets[idx * chunks + chunk];
        while (hits) {
            const int q = __ffsll((long long)hits) - 1;
            hits &= hits - 1ull;
            const int64_t Q = chunk * QCHUNK + q;
            unsigned long long mask = p.ket_masks[Q];
            int flips;
            apply_string(p, orb, mask, flips);
            entries[at] = make_uint2((unsigned)config_rank(mask, p) | ((unsigned)(flips & 1) << 31), (unsigned)Q);
            if (entry_weights) entry_weights[at] = (flips & 1) ? -weights[idx] : weights[idx];       // parity * weights[index]
            ++at;
        }
    }
}

// CI vectors transposed and zero-padded: zT[cfg * n_pad + state], so that the IT (JT) states a thread needs for one
// coupling are one (two) 32-byte sectors instead of IT (JT) scattered ones.
__global__ void density_transpose_kernel(const double* __restrict__ z, int64_t n_states, int64_t n_cfg, int64_t n_pad, double* __restrict__ zT) {
    const int64_t total = n_cfg * n_pad;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / n_pad, s = t % n_pad;
        zT[t] = s < n_states ? z[s * n_cfg + c] : 0.0;
    }
}

// acc[i][j] += sum over the list entries [e0, e1) of (parity * z_bra[I0+i, P]) * z_ket[J0+j, Q], in list order
__device__ __forceinline__ void sum_list(double (&acc)[IT][JT], long long e0, long long e1, const uint2* __restrict__ entries,
                                         const double* __restrict__ zT_bra, const double* __restrict__ zT_ket, int64_t nb_pad,
                                         int64_t nk_pad, int64_t I0, int64_t J0) {
    for (long long e = e0; e < e1; ++e) {
        const uint2 entry = entries[e];
        const double4 b4 = *reinterpret_cast<const double4*>(zT_bra + (size_t)(entry.x & 0x7fffffffu) * nb_pad + I0);
        const double4 k0 = *reinterpret_cast<const double4*>(zT_ket + (size_t)entry.y * nk_pad + J0);
        const double4 k1 = *reinterpret_cast<const double4*>(zT_ket + (size_t)entry.y * nk_pad + J0 + 4);
        const bool neg = entry.x >> 31;
        const double left[IT] = {neg ? -b4.x : b4.x, neg ? -b4.y : b4.y, neg ? -b4.z : b4.z, neg ? -b4.w : b4.w};   // parity * zI_P (exact)
        const double zk[JT] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
        for (int i = 0; i < IT; ++i)
#pragma unroll
            for (int j = 0; j < JT; ++j)                                     // (parity*zI_P)*zJ_Q rounded, then added: no FMA, as the C
                acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(left[i], zk[j]));
    }
}

// Phase 2 (FP64): every thread OWNS one tensor index, IT bra states and JT ket states, and sums its list in order.
__global__ void __launch_bounds__(256) density_apply_kernel(const DensityParams p, const long long* __restrict__ offsets,
                                                            const uint2* __restrict__ entries, const double* __restrict__ zT_bra,
                                                            const double* __restrict__ zT_ket, int64_t nb_pad, int64_t nk_pad, int64_t chunks,
                                                            int accumulate) {
    const int64_t itiles = nb_pad / IT, jtiles = nk_pad / JT;
    const int64_t total = p.T * itiles * jtiles;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t idx = t % p.T;
        const int64_t rest = t / p.T;
        const int64_t I0 = (rest % itiles) * IT, J0 = (rest / itiles) * JT;
        double acc[IT][JT];
#pragma unroll
        for (int i = 0; i < IT; ++i)
#pragma unroll
            for (int j = 0; j < JT; ++j)
                acc[i][j] = (accumulate && I0 + i < p.n_bra && J0 + j < p.n_ket) ? p.rho[((I0 + i) * p.n_ket + J0 + j) * p.T + idx] : 0.0;
        sum_list(acc, offsets[idx * chunks], offsets[(idx + 1) * chunks], entries, zT_bra, zT_ket, nb_pad, nk_pad, I0, J0);
#pragma unroll
        for (int i = 0; i < IT; ++i)
#pragma unroll
            for (int j = 0; j < JT; ++j)
                if (I0 + i < p.n_bra && J0 + j < p.n_ket) p.rho[((I0 + i) * p.n_ket + J0 + j) * p.T + idx] = acc[i][j];
    }
}

// Phase 2, contracted: out[I,J] (+)= sum_index weights[index] * rho[I,J,index] without ever storing rho (the reference forms
// the ccaa tensor only to reduce it with V at once, build_density_tensors.py:125-133).  No per-index structure is needed any
// more: out = sum over ALL couplings e of (parity_e weights[index_e]) z_bra[:,P_e] (x) z_ket[:,Q_e] -- a flat, perfectly
// balanced walk of the coupling list with coalesced entry loads.  blockIdx.y = (bra tile, ket tile); every thread folds
// its couplings in order, then a fixed shuffle/shared-memory tree and a fixed-order second pass over the blocks:
// bit-reproducible.
constexpr int CONTRACT_BLOCKS = 512;

__global__ void __launch_bounds__(256) density_contract_kernel(const DensityParams p, long long nnz, const uint2* __restrict__ entries,
                                                               const double* __restrict__ entry_weights, const double* __restrict__ zT_bra,
                                                               const double* __restrict__ zT_ket, int64_t nb_pad, int64_t nk_pad,
                                                               double* __restrict__ partials) {
    const int64_t itiles = nb_pad / IT;
    const int64_t I0 = (blockIdx.y % itiles) * IT, J0 = (blockIdx.y / itiles) * JT;
    double sum[IT][JT];
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int j = 0; j < JT; ++j) sum[i][j] = 0.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
        const uint2 entry = entries[e];
        const double w = entry_weights[e];
        const double4 b4 = *reinterpret_cast<const double4*>(zT_bra + (size_t)(entry.x & 0x7fffffffu) * nb_pad + I0);
        const double4 k0 = *reinterpret_cast<const double4*>(zT_ket + (size_t)entry.y * nk_pad + J0);
        const double4 k1 = *reinterpret_cast<const double4*>(zT_ket + (size_t)entry.y * nk_pad + J0 + 4);
        const double left[IT] = {w * b4.x, w * b4.y, w * b4.z, w * b4.w};
        const double zk[JT] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
        for (int i = 0; i < IT; ++i)
#pragma unroll
            for (int j = 0; j < JT; ++j) sum[i][j] = fma(left[i], zk[j], sum[i][j]);
    }
    __shared__ double red[8][IT * JT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int j = 0; j < JT; ++j) {
            double v = sum[i][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][i * JT + j] = v;
        }
    __syncthreads();
    if (threadIdx.x < IT * JT) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (IT * JT) + threadIdx.x] = v;
    }
}

__global__ void density_contract_finish_kernel(const DensityParams p, const double* __restrict__ partials, int64_t tiles, int blocks_x,
                                               int64_t nb_pad, double* __restrict__ out, int accumulate) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= tiles * IT * JT) return;
    const int64_t tile = t / (IT * JT);
    const int v = (int)(t % (IT * JT));
    const int64_t itiles = nb_pad / IT;
    const int64_t I = (tile % itiles) * IT + v / JT, J = (tile / itiles) * JT + v % JT;
    if (I >= p.n_bra || J >= p.n_ket) return;
    double s = 0.0;
    for (int b = 0; b < blocks_x; ++b) s += partials[((size_t)tile * blocks_x + b) * (IT * JT) + v];
    out[I * p.n_ket + J] = accumulate ? out[I * p.n_ket + J] + s : s;
}

long long binomial(int n, int k) {
    if (k < 0 || k > n) return 0;
    long long r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

}  // namespace

namespace {

// Temporaries from the stream-ordered allocator, returned to the pool (in stream order) on every exit path.
struct StreamBuffers {
    cudaStream_t stream;
    std::vector<void*> owned;
    explicit StreamBuffers(cudaStream_t s) : stream(s) {}
    ~StreamBuffers() {
        for (size_t i = owned.size(); i-- > 0;) cudaFreeAsync(owned[i], stream);
    }
    template <typename T>
    cudaError_t get(T** ptr, size_t count) {
        cudaError_t rc = cudaMallocAsync(reinterpret_cast<void**>(ptr), (count ? count : 1) * sizeof(T), stream);
        if (rc == cudaSuccess) owned.push_back(*ptr);
        return rc;
    }
};

// rho != nullptr: the tensor itself; weights/out != nullptr: its contraction with weights (never stored)
int density_run(xr_ctx* ctx, const char* what, const char* ops, double* rho, const double* weights, double* out, int64_t n_bra_states,
                int64_t n_ket_states, const double* z_bra, int64_t n_configs_bra, const double* z_ket, int64_t n_configs_ket,
                const uint64_t* ket_masks, int64_t n_elec_bra, int64_t n_elec_ket, int64_t n_orbs, int64_t n_core, int accumulate) {
    XR_REQUIRE(ctx && ops, "%s: null ctx or operator string", what);
    DensityParams p{};
    int dchg = 0;
    for (p.k = 0; ops[p.k]; ++p.k) {
        XR_REQUIRE(p.k < MAX_OPS, "%s: more than %d operators in '%s'", what, MAX_OPS, ops);
        XR_REQUIRE(ops[p.k] == 'c' || ops[p.k] == 'a', "%s: operator string '%s' must consist of c and a", what, ops);
        p.create[p.k] = ops[p.k] == 'c';
        dchg += p.create[p.k] ? 1 : -1;
    }
    XR_REQUIRE(p.k >= 1, "%s: empty operator string", what);
    XR_REQUIRE(n_orbs >= 1 && 2 * n_orbs <= 64 && n_core >= 0 && n_core <= n_orbs, "%s: need 1 <= 2*n_orbs <= 64, 0 <= n_core <= n_orbs", what);
    XR_REQUIRE(n_elec_ket + dchg == n_elec_bra, "%s: '%s' does not connect %lld to %lld electrons", what, ops, (long long)n_elec_ket,
               (long long)n_elec_bra);
    XR_REQUIRE(n_elec_bra >= 2 * n_core && n_elec_bra <= 2 * n_orbs, "%s: bra electron count out of range", what);
    if (n_bra_states <= 0 || n_ket_states <= 0 || n_configs_ket <= 0) return XR_OK;
    XR_REQUIRE((rho || (weights && out)) && z_bra && z_ket && ket_masks, "%s: null pointer", what);
    p.dim = 2 * n_orbs;
    p.T = 1;
    for (int o = 0; o < p.k; ++o) p.T *= p.dim;
    p.n_bra = n_bra_states; p.n_ket = n_ket_states; p.ncfg_bra = n_configs_bra; p.ncfg_ket = n_configs_ket;
    p.n_orbs = (int)n_orbs; p.n_core = (int)n_core;
    p.n_val_elec_bra = (int)(n_elec_bra - 2 * n_core);
    p.S = (int)(2 * (n_orbs - n_core));
    XR_REQUIRE(binomial(p.S, p.n_val_elec_bra) == n_configs_bra, "%s: n_configs_bra=%lld is not C(%d,%d): the bra coefficients must "
               "span every valence configuration in find_config_index order", what, (long long)n_configs_bra, p.S, p.n_val_elec_bra);
    XR_REQUIRE(n_configs_bra < (1ll << 31) && n_configs_ket < (1ll << 31), "%s: more than 2^31 configurations", what);
    p.core_mask = 0;
    for (int i = 0; i < n_core; ++i) p.core_mask |= (1ull << i) | (1ull << (n_orbs + i));
    p.z_bra = z_bra; p.z_ket = z_ket; p.ket_masks = reinterpret_cast<const unsigned long long*>(ket_masks); p.rho = rho;
    const int64_t chunks = (p.ncfg_ket + QCHUNK - 1) / QCHUNK;
    const int64_t cells = p.T * chunks;
    XR_REQUIRE(cells + 1 < (1ll << 31), "%s: tensor too large (%lld indices x %lld configuration chunks)", what, (long long)p.T, (long long)chunks);

    XR_CUDA(cudaSetDevice(ctx->device));
    {   // keep freed temporaries in the pool across the synchronisations below
        cudaMemPool_t pool;
        XR_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        unsigned long long keep = ~0ull;
        XR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    StreamBuffers tmp(ctx->stream);
    // prefix table of the ranking binomials (density_tensors.c:34-45): G[i][m] = sum_{n=1..m} C(S-n, e-i-1)
    std::vector<long long> G((size_t)(p.n_val_elec_bra > 0 ? p.n_val_elec_bra : 1) * (p.S + 1), 0);
    for (int i = 0; i < p.n_val_elec_bra; ++i)
        for (int m = 1; m <= p.S; ++m) G[(size_t)i * (p.S + 1) + m] = G[(size_t)i * (p.S + 1) + m - 1] + binomial(p.S - m, p.n_val_elec_bra - i - 1);
    long long* dG = nullptr;
    XR_CUDA(tmp.get(&dG, G.size()));
    XR_CUDA(cudaMemcpyAsync(dG, G.data(), G.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before return
    p.G = dG;

    // phase 1: hits -> exclusive scan -> fill
    const int64_t cap = (int64_t)ctx->sm_count * 64;
    int64_t blocks = (cells + 255) / 256;
    if (blocks > cap) blocks = cap;
    long long *counts = nullptr, *offsets = nullptr;
    unsigned long long* hitmasks = nullptr;
    void* scan_tmp = nullptr;
    size_t scan_bytes = 0;
    XR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts, offsets, (int)(cells + 1), ctx->stream));
    XR_CUDA(tmp.get(&counts, (size_t)cells + 1));
    XR_CUDA(tmp.get(&offsets, (size_t)cells + 1));
    XR_CUDA(tmp.get(&hitmasks, (size_t)cells));
    XR_CUDA(tmp.get(reinterpret_cast<unsigned char**>(&scan_tmp), scan_bytes));
    XR_CUDA(cudaMemsetAsync(counts + cells, 0, sizeof(long long), ctx->stream));
    density_hits_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, chunks, counts, hitmasks);
    XR_CUDA(cudaGetLastError());
    XR_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, offsets, (int)(cells + 1), ctx->stream));
    long long nnz = 0;
    XR_CUDA(cudaMemcpyAsync(&nnz, offsets + cells, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    XR_CUDA(cudaStreamSynchronize(ctx->stream));
    uint2* entries = nullptr;
    double* entry_weights = nullptr;
    XR_CUDA(tmp.get(&entries, (size_t)nnz));
    if (!rho) XR_CUDA(tmp.get(&entry_weights, (size_t)nnz));
    density_fill_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, chunks, offsets, hitmasks, entries, weights, entry_weights);
    XR_CUDA(cudaGetLastError());

    // phase 2
    const int64_t nb_pad = (p.n_bra + IT - 1) / IT * IT, nk_pad = (p.n_ket + JT - 1) / JT * JT;
    double *zT_bra = nullptr, *zT_ket = nullptr;
    XR_CUDA(tmp.get(&zT_bra, (size_t)(p.ncfg_bra * nb_pad)));
    XR_CUDA(tmp.get(&zT_ket, (size_t)(p.ncfg_ket * nk_pad)));
    density_transpose_kernel<<<(unsigned)((p.ncfg_bra * nb_pad + 255) / 256), 256, 0, ctx->stream>>>(z_bra, p.n_bra, p.ncfg_bra, nb_pad, zT_bra);
    density_transpose_kernel<<<(unsigned)((p.ncfg_ket * nk_pad + 255) / 256), 256, 0, ctx->stream>>>(z_ket, p.n_ket, p.ncfg_ket, nk_pad, zT_ket);
    XR_CUDA(cudaGetLastError());
    const int64_t tiles = (nb_pad / IT) * (nk_pad / JT);
    if (rho) {
        const int64_t total = p.T * tiles;
        blocks = (total + 255) / 256;
        if (blocks > cap) blocks = cap;
        density_apply_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, offsets, entries, zT_bra, zT_ket, nb_pad, nk_pad, chunks, accumulate);
        XR_CUDA(cudaGetLastError());
        ctx->launches += 5;
    } else {
        XR_REQUIRE(tiles < 65536, "%s: too many state tiles (%lld)", what, (long long)tiles);
        int bx = (int)((nnz + 255) / 256);
        if (bx > CONTRACT_BLOCKS) bx = CONTRACT_BLOCKS;
        if (bx < 1) bx = 1;
        double* partials = nullptr;
        XR_CUDA(tmp.get(&partials, (size_t)tiles * bx * IT * JT));
        density_contract_kernel<<<dim3((unsigned)bx, (unsigned)tiles), 256, 0, ctx->stream>>>(p, nnz, entries, entry_weights, zT_bra, zT_ket, nb_pad,
                                                                                             nk_pad, partials);
        XR_CUDA(cudaGetLastError());
        density_contract_finish_kernel<<<(unsigned)((tiles * IT * JT + 255) / 256), 256, 0, ctx->stream>>>(p, partials, tiles, bx, nb_pad, out, accumulate);
        XR_CUDA(cudaGetLastError());
        ctx->launches += 6;
    }
    return XR_OK;
}

}  // namespace

extern "C" int xr_density_tensor(xr_ctx* ctx, const char* ops, double* rho, int64_t n_bra_states, int64_t n_ket_states,
                                 const double* z_bra, int64_t n_configs_bra, const double* z_ket, int64_t n_configs_ket,
                                 const uint64_t* ket_masks, int64_t n_elec_bra, int64_t n_elec_ket, int64_t n_orbs, int64_t n_core,
                                 int accumulate) {
    XR_REQUIRE(rho || n_bra_states <= 0 || n_ket_states <= 0, "xr_density_tensor: null rho");
    return density_run(ctx, "xr_density_tensor", ops, rho, nullptr, nullptr, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket,
                       n_configs_ket, ket_masks, n_elec_bra, n_elec_ket, n_orbs, n_core, accumulate);
}

extern "C" int xr_density_contracted(xr_ctx* ctx, const char* ops, double* out, const double* weights, int64_t n_bra_states,
                                     int64_t n_ket_states, const double* z_bra, int64_t n_configs_bra, const double* z_ket,
                                     int64_t n_configs_ket, const uint64_t* ket_masks, int64_t n_elec_bra, int64_t n_elec_ket, int64_t n_orbs,
                                     int64_t n_core, int accumulate) {
    XR_REQUIRE((out && weights) || n_bra_states <= 0 || n_ket_states <= 0, "xr_density_contracted: null out or weights");
    return density_run(ctx, "xr_density_contracted", ops, nullptr, weights, out, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket,
                       n_configs_ket, ket_masks, n_elec_bra, n_elec_ket, n_orbs, n_core, accumulate);
}

// ------------------------------------------------------------------------------------------------------------------------
// The eight entry points of general-XRCC/density_tensors.c with the reference's own C signature (host pointers; PyInt =
// BigInt = int64, Double = double), so that build_density_tensors.py:23 `import_C("density_tensors")` can bind this
// library instead.  Each call uploads the two CI-vector blocks and the ket configurations (as occupation masks), runs
// density_kernel, and adds into `storage` exactly as the reference's `tensor[index] +=` does.  `combinatorics` is accepted
// for signature compatibility (the ranking table is rebuilt from n_orbs/n_core); n_threads is ignored, as in the reference.
namespace {

std::mutex g_density_mutex;
xr_ctx* g_density_ctx = nullptr;

// returns false (with xr_last_error set) on any failure
bool legacy_density_run(const char* ops, double* storage, int64_t bra, int64_t ket, const int64_t* n_elec, const int64_t* n_states,
                        double* const* z_list, const int64_t* n_configs, int64_t* const* configs, int64_t n_orbs, int64_t n_core,
                        int64_t* storage_elements) {
    std::lock_guard<std::mutex> lock(g_density_mutex);
    *storage_elements = 0;
    if (!storage || !n_elec || !n_states || !z_list || !n_configs || !configs) {
        xr_set_error("%s_tensor: null argument", ops);
        return false;
    }
    int k = 0;
    while (ops[k]) ++k;
    int64_t T = 1;
    for (int o = 0; o < k; ++o) T *= 2 * n_orbs;
    const int64_t nb = n_states[bra], nk = n_states[ket], cb = n_configs[bra], ck = n_configs[ket], ne = n_elec[ket];
    if (nb <= 0 || nk <= 0 || ck <= 0) return true;           // an empty sector: nothing to add
    *storage_elements = nb * nk * T;
    if (!g_density_ctx && xr_ctx_create(0, nullptr, 1, &g_density_ctx) != XR_OK) return false;
    xr_ctx* ctx = g_density_ctx;
    std::vector<unsigned long long> masks((size_t)ck, 0ull);
    for (int64_t Q = 0; Q < ck; ++Q)
        for (int64_t e = 0; e < ne; ++e) {
            const int64_t orb = configs[ket][Q * ne + e];
            if (orb < 0 || orb >= 64) {
                xr_set_error("%s_tensor: orbital index %lld outside 0..63", ops, (long long)orb);
                return false;
            }
            masks[(size_t)Q] |= 1ull << orb;
        }
    const size_t bytes_rho = (size_t)(nb * nk * T) * sizeof(double), bytes_zb = (size_t)(nb * cb) * sizeof(double),
                 bytes_zk = (size_t)(nk * ck) * sizeof(double), bytes_m = (size_t)ck * sizeof(unsigned long long);
    double *d_rho = nullptr, *d_zb = nullptr, *d_zk = nullptr;
    unsigned long long* d_m = nullptr;
    bool ok = cudaSetDevice(ctx->device) == cudaSuccess && cudaMalloc(&d_rho, bytes_rho) == cudaSuccess &&
              cudaMalloc(&d_zb, bytes_zb) == cudaSuccess && cudaMalloc(&d_zk, bytes_zk) == cudaSuccess && cudaMalloc(&d_m, bytes_m) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(d_rho, storage, bytes_rho, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(d_zb, z_list[bra], bytes_zb, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(d_zk, z_list[ket], bytes_zk, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(d_m, masks.data(), bytes_m, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    if (!ok) xr_set_error("%s_tensor: device allocation or upload failed: %s", ops, cudaGetErrorString(cudaGetLastError()));
    ok = ok && xr_density_tensor(ctx, ops, d_rho, nb, nk, d_zb, cb, d_zk, ck, reinterpret_cast<const uint64_t*>(d_m), n_elec[bra], ne,
                                 n_orbs, n_core, /*accumulate=*/1) == XR_OK;
    if (ok && (cudaMemcpyAsync(storage, d_rho, bytes_rho, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
               cudaStreamSynchronize(ctx->stream) != cudaSuccess)) {
        xr_set_error("%s_tensor: download failed: %s", ops, cudaGetErrorString(cudaGetLastError()));
        ok = false;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rho); cudaFree(d_zb); cudaFree(d_zk); cudaFree(d_m);
    return ok;
}

// The reference's entry points return void, so there is no error channel: a failed build must not pass as the zeros the
// caller allocated.  On any failure the message goes to stderr and the requested block is filled with NaN.
void legacy_density(const char* ops, double* storage, int64_t bra, int64_t ket, const int64_t* n_elec, const int64_t* n_states,
                    double* const* z_list, const int64_t* n_configs, int64_t* const* configs, int64_t n_orbs, int64_t n_core) {
    int64_t elements = 0;
    if (legacy_density_run(ops, storage, bra, ket, n_elec, n_states, z_list, n_configs, configs, n_orbs, n_core, &elements)) return;
    fprintf(stderr, "libxr_b200: %s (there is no CPU fallback; the block is filled with NaN)\n", xr_last_error());
    if (storage && elements == 0 && n_states && n_orbs > 0) {      // the failure came before the size was known
        int k = 0;
        while (ops[k]) ++k;
        elements = n_states[bra] * n_states[ket];
        for (int o = 0; o < k; ++o) elements *= 2 * n_orbs;
    }
    for (int64_t e = 0; storage && e < elements; ++e) storage[e] = NAN;
}

}  // namespace

#define XR_LEGACY_DENSITY(name, ops)                                                                                          \
    extern "C" void name(double* storage, int64_t bra_chg_idx, int64_t ket_chg_idx, int64_t* n_elec, int64_t* n_states,        \
                         double** z_list, int64_t* n_configs, int64_t** configs, int64_t n_orbs, int64_t n_core,               \
                         int64_t** combinatorics, int64_t n_threads) {                                                         \
        (void)combinatorics; (void)n_threads;                                                                                  \
        legacy_density(ops, storage, bra_chg_idx, ket_chg_idx, n_elec, n_states, z_list, n_configs, configs, n_orbs, n_core);  \
    }

XR_LEGACY_DENSITY(a_tensor, "a")
XR_LEGACY_DENSITY(c_tensor, "c")
XR_LEGACY_DENSITY(aa_tensor, "aa")
XR_LEGACY_DENSITY(cc_tensor, "cc")
XR_LEGACY_DENSITY(ca_tensor, "ca")
XR_LEGACY_DENSITY(caa_tensor, "caa")
XR_LEGACY_DENSITY(cca_tensor, "cca")
XR_LEGACY_DENSITY(ccaa_tensor, "ccaa")
